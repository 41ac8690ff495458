# 8-GPU hint generation: how the parities reach rank 0 (run under gpurun --gpus 8)
cd $GRAFT_REPO_ROOT
run() {
  tag=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 295$((RANDOM % 90 + 10)) bench.py --gpus 8 --steps 20 --warmup 3 --no-search --no-other-configs --no-cpu-baseline $EXTRA 2>gpurun_out/n8_$tag.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$tag:', round(d['value'],1),'GB/s step',round(d['ms_per_step'],4),'ms kernel',round(d['roofline']['kernel_ms'],4),'verified',d['verified_vs_oracle_prf'],'e2e',round(d['e2e']['ms_per_step'],3))"
}
EXTRA="" run p2p A=1
EXTRA="--exchange nccl" run nccl A=1
EXTRA="--exchange ce --pieces 4" run ce4 A=1
EXTRA="--exchange ce --pieces 8" run ce8 A=1
EXTRA="" run p2p_nosync PM_HG_SYNC=0
EXTRA="--sharding hintset" run p2p_hintset A=1
