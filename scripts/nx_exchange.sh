# N-GPU hint generation: how the parities reach rank 0.  usage (under gpurun --gpus N): bash scripts/nx_exchange.sh N "<variants>"
cd $GRAFT_REPO_ROOT
N=$1
port=29600
run() {
  tag=$1; shift
  port=$((port + 1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 3 --no-search --no-other-configs --no-cpu-baseline "$@" 2>gpurun_out/n${N}_$tag.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$tag:', round(d['value'],1),'GB/s step',round(d['ms_per_step'],4),'ms kernel',round(d['roofline']['kernel_ms'],4),'per rank',[round(x,3) for x in d['roofline']['kernel_ms_per_rank']],'verified',d['verified_vs_oracle_prf'],'e2e',round(d['e2e']['ms_per_step'],3),'|',d['parallelism']['sharding'][-60:])" || tail -5 gpurun_out/n${N}_$tag.err | cut -c1-300
}
for v in $2; do
  case $v in
    pipe) run pipe --exchange pipe;;
    pipe_r0) run pipe_r0 --exchange pipe --relief 0;;
    pipe_r20) run pipe_r20 --exchange pipe --relief 0.2;;
    pipe_r30) run pipe_r30 --exchange pipe --relief 0.3;;
    p2p) run p2p --exchange p2p;;
    nccl) run nccl --exchange nccl;;
    pipe_hintset) run pipe_hintset --exchange pipe --sharding hintset;;
  esac
done
