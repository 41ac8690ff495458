set -x
cd $GRAFT_REPO_ROOT
for v in "" "PM_HG_XBYTES=4" "PM_HG_XBYTES=2" "PM_HG_SERPENTINE=0" "PM_HG_TAIL_SPLIT=0" "PM_HG_SYNC=0" "PM_HG_NTAB=4" "PM_HG_WARPS=14" "PM_HG_WARPS=12"; do
  echo "== $v"; env $v python scripts/time_hintgen.py --iters 8 2>&1 | tail -1
done
echo "== shard 0/8"; python scripts/time_hintgen.py --iters 8 --shard 0/8 2>&1 | tail -1
echo "== shard 0/8 notail"; PM_HG_TAIL_SPLIT=0 python scripts/time_hintgen.py --iters 8 --shard 0/8 2>&1 | tail -1
echo "== batch 4 (partition-sharded rank: 2 parts ~)"; python scripts/time_hintgen.py --iters 8 --n 400228 --batch 4 2>&1 | tail -1
echo "== cfg0"; python scripts/time_hintgen.py --iters 8 --n 1048576 --entry-u64 4 --batch 0 --fail 40 2>&1 | tail -1
echo "== cfg0 xb4"; PM_HG_XBYTES=4 python scripts/time_hintgen.py --iters 8 --n 1048576 --entry-u64 4 --batch 0 --fail 40 2>&1 | tail -1
echo "== sift"; python scripts/time_hintgen.py --iters 8 --n 1000000 --entry-u64 80 2>&1 | tail -1
echo "== sift xb4"; PM_HG_XBYTES=4 python scripts/time_hintgen.py --iters 8 --n 1000000 --entry-u64 80 2>&1 | tail -1
python scripts/time_hintgen.py --iters 2 > gpurun_out/r2_plain_hg.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:hintgen_kernel -s 2 -c 1 -o gpurun_out/r2_hintgen_v4 python scripts/time_hintgen.py --iters 2 > gpurun_out/r2_ncu_hg.log 2>&1
