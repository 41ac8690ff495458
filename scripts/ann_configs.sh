# private-ANN throughput of bench.py's lock-step measurement for several (lanes, groups) shapes
cd $GRAFT_REPO_ROOT
for cfg in "64 2" "42 3" "32 4" "24 5"; do
  set -- $cfg
  python bench.py --steps 3 --warmup 3 --no-other-configs --no-cpu-baseline --search-lanes $1 --search-groups $2 --search-queries 5040 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=d['private_ann']
print('lanes $1 groups $2:', round(p['queries_per_s']), 'q/s incl maint,', round(p['queries_per_s_excl_maintenance']), 'excl; queries', p['queries'], 'timed', round(p['timed_region_s'],3))"
done
