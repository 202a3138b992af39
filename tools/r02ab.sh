#!/bin/bash
out=gpurun_out
for i in 1 2; do
for t in Anymal AnymalTerrain; do
  python bench.py --task $t --steps 300 --warmup 30 --ppo 0 --other-configs 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$t', round(d['ms_per_step']*1e3,1), round(d['e2e']['ms_per_step']*1e3,1))"
done
done
