#!/bin/bash
out=gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for t in Anymal Hound; do
  python bench.py --task $t --steps 300 --warmup 30 --ppo 0 --other-configs 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$t', round(d['ms_per_step']*1e3,1), round(d['e2e']['ms_per_step']*1e3,1))"
done
timeout 300 python tools/train_ppo.py --task Anymal --epochs 300 --tf32 --cuda-graphs --fused-rollout --fused-update --yaml 2>&1 | tail -1 | cut -c1-500
