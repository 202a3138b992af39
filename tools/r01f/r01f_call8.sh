#!/bin/bash
# GPU call: shared-memory carve-out sweep for the long-chain physics kernel (UsefulHound)
out=gpurun_out; tag=${1:-r01m}
for c in none 40 50 60 none 40 50 60; do
  if [ $c = none ]; then unset B2G_CARVEOUT; else export B2G_CARVEOUT=$c; fi
  python bench.py --task UsefulHound --steps 200 --warmup 20 2>/dev/null | tail -n 1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('carveout $c', f\"{d['ms_per_step']*1e3:.1f}us warm {d['ms_per_step_warm_l2']*1e3:.1f}us\")" | tee -a $out/${tag}_carveout.log
done
