#!/bin/bash
out=gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_physics_gpu.py tests/test_tasks_gpu.py -m gpu -q -x 2>&1 | tail -4
timeout 600 python tools/useful_hound_diag.py 2>&1 | grep "^{"
for t in Hound HoundTerrain UsefulHound; do
  python bench.py --task $t --steps 300 --warmup 30 --ppo 0 --other-configs 0 > $out/r02aa_bench_$t.json 2>/dev/null
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02aa_bench_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split("/")[-1], f"{d['ms_per_step']*1e3:.1f}us {d['value']/1e6:.2f}M/s", d.get("contact_stats"))
PY
