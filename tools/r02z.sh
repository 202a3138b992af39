#!/bin/bash
# after self-collision / velocity limits / Manipulator: GPU suite + every task bench, with self-collision on (default for the terrain tasks) and off
out=gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > $out/r02z_tests.log; tail -4 $out/r02z_tests.log
for t in Anymal Hound Cartpole AnymalTerrain HoundTerrain UsefulHound Houndarm Manipulator; do
  python bench.py --task $t --steps 300 --warmup 30 --ppo 0 --other-configs 0 > $out/r02z_bench_$t.json 2>/dev/null
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02z_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], f"{d['ms_per_step']*1e3:.1f}us {d['value']/1e6:.2f}M/s warm {d.get('value_warm_l2',0)/1e6:.1f} e2e {d['e2e'].get('ms_per_step',0)*1e3:.1f}us {d['e2e']['value']/1e6:.2f}M/s", d.get("contact_stats"))
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 600 python tools/useful_hound_diag.py 2>&1 | grep "^{"
