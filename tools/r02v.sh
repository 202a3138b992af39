#!/bin/bash
# Manipulator + everything touched since the last full run: GPU suite, arm benches
out=gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > $out/r02v_tests.log; tail -6 $out/r02v_tests.log
for t in Manipulator Houndarm; do
  python bench.py --task $t --steps 300 --warmup 30 --ppo 0 --other-configs 0 > $out/r02v_bench_$t.json 2>$out/r02v_bench_$t.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02v_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], f"{d['ms_per_step']*1e3:.1f}us {d['value']/1e6:.2f}M/s warm {d.get('value_warm_l2',0)/1e6:.1f} e2e {d['e2e'].get('ms_per_step',0)*1e3:.1f}us {d['e2e']['value']/1e6:.2f}M/s", d.get("contact_stats"))
    except Exception as e:
        print(f, "ERR", e)
PY
