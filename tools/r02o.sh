#!/bin/bash
# segment variant <8,3> of the long-chain kernels: GPU tests, bench and A/B against the whole-chain kernels
out=gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $out/r02o_tests.log; tail -3 $out/r02o_tests.log
python bench.py --task UsefulHound --steps 300 --warmup 30 --ppo 0 --other-configs 0 > $out/r02o_bench_UsefulHound.json 2>/dev/null
B2G_NO_SEGMENTS=1 python bench.py --task UsefulHound --steps 300 --warmup 30 --ppo 0 --other-configs 0 > $out/r02o_bench_UsefulHound_wholechain.json 2>/dev/null
python bench.py --task UsefulHound --num-envs 8192 --steps 300 --warmup 30 --ppo 0 --other-configs 0 > $out/r02o_bench_UsefulHound_8192.json 2>/dev/null
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02o_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], f"{d['ms_per_step']*1e3:.1f}us {d['value']/1e6:.2f}M/s warm {d.get('value_warm_l2',0)/1e6:.1f} e2e {d['e2e'].get('ms_per_step',0)*1e3:.1f}us {d['e2e']['value']/1e6:.2f}M/s", d.get("contact_stats"))
    except Exception as e:
        print(f, "ERR", e)
PY
