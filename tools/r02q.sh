#!/bin/bash
# block re-alignment barriers inside the sub-step: B2G_BLOCK_ALIGN bit mask sweep
out=gpurun_out
for t in Anymal AnymalTerrain UsefulHound Hound; do
  for ba in 0 1 3 7; do
    B2G_BLOCK_ALIGN=$ba python bench.py --task $t --steps 300 --warmup 30 --ppo 0 --other-configs 0 > $out/r02q_bench_${t}_ba$ba.json 2>/dev/null
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02q_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], f"{d['ms_per_step']*1e3:.1f}us {d['value']/1e6:.2f}M/s warm {d.get('value_warm_l2',0)/1e6:.1f} e2e {d['e2e'].get('ms_per_step',0)*1e3:.1f}us")
    except Exception as e:
        print(f, "ERR", e)
PY
