#!/bin/bash
# one block per SM (2368 envs = 148 x 16): single-block latency of the two long-chain variants
out=gpurun_out
for n in 2368 4096 4736; do
  python bench.py --task UsefulHound --num-envs $n --steps 200 --warmup 30 --ppo 0 --other-configs 0 > $out/r02r_seg_$n.json 2>/dev/null
  B2G_NO_SEGMENTS=1 python bench.py --task UsefulHound --num-envs $n --steps 200 --warmup 30 --ppo 0 --other-configs 0 > $out/r02r_whole_$n.json 2>/dev/null
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02r_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], f"{d['ms_per_step']*1e3:.1f}us {d['value']/1e6:.2f}M/s warm {d.get('ms_per_step_warm_l2',0)*1e3:.1f}us")
    except Exception as e:
        print(f, "ERR", e)
PY
