#!/bin/bash
# A/B run of the experimental builds on the GPU box: parity suite + Anymal / AnymalTerrain bench per variant.
# usage: tools/ab_variants.sh TAG name1 name2 ...   (results in gpurun_out/TAG_<name>_*.{log,json})
tag=$1; shift
for v in "$@"; do
  export B2G_LIB_PATH=$PWD/build/variants/libb200gym_$v.so
  python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${tag}_${v}_tests.log
  python bench.py --steps 600 --warmup 60 > gpurun_out/${tag}_${v}_anymal.json 2> gpurun_out/${tag}_${v}_anymal.err
  python bench.py --task AnymalTerrain --steps 200 --warmup 20 > gpurun_out/${tag}_${v}_terrain.json 2>/dev/null
  python bench.py --num-envs 8192 --steps 300 --warmup 30 > gpurun_out/${tag}_${v}_anymal8192.json 2>/dev/null
  python bench.py --task UsefulHound --steps 100 --warmup 10 > gpurun_out/${tag}_${v}_usefulhound.json 2>/dev/null
done
unset B2G_LIB_PATH
python - "$tag" "$@" <<'PY'
import json, sys
tag, names = sys.argv[1], sys.argv[2:]
for v in names:
    t = open(f"gpurun_out/{tag}_{v}_tests.log").read().strip().splitlines()[-1:]
    row = [v, " ".join(t)]
    for k in ("anymal", "terrain", "anymal8192", "usefulhound"):
        try:
            d = json.loads(open(f"gpurun_out/{tag}_{v}_{k}.json").read().strip().splitlines()[-1])
            row.append(f"{k}: {d['ms_per_step']*1e3:.1f}us warm {d['ms_per_step_warm_l2']*1e3:.1f}us e2e {d['e2e']['ms_per_step']*1e3:.1f}us")
        except Exception as e:
            row.append(f"{k}: ERR {e}")
    print(" | ".join(row))
PY
