#!/bin/bash
# GPU call: parity suite + rough-terrain benches + launch list on the current build (tag)
out=gpurun_out; tag=${1:-r01k}
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $out/${tag}_tests.log
for t in AnymalTerrain HoundTerrain UsefulHound; do
  python bench.py --task $t --steps 300 --warmup 30 > $out/${tag}_bench_$t.json 2>/dev/null
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/${tag}_launches_AnymalTerrain.csv python bench.py --task AnymalTerrain --steps 20 --warmup 5 > $out/${tag}_ncu_launches_AnymalTerrain.log 2>&1
tail -3 $out/${tag}_tests.log
for f in $out/${tag}_bench_*.json; do python - $f <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split("/")[-1], f"{d['ms_per_step']*1e3:.1f}us warm {d['ms_per_step_warm_l2']*1e3:.1f}us {d['value']/1e6:.1f}M/s e2e {d['e2e']['ms_per_step']*1e3:.1f}us {d['e2e']['value']/1e6:.1f}M/s launches {d['gpu_launches']}")
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
