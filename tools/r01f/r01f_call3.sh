#!/bin/bash
# GPU call: parity suite + benches of every task on the current build (tag), optional ncu launch list
out=gpurun_out; tag=${1:-r01h}
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $out/${tag}_tests.log
python bench.py --steps 1000 --warmup 100 > $out/${tag}_bench_plain.json 2> $out/${tag}_bench_plain.err
for t in AnymalTerrain HoundTerrain UsefulHound Houndarm Hound Cartpole; do
  python bench.py --task $t --steps 300 --warmup 30 > $out/${tag}_bench_$t.json 2>/dev/null
done
for t in AnymalTerrain UsefulHound; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/${tag}_launches_$t.csv python bench.py --task $t --steps 20 --warmup 5 > $out/${tag}_ncu_launches_$t.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_terrain_phys -s 12 -c 1 -f -o $out/${tag}_full_UsefulHound python bench.py --task UsefulHound --steps 20 --warmup 5 > $out/${tag}_ncu_full_UsefulHound.log 2>&1
tail -3 $out/${tag}_tests.log
for f in $out/${tag}_bench_*.json; do python - $f <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split("/")[-1], f"{d['ms_per_step']*1e3:.1f}us warm {d['ms_per_step_warm_l2']*1e3:.1f}us {d['value']/1e6:.1f}M/s e2e {d['e2e']['ms_per_step']*1e3:.1f}us launches {d['gpu_launches']}")
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
