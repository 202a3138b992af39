#!/bin/bash
# GPU call: final confirmation of the round's HEAD: parity suite, smoke, default bench line + its ncu launch list, reference arm
out=gpurun_out; tag=${1:-r01n}
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $out/${tag}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1
python bench.py > $out/${tag}_bench_plain.json 2> $out/${tag}_bench_plain.err
python bench.py --impl reference --steps 10 --warmup 2 > $out/${tag}_bench_reference_arm.json 2>/dev/null
for t in AnymalTerrain HoundTerrain UsefulHound; do
  python bench.py --task $t --steps 300 --warmup 30 > $out/${tag}_bench_$t.json 2>/dev/null
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench_steps30.csv python bench.py --steps 30 --warmup 10 > $out/${tag}_ncu_launches.log 2>&1
tail -2 $out/${tag}_tests.log; tail -1 $out/${tag}_smoke.log
for f in $out/${tag}_bench_*.json; do python - $f <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split("/")[-1], f"{d['ms_per_step']*1e3:.1f}us {d['value']/1e6:.2f}M/s e2e {d['e2e']['value']/1e6:.2f}M/s launches {d['gpu_launches']} traffic {d.get('roofline',{}).get('traffic')}")
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
