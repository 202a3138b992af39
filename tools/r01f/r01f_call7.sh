#!/bin/bash
# GPU call: A/B of the 8-lane warp layout (default = legs in lanes 0-15 / upper chains chain-major, consec = consecutive lanes)
out=gpurun_out; tag=${1:-r01l}
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $out/${tag}_tests.log
for v in consec default consec default; do
  if [ $v = default ]; then unset B2G_LIB_PATH; else export B2G_LIB_PATH=$PWD/build/variants/libb200gym_$v.so; fi
  python bench.py --task UsefulHound --steps 300 --warmup 30 2>/dev/null | tail -n 1 >> $out/${tag}_${v}_bench_UsefulHound.json
done
unset B2G_LIB_PATH
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_terrain_phys -s 12 -c 1 -f -o $out/${tag}_full_UsefulHound python bench.py --task UsefulHound --steps 20 --warmup 5 > $out/${tag}_ncu_full_UsefulHound.log 2>&1
tail -3 $out/${tag}_tests.log
for f in $out/${tag}_*_bench_*.json; do python - $f <<'PY'
import json, sys
for l in open(sys.argv[1]).read().strip().splitlines():
    try:
        d = json.loads(l)
        print(sys.argv[1].split("/")[-1], f"{d['ms_per_step']*1e3:.1f}us warm {d['ms_per_step_warm_l2']*1e3:.1f}us {d['value']/1e6:.1f}M/s e2e {d['e2e']['ms_per_step']*1e3:.1f}us")
    except Exception as e:
        print(sys.argv[1], "ERR", e)
PY
done
