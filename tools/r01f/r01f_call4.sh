#!/bin/bash
# GPU call: A/B of the long-chain link-state placement (default = per-thread local arrays, lsh = shared memory)
out=gpurun_out; tag=${1:-r01i}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $out/${tag}_tests.log
for v in default lsh; do
  if [ $v = default ]; then unset B2G_LIB_PATH; else export B2G_LIB_PATH=$PWD/build/variants/libb200gym_$v.so; fi
  [ $v = default ] || python -m pytest tests -m gpu -x -q -k "useful or houndarm or Houndarm or UsefulHound or arm" 2>&1 | tail -3 > $out/${tag}_${v}_tests.log
  for t in UsefulHound Houndarm; do
    python bench.py --task $t --steps 300 --warmup 30 > $out/${tag}_${v}_bench_$t.json 2>/dev/null
  done
done
unset B2G_LIB_PATH
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_terrain_phys -s 12 -c 1 -f -o $out/${tag}_full_UsefulHound python bench.py --task UsefulHound --steps 20 --warmup 5 > $out/${tag}_ncu_full_UsefulHound.log 2>&1
tail -2 $out/${tag}_tests.log; tail -2 $out/${tag}_lsh_tests.log
for f in $out/${tag}_*_bench_*.json; do python - $f <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split("/")[-1], f"{d['ms_per_step']*1e3:.1f}us warm {d['ms_per_step_warm_l2']*1e3:.1f}us {d['value']/1e6:.1f}M/s e2e {d['e2e']['ms_per_step']*1e3:.1f}us")
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
