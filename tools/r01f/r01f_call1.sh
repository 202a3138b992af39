#!/bin/bash
# GPU call: confirm HEAD (tests, smoke, bench) + launch lists / full captures of the rough-terrain kernels
out=gpurun_out; tag=${1:-r01f}
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $out/${tag}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> $out/${tag}_smoke.log
python bench.py > $out/${tag}_bench_plain.json 2> $out/${tag}_bench_plain.err
python bench.py --task AnymalTerrain --steps 300 --warmup 30 > $out/${tag}_bench_AnymalTerrain.json 2>/dev/null
python bench.py --task UsefulHound --steps 100 --warmup 10 > $out/${tag}_bench_UsefulHound.json 2>/dev/null
for t in AnymalTerrain UsefulHound; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/${tag}_launches_$t.csv python bench.py --task $t --steps 20 --warmup 5 > $out/${tag}_ncu_launches_$t.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_terrain_phys -s 12 -c 1 -f -o $out/${tag}_full_$t python bench.py --task $t --steps 20 --warmup 5 > $out/${tag}_ncu_full_$t.log 2>&1
done
tail -3 $out/${tag}_tests.log; tail -2 $out/${tag}_smoke.log; cat $out/${tag}_bench_plain.json | cut -c1-300
