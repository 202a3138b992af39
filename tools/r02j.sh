#!/bin/bash
# ncu --set full captures (cold caches: --cache-control all is ncu's default) of the dominant kernels, after the plain run exited 0
out=gpurun_out
python bench.py --steps 30 --warmup 10 --other-configs 0 --ppo 0 > $out/r02G_plain.json 2>/dev/null || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_anymal_step -s 325 -c 2 -o $out/r02G_anymal -f python bench.py --steps 30 --warmup 10 --other-configs 0 --ppo 0 > $out/r02G_ncu1.log 2>&1
python bench.py --task UsefulHound --steps 20 --warmup 5 --preroll 100 --ppo 0 > $out/r02G_plain_uh.json 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_terrain_phys -s 110 -c 1 -o $out/r02G_usefulhound -f python bench.py --task UsefulHound --steps 20 --warmup 5 --preroll 100 --ppo 0 > $out/r02G_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_terrain -s 220 -c 2 -o $out/r02G_anymalterrain -f python bench.py --task AnymalTerrain --steps 20 --warmup 5 --preroll 100 --ppo 0 > $out/r02G_ncu3.log 2>&1
for k in anymal usefulhound anymalterrain; do python tools/ncu_phases.py $out/r02G_$k.ncu-rep $out/r02G_ncu_phases_$k.csv > /dev/null 2>&1; rm -f $out/r02G_$k.ncu-rep; done
ls $out | grep r02G
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/r02G_launches_bench_steps30.csv python bench.py --steps 30 --warmup 10 --other-configs 0 > $out/r02G_ncu_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/r02G_launches_ppo_iteration_fused.csv python tools/ppo_profile.py > $out/r02G_ncu_ppo.log 2>&1
tail -2 $out/r02G_ncu_ppo.log
