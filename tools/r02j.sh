#!/bin/bash
# ncu --set full captures (cold caches: --cache-control all is ncu's default) of the dominant kernels, after the plain run exited 0
out=gpurun_out
python bench.py --steps 30 --warmup 10 --other-configs 0 --ppo 0 > $out/r02j_plain.json 2>/dev/null || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_anymal_step -s 325 -c 2 -o $out/r02j_anymal -f python bench.py --steps 30 --warmup 10 --other-configs 0 --ppo 0 > $out/r02j_ncu1.log 2>&1
python bench.py --task UsefulHound --steps 20 --warmup 5 --preroll 100 --ppo 0 > $out/r02j_plain_uh.json 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_terrain_phys -s 110 -c 1 -o $out/r02j_usefulhound -f python bench.py --task UsefulHound --steps 20 --warmup 5 --preroll 100 --ppo 0 > $out/r02j_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_terrain -s 220 -c 2 -o $out/r02j_anymalterrain -f python bench.py --task AnymalTerrain --steps 20 --warmup 5 --preroll 100 --ppo 0 > $out/r02j_ncu3.log 2>&1
ls -la $out/*.ncu-rep
