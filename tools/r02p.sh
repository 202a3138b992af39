#!/bin/bash
out=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_terrain_phys -s 110 -c 1 -o $out/r02p_usefulhound_seg -f python bench.py --task UsefulHound --steps 20 --warmup 5 --preroll 100 --ppo 0 --other-configs 0 > $out/r02p_ncu.log 2>&1
tail -3 $out/r02p_ncu.log
