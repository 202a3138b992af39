#!/bin/bash
out=gpurun_out
python tools/ppo_profile.py > $out/r02s_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/r02s_launches_ppo_iteration_fused.csv python tools/ppo_profile.py > $out/r02s_ncu_ppo.log 2>&1
tail -2 $out/r02s_ncu_ppo.log
