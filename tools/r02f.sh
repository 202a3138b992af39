#!/bin/bash
out=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/r02f_launches_ppo_iteration.csv python tools/ppo_profile.py > $out/r02f_ncu_ppo.log 2>&1
tail -2 $out/r02f_ncu_ppo.log
