#!/bin/bash
# final rehearsal of the round: smoke, GPU suite, driver-style bench (both arms), default bench, 1000-epoch Anymal curve, PPO launch list
out=gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) 2>&1 | tail -n 5 > $out/r02W_smoke.log
python -m pytest tests -m gpu -q -x 2>&1 | tail -n 6 > $out/r02W_tests.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $out/r02W_bench_ref.json 2> /dev/null
python bench.py --gpus 1 --steps 20 --warmup 5 > $out/r02W_bench_driver.json 2> $out/r02W_bench_driver.err
python bench.py > $out/r02W_bench_default.json 2> /dev/null
timeout 300 python tools/train_ppo.py --task Anymal --epochs 1000 --tf32 --cuda-graphs --fused-rollout --fused-update --yaml > $out/r02W_ppo_anymal_1000epochs.json 2> $out/r02W_ppo.err
python tools/ppo_profile.py > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 4000 --csv --log-file $out/r02W_launches_ppo_iteration.csv python tools/ppo_profile.py > /dev/null 2>&1
tail -n 2 $out/r02W_smoke.log $out/r02W_tests.log; tail -c 300 $out/r02W_ppo_anymal_1000epochs.json
