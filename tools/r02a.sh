#!/bin/bash
# round-2 GPU pass A: whole GPU suite, the driver's bench invocations, default bench, launch list
out=gpurun_out
python -m pytest tests -m gpu -q --deselect tests/test_ppo.py 2>&1 | tail -40 > $out/r02a_tests.log
python -m pytest tests/test_ppo.py -m gpu -q 2>&1 | tail -15 > $out/r02a_tests_ppo.log
python bench.py --impl reference --steps 20 --warmup 5 > $out/r02a_bench_reference.json 2> $out/r02a_bench_reference.err
python bench.py --steps 20 --warmup 5 > $out/r02a_bench_driver.json 2> $out/r02a_bench_driver.err
python bench.py > $out/r02a_bench_plain.json 2> $out/r02a_bench_plain.err
B2G_CONTACT_STATS=0 python bench.py --other-configs 0 > $out/r02a_bench_nostats.json 2> $out/r02a_bench_nostats.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/r02a_launches_bench_steps30.csv python bench.py --steps 30 --warmup 10 --other-configs 0 > $out/r02a_ncu_launches.log 2>&1
tail -3 $out/r02a_tests.log
