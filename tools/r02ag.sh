#!/bin/bash
# round-end rehearsal on a fresh box: smoke, GPU suite twice (flakiness), driver-style bench (both arms), default bench wall time
out=gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) 2>&1 | tail -5 > $out/r02P_smoke.log
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > $out/r02P_tests1.log
python -m pytest tests -m gpu -q -x -p no:randomly 2>&1 | tail -6 > $out/r02P_tests2.log
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $out/r02P_bench_ref.json 2> $out/r02P_bench_ref.err
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > $out/r02P_bench_driver.json 2> $out/r02P_bench_driver.err
( time python bench.py ) > $out/r02P_bench_default.json 2> $out/r02P_bench_default.err
tail -3 $out/r02P_smoke.log $out/r02P_tests1.log $out/r02P_tests2.log; tail -4 $out/r02P_bench_ref.err $out/r02P_bench_driver.err $out/r02P_bench_default.err
