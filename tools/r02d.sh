#!/bin/bash
out=gpurun_out
python -m pytest tests -m gpu -q 2>&1 > $out/r02d_tests_full.log
tail -40 $out/r02d_tests_full.log
python bench.py --steps 200 --warmup 20 > $out/r02d_bench.json 2> $out/r02d_bench.err
tail -c 600 $out/r02d_bench.err
