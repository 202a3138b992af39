#!/bin/bash
out=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -80 > $out/r02b_tests.log
tail -5 $out/r02b_tests.log
