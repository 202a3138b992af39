#!/bin/bash
out=gpurun_out
python -m pytest tests/test_fused_update.py tests/test_ppo.py -m gpu -q 2>&1 | tail -150 > $out/r02h_tests_learner.log
grep -n "Error\|error\|passed\|failed" $out/r02h_tests_learner.log | head -20
