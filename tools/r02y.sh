#!/bin/bash
out=gpurun_out
python -m pytest tests/test_tasks_gpu.py tests/test_kernels_gpu.py -m gpu -q -x -k "plausibility or velocity_limits or long_run" 2>&1 | tail -8
timeout 600 python tools/useful_hound_diag.py 2>&1 | grep "^{"
timeout 900 python tools/useful_hound_curves.py --epochs 300 --out $out/r02y_useful_hound_curves.json 2>&1 | grep "^refresh"
