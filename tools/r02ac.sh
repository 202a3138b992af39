#!/bin/bash
python -m pytest tests/test_fused_update.py tests/test_ppo.py -m gpu -q -x 2>&1 | tail -3
python - <<'PY' 2>&1 | grep -v Warning | tail -3
import json, sys
sys.path.insert(0, ".")
import torch, bench
for n in (8192, 4096):
    r = bench.measure_ppo("cuda:0", 0, 1, None, num_envs=n, epochs=20, warm=3, fused_update=True)
    print(n, round(r["ms_per_iteration"], 2), round(r["env_steps_per_sec_incl_learner"] / 1e6, 2), r["update_capture_error"])
PY
