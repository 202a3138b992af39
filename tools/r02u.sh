#!/bin/bash
out=gpurun_out
for d in 2 3 4; do
  echo "== DIRECT=$d"
  B2G_PPO_DIRECT=$d python - <<'PY' 2>&1 | grep -v "Warning\|run_backward" | tail -3 | cut -c1-200
import json, sys
sys.path.insert(0, ".")
import torch, bench
r = bench.measure_ppo("cuda:0", 0, 1, None, num_envs=8192, epochs=10, warm=3, fused_update=True)
print(r["ms_per_iteration"], r["env_steps_per_sec_incl_learner"], r["update_capture_error"])
PY
done
