#!/bin/bash
# learner: direct gradient stores, gather kernel, faster column-sum finalize
out=gpurun_out
python -m pytest tests/test_fused_update.py tests/test_ppo.py -m gpu -q -x 2>&1 | tail -40 > $out/r02t_tests_learner.log
tail -5 $out/r02t_tests_learner.log
python - <<'PY' > gpurun_out/r02t_ppo_variants.json 2> gpurun_out/r02t_ppo_variants.err
import json, sys
sys.path.insert(0, ".")
import torch, bench
res = {}
res["fused_update=True"] = bench.measure_ppo("cuda:0", 0, 1, None, num_envs=8192, epochs=20, warm=3, fused_update=True)
res["fused_update=True_4096"] = bench.measure_ppo("cuda:0", 0, 1, None, num_envs=4096, epochs=20, warm=3, fused_update=True)
print(json.dumps(res, indent=1))
PY
grep -n "ms_per_iteration\|env_steps_per_sec\|capture_error\|reward" gpurun_out/r02t_ppo_variants.json; tail -3 gpurun_out/r02t_ppo_variants.err
