#!/bin/bash
out=gpurun_out
python -m pytest tests/test_fused_update.py tests/test_ppo.py -m gpu -q 2>&1 | tail -150 > $out/r02l_tests_learner.log
tail -5 $out/r02l_tests_learner.log
python bench.py --steps 200 --warmup 20 --other-configs 0 > $out/r02l_bench_fusedupdate.json 2> $out/r02l_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02l_bench_fusedupdate.json").read().strip().splitlines()[-1])
print("ppo fused", d.get("ppo_config5"))
PY
python - <<'PY' > gpurun_out/r02l_ppo_variants.json 2> gpurun_out/r02l_ppo_variants.err
import json, sys, time
sys.path.insert(0, ".")
import torch, bench
class D:  # no distributed
    pass
res = {}
for fu in (False, True):
    res["fused_update=%s" % fu] = bench.measure_ppo("cuda:0", 0, 1, None, num_envs=8192, epochs=20, warm=3, fused_update=fu)
    res["fused_update=%s_4096" % fu] = bench.measure_ppo("cuda:0", 0, 1, None, num_envs=4096, epochs=20, warm=3, fused_update=fu)
print(json.dumps(res, indent=1))
PY
tail -30 gpurun_out/r02l_ppo_variants.json
