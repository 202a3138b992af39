#!/bin/bash
out=gpurun_out
N=${1:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > $out/r02F_bench_n$N.json 2> $out/r02F_bench_n$N.err
echo "rc=$?"; python - $N <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads(open(f"gpurun_out/r02F_bench_n{n}.json").read().strip().splitlines()[-1])
print("N", d["n_gpus"], "value %.1fM ms %.4f e2e %.1fM (%.4f ms)" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["e2e"]["ms_per_step"]), d["e2e"]["cpu_affinity"], d["clocks"])
print(json.dumps(d.get("ppo_config5")))
PY
