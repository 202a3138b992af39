#!/bin/bash
# e2e scaling after moving the closing barrier out of the e2e wall-clock interval: N=1 and N=$1 on the same box, K=20 (driver) and K=200
N=$1; out=gpurun_out
for K in 20 200; do
  python bench.py --gpus 1 --steps $K --warmup 5 --ppo 0 --other-configs 0 > $out/r02R_n1_k$K.json 2>/dev/null
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + K % 97)) bench.py --gpus $N --steps $K --warmup 5 --ppo 0 --other-configs 0 > $out/r02R_n${N}_k$K.json 2>/dev/null
done
python - $N <<'PY'
import json, sys
N = sys.argv[1]
for K in (20, 200):
    a = json.loads(open(f"gpurun_out/r02R_n1_k{K}.json").read().strip().splitlines()[-1])
    b = json.loads(open(f"gpurun_out/r02R_n{N}_k{K}.json").read().strip().splitlines()[-1])
    print(f"K={K}: N=1 value {a['value']/1e6:.1f} e2e {a['e2e']['value']/1e6:.1f} | N={N} value {b['value']/1e6:.1f} e2e {b['e2e']['value']/1e6:.1f} | eff value {b['value']/a['value']/int(N):.3f} e2e {b['e2e']['value']/a['e2e']['value']/int(N):.3f}", b['e2e'].get('cpu_affinity'))
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --steps 20 --warmup 5 > $out/r02R_n${N}_driver.json 2> $out/r02R_n${N}_driver.err; echo driver-style rc=$?
tail -c 400 $out/r02R_n${N}_driver.json
