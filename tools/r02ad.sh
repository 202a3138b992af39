#!/bin/bash
out=gpurun_out
N=${1:-8}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/train_ppo.py --task Anymal --num-envs 8192 --epochs 300 --tf32 --cuda-graphs --fused-rollout --fused-update --yaml --out $out/r02H_ppo_anymal_${N}gpu_8192envs_300epochs.json > $out/r02H_ppo_${N}gpu.log 2>&1
echo "rc=$?"; tail -3 $out/r02H_ppo_${N}gpu.log | cut -c1-600
