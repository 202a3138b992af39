#!/bin/bash
# does self-collision against the base (default for the rough-terrain tasks) change what PPO learns?
out=gpurun_out
for t in AnymalTerrain HoundTerrain; do
  for sc in 1 0; do
    timeout 400 python tools/train_ppo.py --task $t --epochs 300 --tf32 --cuda-graphs --fused-rollout --fused-update --yaml --self-collision $sc --out $out/r02I_ppo_${t}_sc$sc.json 2>&1 | tail -1 | cut -c1-400
  done
done
timeout 300 python tools/train_ppo.py --task Hound --epochs 500 --tf32 --cuda-graphs --fused-rollout --fused-update --yaml --out $out/r02I_ppo_Hound.json 2>&1 | tail -1 | cut -c1-400
