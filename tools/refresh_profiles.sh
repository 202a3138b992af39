#!/bin/bash
# Re-measure everything DESIGN.md quotes, on one GPU:  tools/refresh_profiles.sh TAG   -> gpurun_out/TAG_*.json / .csv
# (bench lines per task, env-count sweep, PPO curve with CUDA graphs + fused rollout policy, ncu launch list)
tag=${1:-r}
out=gpurun_out
python bench.py > $out/${tag}_bench_plain.json 2> $out/${tag}_bench_plain.err
for t in Hound Cartpole AnymalTerrain HoundTerrain UsefulHound Houndarm; do
  python bench.py --task $t --steps 300 --warmup 30 > $out/${tag}_bench_$t.json 2>/dev/null
done
echo "[" > $out/${tag}_sweep_envs.json
first=1
for n in 1024 2048 4096 8192 16384 32768 65536; do
  [ $first = 1 ] || echo "," >> $out/${tag}_sweep_envs.json
  first=0
  python bench.py --num-envs $n --steps 300 --warmup 30 2>/dev/null | tail -n 1 >> $out/${tag}_sweep_envs.json
done
echo "]" >> $out/${tag}_sweep_envs.json
python tools/train_ppo.py --task Anymal --epochs 1000 --cuda-graphs --fused-rollout --out $out/${tag}_ppo_anymal_graphs_fused_1000epochs.json > $out/${tag}_ppo.log 2>&1
timeout 600 python tools/train_ppo.py --task UsefulHound --epochs 40 --cuda-graphs --out $out/${tag}_ppo_usefulhound_graphs_40epochs.json > $out/${tag}_ppo_uh.log 2>&1
timeout 600 python tools/train_ppo.py --task AnymalTerrain --epochs 300 --cuda-graphs --out $out/${tag}_ppo_anymal_terrain_graphs_300epochs.json > $out/${tag}_ppo_at.log 2>&1
python tools/bench_policy.py > $out/${tag}_policy_bench.json 2> $out/${tag}_policy_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench_steps30.csv python bench.py --steps 30 --warmup 10 > $out/${tag}_ncu_launches.log 2>&1
python - "$tag" <<'PY'
import json, sys, glob
tag = sys.argv[1]
for f in sorted(glob.glob(f"gpurun_out/{tag}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], f"{d['ms_per_step']*1e3:.1f}us {d['value']/1e6:.1f}M/s e2e {d['e2e']['ms_per_step']*1e3:.1f}us {d['e2e']['value']/1e6:.1f}M/s")
    except Exception as e:
        print(f, "ERR", e)
try:
    for d in json.load(open(f"gpurun_out/{tag}_sweep_envs.json")):
        print("sweep", d["config"]["envs_per_gpu"], f"{d['ms_per_step']*1e3:.1f}us {d['value']/1e6:.1f}M/s e2e {d['e2e']['value']/1e6:.1f}M/s")
except Exception as e:
    print("sweep ERR", e)
for name in ("anymal_graphs_fused_1000epochs", "usefulhound_graphs_40epochs", "anymal_terrain_graphs_300epochs"):
    try:
        d = json.load(open(f"gpurun_out/{tag}_ppo_{name}.json"))
        print("ppo", name, d["mean_episode_reward"][-3:], d["wall_s"][-1], d["env_steps_per_sec_incl_learner"])
    except Exception as e:
        print("ppo ERR", name, e)
PY
