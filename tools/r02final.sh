#!/bin/bash
# round-2 final refresh: every number the docs quote, one GPU
out=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > $out/r02F_tests.log; tail -3 $out/r02F_tests.log
python bench.py > $out/r02F_bench_plain.json 2> $out/r02F_bench_plain.err
python bench.py --impl reference --steps 20 --warmup 5 > $out/r02F_bench_reference.json 2>/dev/null
python bench.py --steps 20 --warmup 5 > $out/r02F_bench_driver.json 2>/dev/null
for t in Hound Cartpole AnymalTerrain HoundTerrain UsefulHound Houndarm Manipulator; do
  python bench.py --task $t --steps 300 --warmup 30 --ppo 0 > $out/r02F_bench_$t.json 2>/dev/null
done
echo "[" > $out/r02F_sweep_envs.json
first=1
for n in 1024 2048 4096 8192 16384 32768 65536; do
  [ $first = 1 ] || echo "," >> $out/r02F_sweep_envs.json
  first=0
  python bench.py --num-envs $n --steps 300 --warmup 30 --ppo 0 2>/dev/null | tail -n 1 >> $out/r02F_sweep_envs.json
done
echo "]" >> $out/r02F_sweep_envs.json
python tools/train_ppo.py --task Anymal --epochs 1000 --tf32 --cuda-graphs --fused-rollout --fused-update --yaml --out $out/r02F_ppo_anymal_fused_update_1000epochs.json > $out/r02F_ppo_anymal.log 2>&1
timeout 600 python tools/train_ppo.py --task AnymalTerrain --epochs 300 --tf32 --cuda-graphs --fused-rollout --fused-update --yaml --out $out/r02F_ppo_anymal_terrain_separate_300epochs.json > $out/r02F_ppo_at.log 2>&1
timeout 300 python tools/train_ppo.py --task Cartpole --num-envs 512 --epochs 100 --tf32 --cuda-graphs --fused-update --yaml --out $out/r02F_ppo_cartpole_100epochs.json > $out/r02F_ppo_cp.log 2>&1
timeout 300 python tools/train_ppo.py --task Manipulator --num-envs 8192 --epochs 150 --tf32 --cuda-graphs --fused-rollout --fused-update --yaml --out $out/r02F_ppo_manipulator_150epochs.json > $out/r02F_ppo_mp.log 2>&1
timeout 300 python tools/train_ppo.py --task Houndarm --num-envs 8192 --epochs 150 --tf32 --cuda-graphs --fused-rollout --fused-update --yaml --out $out/r02F_ppo_houndarm_150epochs.json > $out/r02F_ppo_ha.log 2>&1
timeout 900 python tools/useful_hound_curves.py --epochs 300 --out $out/r02F_useful_hound_curves.json 2>&1 | grep "^refresh" > $out/r02F_useful_hound_curves.log
python - <<'PY' > gpurun_out/r02F_ppo_variants.json 2> gpurun_out/r02F_ppo_variants.err
import json, sys
sys.path.insert(0, ".")
import torch, bench
res = {}
for fu in (False, True):
    res["fused_update=%s" % fu] = bench.measure_ppo("cuda:0", 0, 1, None, num_envs=8192, epochs=20, warm=3, fused_update=fu)
    res["fused_update=%s_4096" % fu] = bench.measure_ppo("cuda:0", 0, 1, None, num_envs=4096, epochs=20, warm=3, fused_update=fu)
print(json.dumps(res, indent=1))
PY
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02F_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], f"{d['ms_per_step']*1e3:.1f}us {d['value']/1e6:.2f}M/s warm {d.get('value_warm_l2',0)/1e6:.1f} e2e {d['e2e'].get('ms_per_step',0)*1e3:.1f}us {d['e2e']['value']/1e6:.2f}M/s", d.get("contact_stats"), d.get("ppo_config5", {}).get("env_steps_per_sec_incl_learner") if isinstance(d.get("ppo_config5"), dict) else None)
    except Exception as e:
        print(f, "ERR", e)
try:
    for d in json.load(open("gpurun_out/r02F_sweep_envs.json")):
        print("sweep", d["config"]["envs_per_gpu"], f"{d['ms_per_step']*1e3:.1f}us {d['value']/1e6:.1f}M/s e2e {d['e2e']['value']/1e6:.1f}M/s")
except Exception as e:
    print("sweep ERR", e)
for name in ("anymal_fused_update_1000epochs", "anymal_terrain_separate_300epochs", "cartpole_100epochs", "manipulator_150epochs", "houndarm_150epochs"):
    try:
        d = json.load(open(f"gpurun_out/r02F_ppo_{name}.json"))
        print("ppo", name, d["mean_episode_reward"][-3:], d["mean_episode_length"][-1], d["wall_s"][-1], d["env_steps"][-1])
    except Exception as e:
        print("ppo ERR", name, e)
try:
    v = json.load(open("gpurun_out/r02F_ppo_variants.json"))
    for k, r in v.items(): print("variants", k, round(r["ms_per_iteration"], 2), round(r["env_steps_per_sec_incl_learner"] / 1e6, 2))
except Exception as e:
    print("variants ERR", e)
print(open("gpurun_out/r02F_useful_hound_curves.log").read())
PY
