#!/bin/bash
# GPU call: parity suite, every task's bench line, PPO sanity on the rough-terrain task, launch lists + full captures (tag)
out=gpurun_out; tag=${1:-r01j}
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $out/${tag}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1
python bench.py > $out/${tag}_bench_plain.json 2> $out/${tag}_bench_plain.err
for t in AnymalTerrain HoundTerrain UsefulHound Houndarm Hound Cartpole; do
  python bench.py --task $t --steps 300 --warmup 30 > $out/${tag}_bench_$t.json 2>/dev/null
done
timeout 300 python tools/train_ppo.py --task AnymalTerrain --epochs 300 --cuda-graphs --fused-rollout --out $out/${tag}_ppo_anymal_terrain_graphs_fused_300epochs.json > $out/${tag}_ppo_at.log 2>&1
for t in AnymalTerrain UsefulHound; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/${tag}_launches_$t.csv python bench.py --task $t --steps 20 --warmup 5 > $out/${tag}_ncu_launches_$t.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_terrain_ -s 24 -c 2 -f -o $out/${tag}_full_AnymalTerrain python bench.py --task AnymalTerrain --steps 20 --warmup 5 > $out/${tag}_ncu_full_AnymalTerrain.log 2>&1
tail -3 $out/${tag}_tests.log; tail -1 $out/${tag}_smoke.log; tail -2 $out/${tag}_ppo_at.log
for f in $out/${tag}_bench_*.json; do python - $f <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split("/")[-1], f"{d['ms_per_step']*1e3:.1f}us warm {d['ms_per_step_warm_l2']*1e3:.1f}us {d['value']/1e6:.1f}M/s e2e {d['e2e']['ms_per_step']*1e3:.1f}us {d['e2e']['value']/1e6:.1f}M/s launches {d['gpu_launches']}")
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
