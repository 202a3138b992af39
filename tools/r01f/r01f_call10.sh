#!/bin/bash
# GPU call: threads per environment in k_terrain_post (variants built with tools/build_variant.sh subN -DB2G_POST_SUB=N)
out=gpurun_out; tag=${1:-r01o}
for v in default sub32 sub8 default sub32 sub8; do
  if [ $v = default ]; then unset B2G_LIB_PATH; else export B2G_LIB_PATH=$PWD/build/variants/libb200gym_$v.so; fi
  python bench.py --task AnymalTerrain --steps 300 --warmup 30 2>/dev/null | tail -n 1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('post threads/env $v', f\"{d['ms_per_step']*1e3:.1f}us warm {d['ms_per_step_warm_l2']*1e3:.1f}us e2e {d['e2e']['ms_per_step']*1e3:.1f}us\")" | tee -a $out/${tag}_post_sub.log
done
