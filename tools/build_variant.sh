#!/bin/bash
# Build an experimental variant of libb200gym.so for A/B timing:  tools/build_variant.sh NAME [extra nvcc flags...]
# -> build/variants/libb200gym_NAME.so ; run with  B2G_LIB_PATH=build/variants/libb200gym_NAME.so python bench.py
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -use_fast_math -Xcompiler -fPIC -shared --threads 2 \
  -Iinclude -Iisaacgymenv_b200/csrc "$@" isaacgymenv_b200/csrc/b200gym.cu isaacgymenv_b200/csrc/b2g_policy.cu \
  -o build/variants/libb200gym_$name.so
echo built build/variants/libb200gym_$name.so
