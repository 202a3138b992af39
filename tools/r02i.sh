#!/bin/bash
out=gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $out/r02i_bench_n2.json 2> $out/r02i_bench_n2.err
echo "rc=$?"; tail -c 1500 $out/r02i_bench_n2.json; tail -5 $out/r02i_bench_n2.err | cut -c1-300
