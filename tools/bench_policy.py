#!/usr/bin/env python
"""Time the fused tcgen05 policy forward (b2g_policy_forward) against torch evaluations of the same network.
    python tools/bench_policy.py [--rows 4096 32768 98304] > gpurun_out/policy_bench.json
Timing: CUDA events on the launch stream around ITERS back-to-back launches after warm-up (the working set -- 106 KB of
weights plus rows x 48 floats -- is L2 resident in the rollout as well, the observations having just been written by the step
kernel).  flops = 2 x rows x sum(in x out) over the five linear maps (useful columns only)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, nargs="+", default=[4096, 32768, 98304])
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--net", default="anymal", choices=["anymal", "terrain", "usefulhound"],
                    help="anymal: 48->[256,128,64]->12 (resident weights); terrain: 188->[512,256,128]->12, usefulhound: 204->[512,256,128]->18 (streamed weights)")
    args = ap.parse_args()
    import torch

    from isaacgymenv_b200.learning.fused_policy import FusedPolicy
    from isaacgymenv_b200.learning.ppo import ActorCritic, RunningMeanStd

    dev = "cuda:0"
    units, n_obs, n_act = {"anymal": ((256, 128, 64), 48, 12), "terrain": ((512, 256, 128), 188, 12), "usefulhound": ((512, 256, 128), 204, 18)}[args.net]
    torch.manual_seed(0)
    model = ActorCritic(n_obs, n_act, units).to(dev)
    rms = RunningMeanStd((n_obs,)).to(dev)
    pol = FusedPolicy(n_obs, n_act, units, dev)
    pol.sync(model, rms)
    macs = n_obs * units[0] + units[0] * units[1] + units[1] * units[2] + units[2] * (n_act + 1)
    out = {"gpu": torch.cuda.get_device_name(0), "network": f"{n_obs}->{units}->{n_act}+1, ELU", "iters": args.iters, "rows": {}}

    def timed(fn):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.iters * 1e3     # us

    for rows in args.rows:
        obs = torch.randn(rows, n_obs, device=dev)
        mu, val = torch.empty(rows, n_act, device=dev), torch.empty(rows, device=dev)
        t_fused = timed(lambda: pol.forward(obs, mu, val))

        @torch.no_grad()
        def torch_fp32():
            return model(rms.normalize(obs))

        @torch.no_grad()
        def torch_bf16():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return model(rms.normalize(obs))

        t32, t16 = timed(torch_fp32), timed(torch_bf16)
        g = torch.cuda.CUDAGraph()
        with torch.no_grad():
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                torch_fp32()
            torch.cuda.current_stream().wait_stream(s)
            with torch.cuda.graph(g):
                torch_fp32()
        t32g = timed(g.replay)
        flops = 2.0 * rows * macs
        out["rows"][str(rows)] = {"fused_tcgen05_us": t_fused, "torch_fp32_us": t32, "torch_fp32_cuda_graph_us": t32g, "torch_bf16_autocast_us": t16,
                                  "fused_tflops": flops / (t_fused * 1e-6) / 1e12, "speedup_vs_torch_fp32": t32 / t_fused,
                                  "speedup_vs_torch_fp32_graph": t32g / t_fused}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
