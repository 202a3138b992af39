#!/usr/bin/env python
"""Stand-alone timing of the update's hand-written kernels at the minibatch shapes of cfg/train/AnymalPPO.yaml (32768 rows; hidden 256 / 128 / 64;
12 actions), warm L2, CUDA events around 200 back-to-back launches -- the regime they run in inside the update graph.
    python tools/bench_learner_kernels.py > gpurun_out/learner_kernels.json"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from isaacgymenv_b200 import _lib  # noqa: E402
from isaacgymenv_b200.learning import fused_update as fu  # noqa: E402

lib = fu._protos()
dev = torch.device("cuda:0")
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def timed(fn, reps=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


out = {}
rows = 32768
for cols in (256, 128, 64):
    dh, h = torch.randn(rows, cols, device=dev), torch.randn(rows, cols, device=dev)
    dz, dz16 = torch.empty_like(h), torch.empty(rows, cols, device=dev, dtype=torch.bfloat16)
    db = torch.empty(cols, device=dev)
    ws = torch.empty(int(lib.b2g_mlp_elu_backward_workspace_floats(rows, cols)), device=dev)
    us = timed(lambda: _lib.check(lib.b2g_mlp_elu_backward(p(dh), p(h), p(dz), p(db), p(ws), rows, cols, p(dz16), st), "elu_bwd"))
    byts = rows * cols * (4 + 4 + 4 + 2)
    out[f"elu_backward_{cols}"] = {"us": round(us, 2), "GB/s": round(byts / us * 1e-3, 1)}
    z, b = torch.randn(rows, cols, device=dev), torch.randn(cols, device=dev)
    us = timed(lambda: _lib.check(lib.b2g_mlp_bias_elu(p(z), p(b), rows, cols, p(dz16), st), "bias_elu"))
    out[f"bias_elu_{cols}"] = {"us": round(us, 2), "GB/s": round(rows * cols * 10 / us * 1e-3, 1)}
H, A = 64, 12
h, dmu, dv = torch.randn(rows, H, device=dev), torch.randn(rows, A, device=dev), torch.randn(rows, device=dev)
w_mu, w_v = torch.randn(A, H, device=dev), torch.randn(1, H, device=dev)
dh, cat = torch.empty_like(h), torch.empty(A + 1, H + 1, device=dev)
ws = torch.empty(int(lib.b2g_mlp_heads_backward_workspace_floats(rows, H, A)), device=dev)
us = timed(lambda: _lib.check(lib.b2g_mlp_heads_backward(p(h), p(dmu), p(dv), p(w_mu), p(w_v), rows, H, A, p(dh), p(cat), p(ws), st), "heads_bwd"))
out["heads_backward_64x12"] = {"us": round(us, 2), "GB/s": round(rows * (H * 8 + (A + 1) * 4) / us * 1e-3, 1)}
# reference values for the check
dy = torch.cat([dmu, dv[:, None]], 1)
ref_dh = dy @ torch.cat([w_mu, w_v], 0)
ref_cat = dy.t() @ torch.cat([h, torch.ones(rows, 1, device=dev)], 1)
out["heads_backward_64x12"]["max_rel_err_dh"] = float(((dh - ref_dh).abs().max() / ref_dh.abs().max()))
out["heads_backward_64x12"]["max_rel_err_dw"] = float(((cat - ref_cat).abs().max() / ref_cat.abs().max()))
print(json.dumps(out, indent=1))
