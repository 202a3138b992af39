"""Which cuBLAS path is fastest for the update's weight-gradient GEMMs dW = dZ^T X (K = 32768 minibatch rows, small M, N)?"""
import json, sys, torch
torch.backends.cuda.matmul.allow_tf32 = True
dev = "cuda:0"
B = 32768
shapes = [(256, 48), (128, 256), (64, 128), (13, 64)]          # (cols of dZ, cols of X)
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
out = {}
for M, N in shapes:
    dz, x = torch.randn(B, M, device=dev), torch.randn(B, N, device=dev)
    dzb, xb = dz.bfloat16(), x.bfloat16()
    dzt = dz.t().contiguous()
    xt = x.t().contiguous()
    r = {}
    r["tf32 mm(dz.t(), x)"] = t(lambda: torch.mm(dz.t(), x))
    r["tf32 mm(x.t(), dz).t()"] = t(lambda: torch.mm(x.t(), dz))
    r["tf32 mm(dzt_contig, x)"] = t(lambda: torch.mm(dzt, x))
    r["tf32 mm(dzt_contig, xt_contig.t())"] = t(lambda: torch.mm(dzt, xt.t()))
    r["bf16 mm(dz.t(), x) pre-cast"] = t(lambda: torch.mm(dzb.t(), xb))
    r["bf16 incl casts"] = t(lambda: torch.mm(dz.bfloat16().t(), x.bfloat16()))
    r["tf32 split 8 x bmm"] = t(lambda: torch.bmm(dz.view(8, B // 8, M).transpose(1, 2), x.view(8, B // 8, N)).sum(0))
    r["tf32 split 32 x bmm"] = t(lambda: torch.bmm(dz.view(32, B // 32, M).transpose(1, 2), x.view(32, B // 32, N)).sum(0))
    torch.backends.cuda.matmul.allow_tf32 = False
    r["fp32 mm(dz.t(), x)"] = t(lambda: torch.mm(dz.t(), x))
    torch.backends.cuda.matmul.allow_tf32 = True
    # forward / dX shapes for reference
    w = torch.randn(M, N, device=dev)
    r["fwd tf32 mm(x, w.t())"] = t(lambda: torch.mm(x, w.t()))
    r["dx tf32 mm(dz, w)"] = t(lambda: torch.mm(dz, w))
    wb = w.bfloat16()
    r["fwd bf16 pre-cast"] = t(lambda: torch.mm(xb, wb.t()))
    out[f"dZ {B}x{M}, X {B}x{N}"] = {k: round(v, 1) for k, v in r.items()}
print(json.dumps(out, indent=1))
