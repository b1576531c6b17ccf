"""Dev check of the tcgen05 GEMM against fp64 (run under gpurun with a timeout)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sug_b200 import ops

def rel(a, b):
    return float((a.double() - b).norm() / b.norm())

torch.manual_seed(0)
dev = "cuda:0"
cases = [("NT", 256, 128, 64), ("NT", 128, 64, 32), ("NT", 1000, 200, 100), ("NT", 65536, 128, 64), ("NT", 65536, 512, 128),
         ("NT", 65536, 512, 512), ("NT", 4096, 64, 128)]
for kind, M, N, K in cases:
    a = torch.randn(M, K, device=dev); b = torch.randn(N, K, device=dev)
    bias = torch.randn(N, device=dev) if N % 2 == 0 and M < 5000 else None
    ref = a.double() @ b.double().t() + (bias.double() if bias is not None else 0)
    c = ops.gemm_tc(a, b, bias); torch.cuda.synchronize()
    e = rel(c, ref)
    e_simt = rel(ops.gemm(a, b, bias), ref)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5): ops.gemm_tc(a, b, bias)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 5
    print(f"{kind} M={M} N={N} K={K}: rel err tc {e:.2e} simt {e_simt:.2e}  {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
# strided A (slice of a wider buffer)
buf = torch.randn(4096, 512, device=dev); a = buf[:, 128:256]; b = torch.randn(256, 128, device=dev)
print("strided A:", rel(ops.gemm_tc(a, b), a.double() @ b.double().t()), flush=True)
# MN-major operands: dW = dY^T X
for (P, Co, C) in [(4096, 128, 64), (65536, 512, 128), (20000, 96, 40), (65536, 512, 512)]:
    dy = torch.randn(P, Co, device=dev); x = torch.randn(P, C, device=dev)
    ref = dy.double().t() @ x.double()
    c = ops.gemm_tc(dy.t(), x.t()); torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5): ops.gemm_tc(dy.t(), x.t())
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 5
    print(f"TN P={P} Co={Co} C={C}: rel err {rel(c, ref):.2e}  {ms*1e3:.1f} us {2*P*Co*C/ms/1e9:.1f} TFLOP/s", flush=True)
# NN: dX = dY W  (B MN-major)
dy = torch.randn(8192, 256, device=dev); w = torch.randn(256, 128, device=dev)
print("NN:", rel(ops.gemm_tc(dy, w.t()), dy.double() @ w.double()), flush=True)
print("DONE")
