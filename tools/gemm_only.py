"""One tcgen05 GEMM shape for ncu captures: python tools/gemm_only.py M N K [iters]   (C = A[M,K] * B[N,K]^T)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sug_b200 import ops

M, N, K = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (65536, 512, 128)
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 10
dev = torch.device("cuda:0")
a = torch.randn(M, K, device=dev)
b = torch.randn(N, K, device=dev)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
for _ in range(2):
    ops.gemm_tc(a, b)
torch.cuda.synchronize()
ts = []
for _ in range(iters):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c = ops.gemm_tc(a, b); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
t = sorted(ts)[len(ts) // 2]
by = 4.0 * (M * K + N * K + M * N)
print(f"gemm_tc {M}x{N}x{K}: {t:.1f} us, {by / t / 1e3:.0f} GB/s, {2.0 * M * N * K / t / 1e6:.1f} TFLOP/s", flush=True)
