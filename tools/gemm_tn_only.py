"""Weight-gradient shaped GEMMs (dY^T X, both operands MN-major, K = B*N) on the tcgen05 path: time and error vs fp64.
    python tools/gemm_tn_only.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sug_b200 import ops
dev = torch.device("cuda:0")
for (P, Co, C) in [(65536, 512, 128), (65536, 256, 64), (65536, 128, 64)]:
    dy, x = torch.randn(P, Co, device=dev), torch.randn(P, C, device=dev)
    ref = dy.double().t() @ x.double()
    for _ in range(2): out = ops.gemm_tc(dy.t(), x.t())
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    ts = []
    for _ in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.gemm_tc(dy.t(), x.t()); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    print(f"TN {P}x{Co}x{C}: {sorted(ts)[4]:.1f} us, rel err {err:.2e}", flush=True)
