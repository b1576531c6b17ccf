"""One kNN launch shape for ncu source-level captures: python tools/knn_only.py [C] [k] [N] [B] [iters]."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sug_b200 import ops

C, k, N, B, iters = [int(a) if i < len(sys.argv) - 1 else d for i, (a, d) in enumerate(
    zip(sys.argv[1:] + [None] * 5, (64, 20, 1024, 64, 5)))]
C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
k = int(sys.argv[2]) if len(sys.argv) > 2 else 20
N = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
B = int(sys.argv[4]) if len(sys.argv) > 4 else 64
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 5
dev = torch.device("cuda:0")
x = torch.randn(B, N, C, generator=torch.Generator().manual_seed(1)).to(dev)
for _ in range(2):
    ops.knn_pm(x, k)
torch.cuda.synchronize()
ts = []
for _ in range(iters):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); idx = ops.knn_pm(x, k); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
print(f"knn C={C} k={k} N={N} B={B}: {sorted(ts)[len(ts)//2]:.1f} us (incl. prep)", flush=True)
