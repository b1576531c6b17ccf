"""One EdgeConv block forward+backward for ncu captures: python tools/edge_only.py [C] [Cout] [iters]."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sug_b200 import ops, model_utils, synth

C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Co = int(sys.argv[2]) if len(sys.argv) > 2 else 128
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
B, N, k = 64, 1024, 20
dev = torch.device("cuda:0")
xyz = synth.synth_clouds(B, N, 1)[0].squeeze(-1).to(dev)  # [B,3,N]
idx = ops.knn_cm(xyz, k)                                   # a realistic (geometric) neighbour graph
x = torch.randn(B, N, C, generator=torch.Generator().manual_seed(1)).to(dev).requires_grad_(True)
blk = model_utils.conv_2d(2 * C, Co, 1, activation="leakyrelu", bias=False).to(dev).train()
g = torch.randn(B, N, Co, device=dev)
for _ in range(2):
    blk.edgeconv(x, idx).backward(g)
torch.cuda.synchronize()
ts = []
for _ in range(iters):
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record(); o = blk.edgeconv(x, idx); b.record(); o.backward(g); c.record(); torch.cuda.synchronize()
    ts.append((a.elapsed_time(b) * 1e3, b.elapsed_time(c) * 1e3))
ts.sort()
print(f"edgeconv C={C}->{Co}: fwd {ts[len(ts)//2][0]:.1f} us, bwd {ts[len(ts)//2][1]:.1f} us", flush=True)
