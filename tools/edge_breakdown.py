"""Per-kernel-class time of one EdgeConv block forward + backward (library event timers):
python tools/edge_breakdown.py [C] [Cout] [feature-space graph: 0|1]."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sug_b200 import ops, model_utils, synth, _lib

C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Co = int(sys.argv[2]) if len(sys.argv) > 2 else 128
feat = int(sys.argv[3]) if len(sys.argv) > 3 else 1
B, N, k = 64, 1024, 20
dev = torch.device("cuda:0")
x = torch.randn(B, N, C, generator=torch.Generator().manual_seed(1)).to(dev).requires_grad_(C != 3)
if feat and C != 3:
    idx = ops.knn_pm(x.detach(), k)  # feature-space graph (hubs)
else:
    xyz = synth.synth_clouds(B, N, 1)[0].squeeze(-1).to(dev)
    idx = ops.knn_cm(xyz, k)
    if C == 3:
        x = xyz.transpose(1, 2).contiguous()
deg = torch.zeros(B, N, device=dev).scatter_add_(1, idx.long().reshape(B, -1), torch.ones(B, N * k, device=dev))
print(f"in-degree: max {int(deg.max())}, zero-degree rows {float((deg == 0).float().mean()) * 100:.1f} %")
blk = model_utils.conv_2d(2 * C, Co, 1, activation="leakyrelu", bias=False).to(dev).train()
g = torch.randn(B, N, Co, device=dev)
for _ in range(3):
    o = blk.edgeconv(x, idx)
    if C != 3:
        o.backward(g)
    else:
        torch.autograd.backward(o, g)
for phase in ("fwd", "bwd"):
    tot = {}
    for _ in range(5):
        if phase == "fwd":
            _lib.prof_reset(mask=0xFFFFFFFF)
            o = blk.edgeconv(x, idx)
            pr = _lib.prof_collect()
        else:
            o = blk.edgeconv(x, idx)
            _lib.prof_reset(mask=0xFFFFFFFF)
            torch.autograd.backward(o, g)
            pr = _lib.prof_collect()
        for kname, v in pr.items():
            if v["launches"]:
                tot.setdefault(kname, []).append((v["ms"] * 1e3, v["launches"]))
    print(phase, {kname: f"{sorted(t for t, _ in v)[len(v)//2]:.1f} us / {v[0][1]} launches" for kname, v in tot.items()})
