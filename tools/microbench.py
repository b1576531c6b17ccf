"""BASELINE.json configs[3]: kNN / EdgeConv micro-benchmark sweep (N, k, C) against the HBM and tensor
rooflines, and configs[4]: LiDAR-scale DGCNN inference (N = 16384).  Prints a markdown table.
Timing: CUDA events, 3 warm-up + 10 timed launches; every case touches more than the 126 MB L2 or is
preceded by an L2 flush (256 MB write)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sug_b200 import ops, model_utils, model_pointnet, synth

dev = torch.device("cuda:0")
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) \
    else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


print("| op | N | k | C->Cout | B | time us | alg GB/s (% HBM) | alg TFLOP/s (% bf16) |")
print("|---|---|---|---|---|---|---|---|")
QUICK = "--quick" in sys.argv  # step shapes only (B = 64, N = 1024, k = 20), no LiDAR-scale inference
for N in (() if "--lidar" in sys.argv else (1024,) if QUICK else (1024, 2048, 4096)):
    B = 65536 // N
    for k in ((20,) if QUICK else (20, 40)):
        for C, Co in ((3, 64), (64, 64), (64, 128), (128, 256)):
            g = torch.Generator(device="cpu").manual_seed(N + k + C)
            if C == 3:
                x = synth.synth_clouds(B, N, 1)[0].squeeze(-1).to(dev)  # [B,3,N]
                xpm = x.transpose(1, 2).contiguous()
                f = lambda: ops.knn_cm(x, k)
            else:
                xpm = torch.randn(B, N, C, generator=g).to(dev)
                f = lambda: ops.knn_pm(xpm, k)
            t = timeit(f)
            by, fl = 4.0 * B * N * (C + k), 2.0 * B * N * N * C
            print(f"| knn | {N} | {k} | {C} | {B} | {t*1e6:.1f} | {by/t/1e9:.0f} ({100*by/t/1e9/peaks['hbm_gbs']:.1f}%) | {fl/t/1e12:.1f} ({100*fl/t/1e12/peaks['bf16_tflops']:.2f}%) |")
            idx = f()
            blk = model_utils.conv_2d(2 * C, Co, 1, activation="leakyrelu", bias=False).to(dev).train()
            xin = xpm.clone().requires_grad_(C != 3)
            fwd = lambda: blk.edgeconv(xin, idx)
            t = timeit(fwd)
            by = B * N * (4.0 * C + 4.0 * k + 4.0 * Co + Co + 4.0 * Co)
            fl = 2.0 * B * N * 2 * C * Co
            print(f"| edgeconv fwd (train) | {N} | {k} | {C}->{Co} | {B} | {t*1e6:.1f} | {by/t/1e9:.0f} ({100*by/t/1e9/peaks['hbm_gbs']:.1f}%) | {fl/t/1e12:.1f} ({100*fl/t/1e12/peaks['bf16_tflops']:.2f}%) |")
            out = fwd()
            gout = torch.randn_like(out)

            def bwd():
                o = blk.edgeconv(xin, idx)
                o.backward(gout)
            t2 = timeit(bwd) - t
            by = B * N * (4.0 * (2 * Co + 2 * C) + 4.0 * k + Co)
            print(f"| edgeconv bwd | {N} | {k} | {C}->{Co} | {B} | {t2*1e6:.1f} | {by/t2/1e9:.0f} ({100*by/t2/1e9/peaks['hbm_gbs']:.1f}%) | {3*fl/t2/1e12:.1f} |")

if QUICK and "--lidar" not in sys.argv:
    sys.exit(0)
# configs[4]: LiDAR-scale inference, N = 16384, k = 20 (the reference would need four 1.07 GB N x N tensors per cloud)
net = model_pointnet.DGCNN().to(dev).eval()
for B in (1, 4):
    x = synth.synth_clouds(B, 16384, 7)[0].to(dev)
    with torch.no_grad():
        t = timeit(lambda: net(x), iters=5)
    print(f"| DGCNN inference (model_pointnet.DGCNN, eval) | 16384 | 20 | - | {B} | {t*1e6:.0f} | {B/t:.1f} clouds/s | peak mem {torch.cuda.max_memory_allocated()/2**30:.2f} GiB |")
    from sug_b200 import step as _step, _lib as _l
    fwd = _step.GraphedEval(net, x)
    t = timeit(lambda: fwd(x), iters=5)
    print(f"| same, one CUDA graph (step.GraphedEval) | 16384 | 20 | - | {B} | {t*1e6:.0f} | {B/t:.1f} clouds/s | |")
    if B == 1:
        _l.prof_reset(mask=0xFFFFFFFF)
        with torch.no_grad():
            net(x)
        pr = _l.prof_collect()
        print("| per-class us of one eager forward: " + ", ".join(f"{k} {v['ms']*1e3:.0f}" for k, v in pr.items() if v["launches"]) + " | | | | | | | |")
