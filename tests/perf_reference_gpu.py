"""Not a test (pytest does not collect it): times the REFERENCE ALGORITHM on the GPU, i.e. the CPU
oracle's functional PyTorch code moved to cuda:0 — the same cuBLAS / cuDNN / ATen kernels, Python FPS
loop, boolean-mask syncs and `.cpu()` round trip that the reference's own modules would run
(SURVEY.md §2b).  This is the denominator of north_star's "x3 over the reference's PyTorch-CUDA path";
the reference tree itself is not available on the GPU box.

    python tests/perf_reference_gpu.py [B]
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sug_oracle as O  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    dev = torch.device("cuda:0")
    sd = {k: v.to(dev) for k, v in O.clone_state(O.synth_state("Net_MDA:DGCNN")).items()}
    for v in sd.values():
        if v.is_floating_point() and v.dim() > 0:
            v.requires_grad_(True)
    for k in sd:
        if "running_" in k:
            sd[k].requires_grad_(False)
    params = [v for v in sd.values() if v.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-4, weight_decay=5e-4)
    data, label = O.synth_clouds(B, 1024, 0)
    data_t, label_t = O.synth_clouds(B, 1024, 1)
    data, label, data_t, label_t = (t.to(dev) for t in (data, label, data_t, label_t))
    crit = O.FocalLoss([0.1] * 10, 0.0)
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32  # PyTorch default is True (what the reference gets)
        ts = []
        for it in range(8):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = O.sug_losses(sd, data, label, data_t, label_t, crit)
            out["loss"].backward()
            opt.step()
            opt.zero_grad()
            float(out["loss"].detach())
            torch.cuda.synchronize()
            if it >= 3:
                ts.append(time.perf_counter() - t0)
        ms = 1e3 * sum(ts) / len(ts)
        print(f"reference algorithm on GPU (torch ops, cudnn tf32={tf32}), B={B}+{B}: {ms:.1f} ms/step = "
              f"{2 * B / ms * 1e3:.0f} clouds/s; peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)


if __name__ == "__main__":
    main()
