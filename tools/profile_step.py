"""Dev: torch.profiler breakdown of one SUG step (GPU time per kernel/op, CPU launch overhead)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from sug_b200 import Model, model_utils, step, synth

dev = torch.device("cuda:0")
torch.manual_seed(666)
model = Model.Net_MDA("DGCNN").to(dev).train()
opts = step.make_optimizers(model)
crit = model_utils.focal_loss(num_classes=10, gamma=0.0, alpha=[0.1] * 10)
B = 64
d, l = synth.synth_clouds(B, 1024, 0); dt, lt = synth.synth_clouds(B, 1024, 1)
d, l, dt, lt = (t.to(dev) for t in (d, l, dt, lt))
for _ in range(4):
    step.train_step(model, opts, d, l, dt, lt, crit)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step.train_step(model, opts, d, l, dt, lt, crit)
torch.cuda.synchronize()
print(f"wall per step: {(time.perf_counter()-t0)/5*1e3:.2f} ms")
# CPU-side time only (no sync inside)
t0 = time.perf_counter()
out = step.sug_losses(model, d, l, dt, lt, crit)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
out["loss"].backward()
t3 = time.perf_counter(); torch.cuda.synchronize(); t4 = time.perf_counter()
for o in opts: o.step()
for o in opts: o.zero_grad()
t5 = time.perf_counter(); torch.cuda.synchronize(); t6 = time.perf_counter()
print(f"forward: cpu {1e3*(t1-t0):.2f} ms, +gpu drain {1e3*(t2-t1):.2f}; backward: cpu {1e3*(t3-t2):.2f}, drain {1e3*(t4-t3):.2f}; optim: cpu {1e3*(t5-t4):.2f}, drain {1e3*(t6-t5):.2f}")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step.train_step(model, opts, d, l, dt, lt, crit)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
# kernel-only view: library kernels vs PyTorch glue
from collections import defaultdict
agg = defaultdict(lambda: [0.0, 0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        agg[ev.name][0] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
        agg[ev.name][1] += 1
ours = sum(v[0] for k, v in agg.items() if "sug::" in k)
rest = {k: v for k, v in agg.items() if "sug::" not in k}
print(f"KERNELS: library {ours/1e3:.3f} ms, other {sum(v[0] for v in rest.values())/1e3:.3f} ms")
for k, v in sorted(rest.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"  {v[0]:9.1f} us  x{v[1]:4d}  {k[:150]}")
