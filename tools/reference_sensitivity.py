"""How chaotic is the reference step itself?  Runs the oracle twice -- once with every weight perturbed by a
relative `eps` -- and prints the largest relative gradient change per parameter.
    python tools/reference_sensitivity.py [dgcnn|pointnet] [eps]        (CPU; about a minute)
Measured: DGCNN, eps 3e-7 -> 3.4e-3;  PointNet, eps 3e-7 -> 1.0e-2, eps 2e-6 -> 3.95e-2."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sug_oracle as O

model = sys.argv[1].lower() if len(sys.argv) > 1 else "dgcnn"
eps = float(sys.argv[2]) if len(sys.argv) > 2 else 3e-7
name, spec, seed, ds = ("Pointnet", "Net_MDA:Pointnet", 668, (4, 5)) if model == "pointnet" else ("DGCNN", "Net_MDA:DGCNN", 666, (0, 1))
torch.set_num_threads(os.cpu_count())
Bs = 12
data, label = O.synth_clouds(Bs, 1024, ds[0])
data_t, label_t = O.synth_clouds(Bs, 1024, ds[1])


def run(perturb):
    sd = O.clone_state(O.synth_state(spec, seed), requires_grad=True)
    if perturb:
        g = torch.Generator().manual_seed(5)
        for v in sd.values():
            if v.requires_grad:
                v.data.mul_(1 + perturb * torch.randn(v.shape, generator=g))
    torch.manual_seed(101)
    r = O.sug_losses(sd, data, label, data_t, label_t, O.FocalLoss([0.1] * 10, 0.0), model_name=name, drop_p=0.0,
                     mmd_dtype=torch.float64)
    r["loss"].backward()
    return sd, r


a, ra = run(0)
b, rb = run(eps)
print("loss", float(ra["loss"]), float(rb["loss"]))
gmax = max(float(v.grad.norm()) for v in a.values() if v.grad is not None)
rows = []
for k, v in a.items():
    if v.grad is None:
        continue
    d = float((v.grad.double() - b[k].grad.double()).norm())
    rows.append((d / max(float(v.grad.norm()), 1e-4 * gmax), k))
rows.sort(reverse=True)
print(rows[:6])
print("median", rows[len(rows) // 2][0], "of", len(rows), "parameters")
