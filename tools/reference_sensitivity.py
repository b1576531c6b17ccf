import sys, torch, numpy as np
sys.path.insert(0,'/root/repo')
from oracle import sug_oracle as O
torch.set_num_threads(8)
Bs=12
data,label=O.synth_clouds(Bs,1024,0); data_t,label_t=O.synth_clouds(Bs,1024,1)
def run(perturb, forced=None):
    sd=O.clone_state(O.synth_state("Net_MDA:DGCNN"),requires_grad=True)
    if perturb:
        g=torch.Generator().manual_seed(5)
        for k,v in sd.items():
            if v.requires_grad: v.data.mul_(1+perturb*torch.randn(v.shape,generator=g))
    torch.manual_seed(101)
    O.KNN_TRACE=[]
    r=O.sug_losses(sd,data,label,data_t,label_t,O.FocalLoss([0.1]*10,0.0),drop_p=0.0,mmd_dtype=torch.float64)
    tr=O.KNN_TRACE; O.KNN_TRACE=None
    r["loss"].backward()
    return sd,r,tr
a,ra,ta=run(0); b,rb,tb=run(3e-7)
print("loss", float(ra["loss"]), float(rb["loss"]))
print("knn rows differing:", [int((x.sort(-1)[0]!=y.sort(-1)[0]).any(-1).sum()) for x,y in zip(ta,tb)])
gmax=max(float(v.grad.norm()) for v in a.values() if v.grad is not None)
rows=[]
for k,v in a.items():
    if v.grad is None: continue
    d=float((v.grad.double()-b[k].grad.double()).norm())
    rows.append((d/max(float(v.grad.norm()),1e-4*gmax),k))
rows.sort(reverse=True); print(rows[:6])
