import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sug_b200 import Model, model_utils, step, synth
dev = torch.device("cuda:0")
torch.manual_seed(666)
model = Model.Net_MDA("DGCNN").to(dev).train()
opts = step.make_optimizers(model, capturable=True)
crit = model_utils.focal_loss(num_classes=10, gamma=0.0, alpha=[0.1] * 10)
B = 16
d, l = synth.synth_clouds(B, 1024, 0); dt, lt = synth.synth_clouds(B, 1024, 1)
batch = tuple(t.to(dev) for t in (d, l, dt, lt))
g = step.GraphedTrainStep(model, opts, crit, B, 1024, dev)
try:
    g.warm(*batch)
    print("warm ok")
    g.capture()
    print("capture ok")
    for i in range(3):
        out = g(*batch)
    torch.cuda.synchronize()
    print("replay ok", float(out["loss"]))
except Exception:
    traceback.print_exc()
