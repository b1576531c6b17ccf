"""Dev diagnostic: layer-by-layer comparison of the GPU DGCNN trunk with the CPU oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import sug_oracle as O
from sug_b200 import Model, ops, point_utils

dev = "cuda:0"
def rel(a, b):
    a = a.detach().double().cpu(); b = b.detach().double()
    return float((a - b).norm() / (b.norm() + 1e-30))

for seed in (0, 1):
    x, _ = O.synth_clouds(12, 1024, seed)
    sd = O.synth_state("Net_MDA:DGCNN")
    net = Model.Net_MDA("DGCNN"); net.load_state_dict(sd); net = net.to(dev).train()
    g = net.g
    xl = x.squeeze(-1)
    # oracle, step by step
    k = 20
    i1 = O.knn(xl, k); o1 = O.edgeconv(xl, sd, "g.conv1", True, idx=i1)
    i2 = O.knn(o1, k); o2 = O.edgeconv(o1, sd, "g.conv2", True, idx=i2)
    torch.manual_seed(3)
    start = torch.randint(0, 1024, (12,))
    oa, onode, ooff = O.adapt_layer_off(o2, xl, sd, "g.node_fea_adapt", True, fps_start=start)
    o2b = torch.nn.functional.conv1d(oa, sd["g.conv1d.weight"], sd["g.conv1d.bias"])
    i3 = O.knn(o2b, k); o3 = O.edgeconv(o2b, sd, "g.conv3", True, idx=i3)
    i4 = O.knn(o3, k); o4 = O.edgeconv(o3, sd, "g.conv4", True, idx=i4)
    # ours
    xg = xl.to(dev)
    j1 = ops.knn_cm(xg, k); g1 = g.conv1.edgeconv(xg.transpose(1, 2).contiguous(), j1)
    j2 = ops.knn_pm(g1, k); g2 = g.conv2.edgeconv(g1, j2)
    torch.manual_seed(3)
    ga, gnode, goff = g.node_fea_adapt(g2.transpose(1, 2).unsqueeze(3), xg)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        g2b = g.conv1d(ga.squeeze(-1)).transpose(1, 2).contiguous()
    j3 = ops.knn_pm(g2b, k); g3 = g.conv3.edgeconv(g2b, j3)
    j4 = ops.knn_pm(g3, k); g4 = g.conv4.edgeconv(g3, j4)
    def setdiff(a, b):
        return int((a.cpu().long().sort(-1)[0] != b.sort(-1)[0]).any(-1).sum())
    print(f"seed {seed}: knn rows differing L1..4:", setdiff(j1, i1), setdiff(j2, i2), setdiff(j3, i3), setdiff(j4, i4))
    print("  x1", rel(g1.transpose(1, 2), o1), "x2", rel(g2.transpose(1, 2), o2), "adapt", rel(ga.squeeze(-1), oa),
          "node", rel(gnode, onode), "off", rel(goff, ooff), "x2b", rel(g2b.transpose(1, 2), o2b),
          "x3", rel(g3.transpose(1, 2), o3), "x4", rel(g4.transpose(1, 2), o4))
    # teacher-forced deeper layers: feed the oracle's activations / indices to our kernels
    t3 = g.conv3.edgeconv(o2b.to(dev).transpose(1, 2).contiguous(), i3.to(dev).int())
    print("  teacher-forced x3", rel(t3.transpose(1, 2), o3))
    # adapt-layer indices on the oracle's x2
    fo = O.farthest_point_sample(xl, 64, start); fg = ops.fps(xg, 64, start).cpu().long()
    print("  fps equal:", bool((fo == fg).all()))
    floc = O.index_points(xl, fo)
    bo = O.query_ball_point(0.3, 64, xl, floc); bg = ops.ball_query(xg, floc.to(dev), 0.3, 64).cpu().long()
    print("  ball query differing entries:", int((bo != bg).sum()))
