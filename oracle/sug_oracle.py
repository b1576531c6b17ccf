"""CPU oracle for the SUG point-cloud hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain functional PyTorch on the CPU, the algorithm of the
reference (SiyuanHuang95/SUG) for the one hot path this repository accelerates:
DGCNN kNN / EdgeConv, the PointNet shared MLP + global max-pool, the self-adaptive
node layer that sits inside both encoders, the two-head classifier wrapper and the
multi-kernel Gaussian MMD with its SDA sample weights.  Every function cites the
reference file:line it follows.  Nothing under ``sug_b200/`` may import this file;
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker or the CPU arm.

Parity pinning: the reference ships no tests and no golden vectors, so
this oracle is pinned against outputs of the reference itself, produced in the build
container by ``tests/golden/make_golden.py`` (which imports /root/reference) and
committed as ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks every
function below against those fixtures.

Third-party arithmetic that is not under /root/reference:
  * ``chamfer_distance`` (github.com/otaheri/chamfer_distance, no commit pinned by the
    reference, README.md:58-62).  Published algorithm: for every point of cloud 1 the
    squared L2 distance to its nearest neighbour in cloud 2, and vice versa.  Restated
    in ``chamfer`` below; parity is anchored on the reference's call sites
    mmd.py:126-128,169-175.
  * ``scipy.special.kl_div`` (dataset_splitter.py:24,244-245): x*log(x/y) - x + y.

All tensors are float32 unless stated.  State is a flat ``dict`` that uses the
reference's ``state_dict`` key names (e.g. ``g.conv1.conv.0.weight``).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

SIGMA_LIST = [0.01, 0.1, 1, 10, 100]  # mmd.py:23
MIN_VAR_EST = 1e-8  # mmd.py:22
K_NEIGH = 20  # Model.py:52
BN_EPS = 1e-5
BN_MOM = 0.1

State = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------
# kNN graph + edge features                                   model_utils.py:178-210
# --------------------------------------------------------------------------------------
def pairwise_neg_sqdist(x: torch.Tensor) -> torch.Tensor:
    """x [B,C,N] -> D [B,N,N] = -|x_i - x_j|^2 in the reference's operation order
    (model_utils.py:179-181): inner=-2 x^T x; xx=sum x^2; D = -xx - inner - xx^T."""
    inner = -2 * torch.matmul(x.transpose(2, 1), x)
    xx = torch.sum(x ** 2, dim=1, keepdim=True)
    return -xx - inner - xx.transpose(2, 1)


def knn_row_hash(idx):
    """16-bit hash of every row's neighbour SET (order-independent): idx [B,N,k] integer -> int32 [B,N] in
    [0, 65536).  Used to compare neighbour graphs at sizes where storing the lists is impractical."""
    s = idx.long().sort(dim=-1)[0]
    w = torch.arange(1, 2 * s.shape[-1], 2, device=s.device, dtype=torch.long)
    h = (s * w).sum(-1) * 40503 + (s * s).sum(-1) * 2654435761
    return ((h ^ (h >> 16)) & 0xFFFF).to(torch.int32)


KNN_TRACE = None  # set to a list to record every neighbour list knn() returns (teacher forcing in tests)


def knn(x: torch.Tensor, k: int) -> torch.Tensor:
    """model_utils.py:178-185.  int64 [B,N,k], nearest first, self included."""
    idx = pairwise_neg_sqdist(x).topk(k=k, dim=-1)[1]
    if KNN_TRACE is not None:
        KNN_TRACE.append(idx)
    return idx


def get_graph_feature(x: torch.Tensor, k: int = 20, idx: Optional[torch.Tensor] = None) -> torch.Tensor:
    """model_utils.py:188-210.  x [B,C,N] or [B,C,N,1] -> [B,2C,N,k] with channel
    order [x_j - x_i ; x_i] (line 208)."""
    B, N = x.size(0), x.size(2)
    x = x.reshape(B, -1, N)
    if idx is None:
        idx = knn(x, k)
    k = idx.shape[-1]
    C = x.shape[1]
    xt = x.transpose(2, 1).contiguous()  # [B,N,C]
    flat = (idx + torch.arange(B, device=x.device).view(-1, 1, 1) * N).reshape(-1)
    nb = xt.reshape(B * N, C)[flat].view(B, N, k, C)
    ctr = xt.view(B, N, 1, C).expand(B, N, k, C)
    return torch.cat((nb - ctr, ctr), dim=3).permute(0, 3, 1, 2)


# --------------------------------------------------------------------------------------
# conv_2d / fc_layer blocks                                   model_utils.py:8-57
# --------------------------------------------------------------------------------------
def _bn(x, sd: State, p: str, training: bool):
    """nn.BatchNorm{1,2}d with default eps/momentum; updates running stats in ``sd``."""
    if training:
        sd[p + ".num_batches_tracked"] += 1
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                        sd[p + ".bias"], training, BN_MOM, BN_EPS)


def conv_2d(x, sd: State, p: str, training: bool, act: str = "relu"):
    """model_utils.py:8-32: Conv2d(1x1) -> BatchNorm2d -> {ReLU | Tanh | LeakyReLU(0.01)}."""
    y = F.conv2d(x, sd[p + ".conv.0.weight"], sd.get(p + ".conv.0.bias"))
    y = _bn(y, sd, p + ".conv.1", training)
    if act == "relu":
        return F.relu(y)
    if act == "tanh":
        return torch.tanh(y)
    if act == "leakyrelu":
        return F.leaky_relu(y, 0.01)  # nn.LeakyReLU() default slope, model_utils.py:27
    raise ValueError(act)


def fc_layer(x, sd: State, p: str, act: str):
    """model_utils.py:35-57 with bn=True: Linear -> LayerNorm -> ReLU | LeakyReLU(0.2)."""
    y = F.linear(x, sd[p + ".fc.0.weight"], sd.get(p + ".fc.0.bias"))
    y = F.layer_norm(y, (y.shape[-1],), sd[p + ".fc.1.weight"], sd[p + ".fc.1.bias"])
    return F.relu(y) if act == "relu" else F.leaky_relu(y, 0.2)


def edgeconv(x, sd: State, p: str, training: bool, k: int = K_NEIGH, idx=None):
    """get_graph_feature -> conv_2d('leakyrelu', bias=False) -> max over k
    (Model.py:88-94).  x [B,C,N] -> [B,Cout,N]."""
    e = get_graph_feature(x, k=k, idx=idx)
    return conv_2d(e, sd, p, training, "leakyrelu").max(dim=-1)[0]


# --------------------------------------------------------------------------------------
# point_utils.py: FPS, ball query, 3-NN interpolation
# --------------------------------------------------------------------------------------
def farthest_point_sample(xyz: torch.Tensor, npoint: int, start: Optional[torch.Tensor] = None):
    """point_utils.py:5-26.  xyz [B,3,N] -> int64 [B,npoint].  The random start index is
    drawn with torch.randint on the CPU generator exactly like line 17 unless given."""
    B, _, N = xyz.shape
    dev = xyz.device
    cent = torch.zeros(B, npoint, dtype=torch.long, device=dev)
    dist = torch.ones(B, N, device=dev) * 1e10
    far = (torch.randint(0, N, (B,), dtype=torch.long) if start is None else start.clone()).to(dev)
    ar = torch.arange(B, device=dev)
    for i in range(npoint):
        cent[:, i] = far
        c = xyz[ar, :, far].view(B, 3, 1)
        d = torch.sum((xyz - c) ** 2, 1)
        dist = torch.minimum(dist, d)  # == masked overwrite of lines 23-24
        far = torch.max(dist, -1)[1]
    return cent


def index_points(points: torch.Tensor, idx: torch.Tensor):
    """point_utils.py:60-83.  points [B,C,N], idx [B,S] or [B,S,K] -> [B,C,S] / [B,C,S,K]."""
    B, C, _ = points.shape
    pt = points.permute(0, 2, 1)
    out = pt[torch.arange(B, device=points.device).view(B, *([1] * (idx.dim() - 1))), idx]
    return out.permute(0, 2, 1) if idx.dim() == 2 else out.permute(0, 3, 1, 2)


def square_distance(src, dst):
    """point_utils.py:112-131.  [B,C,N],[B,C,M] -> [B,N,M] = -2 s.d + |s|^2 + |d|^2."""
    B, _, N = src.shape
    M = dst.shape[2]
    d = -2 * torch.matmul(src.permute(0, 2, 1), dst)
    d = d + torch.sum(src ** 2, 1).view(B, N, 1)
    d = d + torch.sum(dst ** 2, 1).view(B, 1, M)
    return d


def query_ball_point(radius, nsample, xyz, new_xyz):
    """point_utils.py:86-109.  With a radius: the ``nsample`` lowest-index points whose
    squared distance is <= r^2, padded with the first hit.  radius None: the nsample
    nearest points in ascending-distance order (full sort)."""
    B, _, N = xyz.shape
    S = new_xyz.shape[2]
    sq = square_distance(new_xyz, xyz)
    if radius is None:
        return torch.sort(sq, dim=-1)[1][:, :, :nsample]
    g = torch.arange(N, device=xyz.device).view(1, 1, N).repeat(B, S, 1)
    g[sq > radius ** 2] = N
    g = g.sort(dim=-1)[0][:, :, :nsample]
    first = g[:, :, :1].expand(-1, -1, nsample)
    return torch.where(g == N, first, g)


def upsample_inter(xyz1, xyz2, points1, points2, k=3):
    """point_utils.py:134-165: inverse-squared-distance interpolation of points2 (at xyz2)
    onto xyz1 using the k nearest; concatenated after points1."""
    B, _, N = xyz1.shape
    d, idx = square_distance(xyz1, xyz2).sort(dim=-1)
    d, idx = d[:, :, :k], idx[:, :, :k]
    d = torch.where(d < 1e-10, torch.full_like(d, 1e-10), d)
    w = 1.0 / d
    w = w / torch.sum(w, dim=-1).view(B, N, 1)
    interp = torch.sum(index_points(points2, idx) * w.view(B, 1, N, k), dim=3)
    return torch.cat([points1, interp], dim=1)


def adapt_layer_off(input_fea, input_loc, sd: State, p: str, training: bool, num_node=64,
                    fps_start=None):
    """model_utils.py:92-128.  input_fea [B,64,N], input_loc [B,3,N] ->
    (out [B,128,N], node_fea [B,64,num_node,1], node_offset [B,3,num_node])."""
    fidx = farthest_point_sample(input_loc, num_node, fps_start)
    f_loc = index_points(input_loc, fidx)
    f_fea = index_points(input_fea, fidx)
    gidx = query_ball_point(0.3, 64, input_loc, f_loc)
    g_fea = index_points(input_fea, gidx) - f_fea.unsqueeze(3)
    sem = torch.tanh(F.conv2d(g_fea, sd[p + ".pred_offset.0.weight"]))  # lines 98-100,113
    g_loc = index_points(input_loc, gidx) - f_loc.unsqueeze(3)
    node_off = (sem * g_loc).mean(dim=-1)
    node_loc = f_loc + node_off
    gidx2 = query_ball_point(None, 64, input_loc, node_loc)
    res = conv_2d(input_fea.unsqueeze(3), sd, p + ".residual", training, "relu").squeeze(3)
    node_fea = index_points(res, gidx2).max(dim=-1, keepdim=True)[0]
    out = upsample_inter(input_loc, node_loc, input_fea, node_fea.squeeze(3), k=3)
    return out, node_fea, node_off


# --------------------------------------------------------------------------------------
# Encoders                                                    Model.py:54-121, 235-283
# --------------------------------------------------------------------------------------
def dgcnn_trunk(x, sd: State, p: str, training: bool, adapt: bool, fps_start=None, knn_idx=None):
    """Model.py:73-116 (adapt=True) / model_pointnet.py:114-155 (adapt=False).
    x [B,3,N,1] -> (feat [B,1024], node_fea or None).  ``knn_idx`` optionally teacher-forces
    the four neighbour lists (get_graph_feature accepts idx, model_utils.py:188-193)."""
    B, N = x.size(0), x.size(2)
    x_loc = x.reshape(B, 3, N)
    ti = knn_idx if knn_idx is not None else [None] * 4
    x1 = edgeconv(x_loc, sd, p + "conv1", training, idx=ti[0])
    x2 = edgeconv(x1, sd, p + "conv2", training, idx=ti[1])
    node_fea = None
    if adapt:
        x_, node_fea, _ = adapt_layer_off(x2, x_loc, sd, p + "node_fea_adapt", training, fps_start=fps_start)
        x2 = F.conv1d(x_, sd[p + "conv1d.weight"], sd[p + "conv1d.bias"])
    x3 = edgeconv(x2, sd, p + "conv3", training, idx=ti[2])
    x4 = edgeconv(x3, sd, p + "conv4", training, idx=ti[3])
    xc = torch.cat((x1, x2, x3, x4), dim=1)
    x5 = F.conv1d(xc, sd[p + "conv5.weight"])
    x5 = F.leaky_relu(_bn(x5, sd, p + "bn5", training), 0.2)
    feat = torch.cat((x5.max(dim=2)[0], x5.mean(dim=2)), 1)
    return feat, node_fea


def transform_net(x, sd: State, p: str, training: bool, K: int):
    """model_utils.py:60-89 (DGCNN_Flag=False).  x [B,Cin,N,1] -> [B,K,K]."""
    y = conv_2d(x, sd, p + ".conv2d1", training)
    y = conv_2d(y, sd, p + ".conv2d2", training)
    y = conv_2d(y, sd, p + ".conv2d3", training)
    y = torch.max(y, dim=2)[0].view(x.size(0), -1)
    y = fc_layer(y, sd, p + ".fc1", "leakyrelu")
    y = fc_layer(y, sd, p + ".fc2", "leakyrelu")
    y = F.linear(y, sd[p + ".fc3.weight"], sd[p + ".fc3.bias"])
    return (y + torch.eye(K, device=y.device).view(1, K * K)).view(-1, K, K)


def pointnet_g(x, sd: State, p: str, training: bool, fps_start=None):
    """Model.py:248-283.  x [B,3,N,1] -> (feat [B,1024], node_fea [B,64,64,1], node_off)."""
    B, N = x.size(0), x.size(2)
    x_loc = x.reshape(B, 3, N)
    t1 = transform_net(x, sd, p + "trans_net1", training, 3)
    y = torch.bmm(x_loc.transpose(2, 1), t1).transpose(2, 1).unsqueeze(3)
    y = conv_2d(y, sd, p + "conv1", training)
    y = conv_2d(y, sd, p + "conv2", training)
    t2 = transform_net(y, sd, p + "trans_net2", training, 64)
    y = torch.bmm(y.squeeze(3).transpose(2, 1), t2).transpose(2, 1)
    y, node_fea, node_off = adapt_layer_off(y, x_loc, sd, p + "conv3", training, fps_start=fps_start)
    y = conv_2d(y.unsqueeze(3), sd, p + "conv4", training)
    y = conv_2d(y, sd, p + "conv5", training)
    y = torch.max(y, dim=2)[0].squeeze(-1)
    y = _bn(y, sd, p + "bn1", training)
    return y, node_fea, node_off


# --------------------------------------------------------------------------------------
# Heads + Net_MDA                                             Model.py:16-34, 412-520
# --------------------------------------------------------------------------------------
def pointnet_c(x, sd: State, p: str, training: bool, dgcnn: bool, drop_p: float = 0.4):
    """Model.py:412-449.  Returns (logits [B,10], mid_feature [B,256])."""
    act = "leakyrelu" if dgcnn else "relu"
    y = fc_layer(x, sd, p + ".mlp1", act)
    y = F.dropout(y, drop_p, training)  # Dropout2d on a 2-D input acts element-wise
    mid = fc_layer(y, sd, p + ".mlp2", act)
    y = F.dropout(mid, drop_p, training)
    return F.linear(y, sd[p + ".mlp3.weight"], sd[p + ".mlp3.bias"]), mid


def calayer(x, sd: State, p: str, training: bool):
    """Model.py:16-34.  x [B,4096] -> [B,4096]."""
    v = x.view(x.size(0), -1, 1, 1)
    y = F.relu(F.conv2d(v, sd[p + ".conv_du.0.weight"], sd[p + ".conv_du.0.bias"]))
    y = torch.sigmoid(F.conv2d(y, sd[p + ".conv_du.2.weight"], sd[p + ".conv_du.2.bias"]))
    y = (v * y + v).view(x.size(0), -1)
    return _bn(y, sd, p + ".bn", training)


def net_mda(x, sd: State, training: bool, model_name: str = "DGCNN", node_adaptation_s=False,
            node_adaptation_t=False, semantic_adaption=False, fps_start=None, drop_p: float = 0.4,
            knn_idx=None):
    """Model.py:485-520 (node_vis / mid_feat / adaptation branches omitted: GradReverse is a
    no-op in the reference, Model.py:37-50)."""
    if model_name == "DGCNN":
        feat, node = dgcnn_trunk(x, sd, "g.", training, adapt=True, fps_start=fps_start, knn_idx=knn_idx)
    else:
        feat, node, _ = pointnet_g(x, sd, "g.", training, fps_start=fps_start)
    B = node.size(0)
    if node_adaptation_s:
        return calayer(node.contiguous().view(B, -1), sd, "attention_s", training)
    if node_adaptation_t:
        return calayer(node.contiguous().view(B, -1), sd, "attention_t", training)
    dg = model_name == "DGCNN"
    y1, s1 = pointnet_c(feat, sd, "c1", training, dg, drop_p)
    y2, s2 = pointnet_c(feat, sd, "c2", training, dg, drop_p)
    return (y1, y2, s1, s2) if semantic_adaption else (y1, y2)


def dgcnn_cls(x, sd: State, training: bool, drop_p: float = 0.4):
    """model_pointnet.py:93-161 — DGCNN without the adapt layer + Pointnet_c classifier."""
    feat, _ = dgcnn_trunk(x, sd, "", training, adapt=False)
    return pointnet_c(feat, sd, "classifier", training, True, drop_p)[0]


# --------------------------------------------------------------------------------------
# MMD + SDA weights                                           mmd.py
# --------------------------------------------------------------------------------------
def one_hot(labels, num_class=10):
    """common_utils.py:161-164."""
    o = torch.zeros(labels.shape[0], num_class, device=labels.device)
    o[torch.arange(labels.shape[0], device=labels.device), labels] = 1
    return o


def mix_rbf_kernel(X, Y, sigma_list=SIGMA_LIST):
    """mmd.py:239-254.  Norms come from the Gram diagonal so the exponent is exactly 0 there."""
    m = X.size(0)
    Z = torch.cat((X, Y), 0)
    G = torch.mm(Z, Z.t())
    d = torch.diag(G).unsqueeze(1).expand_as(G)
    E = d - 2 * G + d.t()
    K = 0.0
    for s in sigma_list:
        K = K + torch.exp(-(1.0 / (2 * s ** 2)) * E)
    return K[:m, :m], K[:m, m:], K[m:, m:]


def mmd2(Kxx, Kxy, Kyy, biased=True, sample_weights=None):
    """mmd.py:274-312 with const_diagonal=False."""
    m = Kxx.size(0)
    dx, dy = torch.diag(Kxx), torch.diag(Kyy)
    kxx = (Kxx.sum(dim=1) - dx).sum()
    kyy = (Kyy.sum(dim=1) - dy).sum()
    kxy0 = Kxy.sum(dim=0)
    if sample_weights is not None:
        kxy0 = sample_weights.reshape(-1).to(kxy0.device) * kxy0  # mmd.py:295 (.to(device='cuda'))
    kxy = kxy0.sum()
    if biased:
        return (kxx + dx.sum()) / (m * m) + (kyy + dy.sum()) / (m * m) - 2.0 * kxy / (m * m)
    return kxx / (m * (m - 1)) + kyy / (m * (m - 1)) - 2.0 * kxy / (m * m)


def mix_rbf_mmd2(X, Y, sigma_list=SIGMA_LIST, biased=True, sample_weights=None):
    """mmd.py:257-260."""
    return mmd2(*mix_rbf_kernel(X, Y, sigma_list), biased=biased, sample_weights=sample_weights)


def chamfer(p1, p2):
    """ChamferDistance()(p1,p2)[:2] (third-party, see header).  p1 [B,N,3], p2 [B,M,3] ->
    (dist1 [B,N], dist2 [B,M]) squared nearest-neighbour distances, as explicit differences."""
    d = ((p1[:, :, None, :] - p2[:, None, :, :]) ** 2).sum(-1)
    return d.min(2)[0], d.min(1)[0]


def distance2weights_mean2one(dist):
    """mmd.py:198-201: scale = int(1/mean) (truncation), weights = dist*scale."""
    return dist * (1 / dist.mean()).type(torch.int)


def geometric_weights(pc_s, pc_t):
    """mmd.py:107-131 with weighting='mean2one'.  pc [B,3,N,1] -> [1,B]."""
    a = pc_s.transpose(1, 2).squeeze(-1) if pc_s.shape[1] == 3 else pc_s
    b = pc_t.transpose(1, 2).squeeze(-1) if pc_t.shape[1] == 3 else pc_t
    d1, d2 = chamfer(a, b)
    return distance2weights_mean2one(d1.mean(1) + d2.mean(1)).reshape(1, -1)


def _kl_div(x, y):
    """scipy.special.kl_div for positive arguments."""
    return x * torch.log(x / y) - x + y


def prob_weights_soft(pred_s, pred_t, label_s, label_t, label_weight):
    """mmd.py:134-153 with weighting='mean2one'.  logits [B,10] -> [1,B]."""
    def prep(pred, lab):
        v = torch.cat((torch.softmax(pred.detach().float().cpu(), dim=1).view(-1, 10),
                       one_hot(lab.cpu()) * label_weight), dim=1)
        v = v + MIN_VAR_EST
        return v / torch.sum(v)  # normalised by the sum over the WHOLE batch (mmd.py:151-153)
    a, b = prep(pred_s, label_s), prep(pred_t, label_t)
    d = (_kl_div(a, b) * 0.5 + _kl_div(b, a) * 0.5).sum(1)
    return distance2weights_mean2one(d).reshape(1, -1)


def mmd_cal(label_s, feat_s, label_t, feat_t, args: dict, data_s=None, data_t=None, mmd_dtype=None):
    """mmd.py:25-41, SOFT_MMD / OFF branches (56-66).

    ``mmd_dtype=torch.float64`` evaluates the kernel matrix and its autograd in double precision.
    The reference's fp32 autograd cancels catastrophically on the Gram diagonal (the sigma=0.01
    term puts -5050/m^2 into dL/dE_ii, which is added to and then subtracted from row sums of
    magnitude 1e-3): measured 170 % relative error of dL/dX against fp64 for 4096-d features, so
    gradient parity of the GPU kernel is judged against the fp64 evaluation of the same loss."""
    if mmd_dtype is not None:
        feat_s, feat_t = feat_s.to(mmd_dtype), feat_t.to(mmd_dtype)
    w = None
    if data_s is not None and (args.get("GEO_WEIGHTS") or args.get("SEM_WEIGHTS")):
        if args.get("GEO_WEIGHTS"):
            w = geometric_weights(data_s, data_t)
        else:
            w = prob_weights_soft(data_s, data_t, label_s, label_t, args["LABEL_WEIGHT"])
    if args["NAME"] == "OFF":
        return mix_rbf_mmd2(feat_s, feat_t)
    assert args["NAME"] == "SOFT_MMD"
    ls = float(args["LABEL_SCALE"])
    fs = torch.cat((feat_s, (one_hot(label_s) * ls).to(feat_s.dtype)), dim=1)
    ft = torch.cat((feat_t, (one_hot(label_t) * ls).to(feat_t.dtype)), dim=1)
    if w is not None:
        w = w.to(fs.dtype)
    return mix_rbf_mmd2(fs, ft, sample_weights=w).to(torch.float32)


# --------------------------------------------------------------------------------------
# Losses + the SUG step                                       train_dg_single_gpu.py:246-335
# --------------------------------------------------------------------------------------
class FocalLoss:
    """model_utils.py:131-176, including the reference's stateful quirk: ``self.alpha`` is
    overwritten by its gather on every call (line 168)."""

    def __init__(self, alpha, gamma=0.0):
        self.alpha = torch.tensor(alpha, dtype=torch.float32)
        self.gamma = gamma

    def __call__(self, preds, labels):
        self.alpha = self.alpha.to(preds.device)
        ls = F.log_softmax(preds.view(-1, preds.size(-1)), dim=1)
        pt = torch.exp(ls).gather(1, labels.view(-1, 1))
        lg = ls.gather(1, labels.view(-1, 1))
        self.alpha = self.alpha.gather(0, labels.view(-1))
        loss = -torch.pow(1 - pt, self.gamma) * lg
        return (self.alpha * loss.t()).mean()


SUG_CFG = {  # tools/cfgs/cfgs_sproject/DG_unified_loss_onedataset_shapenet.yaml
    "MMD_WEIGHT": 0.5, "CLS_WEIGHT": 1.0, "TARGET_LOSS": 1.0, "SRC_LOSS_WEIGHT": 1.0,
    "GEO_MMD": {"NAME": "SOFT_MMD", "LABEL_SCALE": 50, "GEO_WEIGHTS": "mean2one", "GEO_SCALE": 1},
    "SEM_MMD": {"NAME": "SOFT_MMD", "LABEL_SCALE": 5, "SEM_WEIGHTS": "mean2one", "LABEL_WEIGHT": 0.5,
                "SEM_SCALE": 1},
}


def sug_losses(sd: State, data, label, data_t, label_t, criterion, cfg=SUG_CFG, model_name="DGCNN",
               fps_starts=None, drop_p: float = 0.4, mmd_dtype=None):
    """Forward half of one SUG step, train_dg_single_gpu.py:260-324: four Net_MDA forwards,
    class-weighted CE on both heads and both sub-domains (note: the target logits are scored
    against the SOURCE labels, line 287-288), geometric + semantic MMD."""
    fs = fps_starts if fps_starts is not None else [None] * 4
    ps1, ps2, ss1, ss2 = net_mda(data, sd, True, model_name, semantic_adaption=True, fps_start=fs[0], drop_p=drop_p)
    pt1, pt2, st1, st2 = net_mda(data_t, sd, True, model_name, semantic_adaption=True, fps_start=fs[1], drop_p=drop_p)
    loss_s = 0.5 * criterion(ps1, label) + 0.5 * criterion(ps2, label)
    if cfg["TARGET_LOSS"] > 0:
        loss_t = 0.5 * criterion(pt1, label) + 0.5 * criterion(pt2, label)
        loss = 0.5 * loss_s + 0.5 * loss_t
    else:
        loss = cfg["SRC_LOSS_WEIGHT"] * loss_s
    loss_cls = cfg["CLS_WEIGHT"] * loss
    node_s = net_mda(data, sd, True, model_name, node_adaptation_s=True, fps_start=fs[2])
    node_t = net_mda(data_t, sd, True, model_name, node_adaptation_t=True, fps_start=fs[3])
    geo, sem = cfg["GEO_MMD"], cfg["SEM_MMD"]
    loss_geo = cfg["MMD_WEIGHT"] * geo["GEO_SCALE"] * mmd_cal(label, node_s, label_t, node_t, geo, data, data_t, mmd_dtype)
    l1 = sem["SEM_SCALE"] * mmd_cal(label, ss1, label_t, st1, sem, ps1, pt1, mmd_dtype)
    l2 = sem["SEM_SCALE"] * mmd_cal(label, ss2, label_t, st2, sem, ps2, pt2, mmd_dtype)
    loss_sem = cfg["MMD_WEIGHT"] * (0.5 * l1 + 0.5 * l2)
    return {"loss": loss_cls + loss_geo + loss_sem, "loss_cls": loss_cls, "loss_geo": loss_geo,
            "loss_sem": loss_sem, "pred_s1": ps1, "pred_t1": pt1}


# --------------------------------------------------------------------------------------
# Synthetic inputs and weights (numpy PCG64 => identical on every box)   SURVEY.md §8d
# --------------------------------------------------------------------------------------
def synth_clouds(B: int, N: int, seed: int):
    """PointDA-10-shaped clouds: uniform in a cube, centred, scaled to the unit sphere
    (data_utils.py:5-15 normal_pc).  Returns x [B,3,N,1] float32 and int64 labels [B]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p = rng.random((B, N, 3), dtype=np.float32) * 2 - 1
    p = p - p.mean(1, keepdims=True)
    p = p / np.sqrt((p ** 2).sum(2)).max(1)[:, None, None]
    lab = rng.integers(0, 10, size=(B,))
    x = torch.from_numpy(np.ascontiguousarray(p.transpose(0, 2, 1)[..., None]).astype(np.float32))
    return x, torch.from_numpy(lab.astype(np.int64))


def _conv2d_bn(p, cin, cout, bias):
    s = {p + ".conv.0.weight": (cout, cin, 1, 1)}
    if bias:
        s[p + ".conv.0.bias"] = (cout,)
    s.update(_bnspec(p + ".conv.1", cout))
    return s


def _bnspec(p, c):
    return {p + ".weight": (c,), p + ".bias": (c,), p + ".running_mean": (c,), p + ".running_var": (c,),
            p + ".num_batches_tracked": ()}


def _fc(p, cin, cout, bias):
    s = {p + ".fc.0.weight": (cout, cin)}
    if bias:
        s[p + ".fc.0.bias"] = (cout,)
    s[p + ".fc.1.weight"] = (cout,)
    s[p + ".fc.1.bias"] = (cout,)
    return s


def _tnet(p, cin, K):
    s = {}
    s.update(_conv2d_bn(p + ".conv2d1", cin, 64, True))
    s.update(_conv2d_bn(p + ".conv2d2", 64, 128, True))
    s.update(_conv2d_bn(p + ".conv2d3", 128, 1024, True))
    s.update(_fc(p + ".fc1", 1024, 512, False))
    s.update(_fc(p + ".fc2", 512, 256, False))
    s[p + ".fc3.weight"] = (K * K, 256)
    s[p + ".fc3.bias"] = (K * K,)
    return s


def _adapt(p):
    s = {}
    s.update(_conv2d_bn(p + ".trans", 64, 64, True))
    s[p + ".pred_offset.0.weight"] = (3, 64, 1, 1)
    s.update(_conv2d_bn(p + ".residual", 64, 64, True))
    return s


def _head(p, dgcnn):
    s = {}
    s.update(_fc(p + ".mlp1", 1024, 512, dgcnn))
    s.update(_fc(p + ".mlp2", 512, 256, True))
    s[p + ".mlp3.weight"] = (10, 256)
    s[p + ".mlp3.bias"] = (10,)
    return s


def _dgcnn_spec(p, adapt):
    s = {}
    s.update(_tnet(p + "input_transform_net", 6, 3))
    for n, (ci, co) in {"conv1": (6, 64), "conv2": (128, 64), "conv3": (128, 128), "conv4": (256, 256)}.items():
        s.update(_conv2d_bn(p + n, ci, co, False))
    s.update(_bnspec(p + "bn5", 512))
    s[p + "conv5.weight"] = (512, 512, 1)
    if adapt:
        s.update(_adapt(p + "node_fea_adapt"))
        s[p + "conv1d.weight"] = (64, 128, 1)
        s[p + "conv1d.bias"] = (64,)
    return s


def state_spec(model: str) -> Dict[str, tuple]:
    """Key -> shape tables equal to the reference modules' ``state_dict()`` (checked by
    tests/golden/make_golden.py): 'Net_MDA:DGCNN', 'Net_MDA:Pointnet', 'DGCNN_cls'."""
    s: Dict[str, tuple] = {}
    if model == "DGCNN_cls":
        s.update(_dgcnn_spec("", False))
        s.update(_head("classifier", True))
        return s
    if model == "Net_MDA:DGCNN":
        s.update(_dgcnn_spec("g.", True))
        dg = True
    elif model == "Net_MDA:Pointnet":
        s.update(_tnet("g.trans_net1", 3, 3))
        s.update(_tnet("g.trans_net2", 64, 64))
        s.update(_conv2d_bn("g.conv1", 3, 64, True))
        s.update(_conv2d_bn("g.conv2", 64, 64, True))
        s.update(_adapt("g.conv3"))
        s.update(_conv2d_bn("g.conv4", 128, 128, True))
        s.update(_conv2d_bn("g.conv5", 128, 1024, True))
        s.update(_bnspec("g.bn1", 1024))
        dg = False
    else:
        raise ValueError(model)
    for a in ("attention_s", "attention_t"):
        s[a + ".conv_du.0.weight"] = (512, 4096, 1, 1)
        s[a + ".conv_du.0.bias"] = (512,)
        s[a + ".conv_du.2.weight"] = (4096, 512, 1, 1)
        s[a + ".conv_du.2.bias"] = (4096,)
        s.update(_bnspec(a + ".bn", 4096))
    s.update(_head("c1", dg))
    s.update(_head("c2", dg))
    return s


def synth_state(model: str, seed: int = 666, neg_gamma_frac: float = 0.25) -> State:
    """Deterministic weights for ``state_spec(model)``.  Matrices ~ N(0, 1/fan_in) (close to
    the default kaiming-uniform scale), biases ~ N(0, 0.05); norm-layer scales are +-U(0.5,1.5)
    with a fraction of NEGATIVE entries so the min-instead-of-max path of the fused
    BN-monotone kernels is always exercised; running stats start at (0, 1, 0)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sd: State = {}
    spec = state_spec(model)
    for k in sorted(spec):
        shp = spec[k]
        if k.endswith("num_batches_tracked"):
            t = np.zeros((), np.int64)
        elif k.endswith("running_mean"):
            t = np.zeros(shp, np.float32)
        elif k.endswith("running_var"):
            t = np.ones(shp, np.float32)
        elif len(shp) == 1 and (".conv.1." in k or k.split(".")[-2] in ("bn5", "bn1", "bn", "1")):
            if k.endswith(".weight"):
                t = rng.uniform(0.5, 1.5, shp).astype(np.float32)
                if ".fc.1." not in k:
                    t = np.where(rng.random(shp) < neg_gamma_frac, -t, t).astype(np.float32)
            else:
                t = (rng.standard_normal(shp) * 0.1).astype(np.float32)
        elif len(shp) == 1:
            t = (rng.standard_normal(shp) * 0.05).astype(np.float32)
        else:
            fan_in = int(np.prod(shp[1:]))
            t = (rng.standard_normal(shp) / math.sqrt(fan_in)).astype(np.float32)
        sd[k] = torch.from_numpy(np.asarray(t))
    return sd


def clone_state(sd: State, requires_grad: bool = False) -> State:
    out = {}
    for k, v in sd.items():
        c = v.detach().clone()
        if requires_grad and c.is_floating_point() and "running_" not in k:
            c.requires_grad_(True)
        out[k] = c
    return out
