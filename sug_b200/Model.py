"""Drop-in for the reference's ``model/Model.py``: ``Net_MDA``, ``DGCNN``, ``Pointnet_g``,
``Pointnet_c``, ``CALayer`` with unchanged constructor / forward signatures and ``state_dict``
keys, the encoders running on the fused sm_100a kernels.

Out of scope (other model families): Pointnet2_g, PTran_g, KPConv_g —
``Net_MDA`` raises NotImplementedError for them.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .model_utils import adapt_layer_off, conv_2d, fc_layer, transform_net

K = 20  # Model.py:52


class CALayer(nn.Module):
    """Model.py:16-34 channel attention over the 64x64 node features."""

    def __init__(self, channel, reduction=8):
        super().__init__()
        self.conv_du = nn.Sequential(
            nn.Conv2d(channel, channel // reduction, 1, padding=0, bias=True),
            nn.ReLU(inplace=False),
            nn.Conv2d(channel // reduction, channel, 1, padding=0, bias=True),
            nn.Sigmoid())
        self.bn = nn.BatchNorm1d(4096)

    def forward(self, x):
        # the two 1x1 convolutions on a [B,C,1,1] tensor are plain linear layers: the library's fp32-accurate GEMM
        # (no cuDNN / cuBLAS kernel on the path, no CPU branch); the gate and the BatchNorm1d stay in ATen
        ops._need_cuda(x)
        c0, c2 = self.conv_du[0], self.conv_du[2]
        v = x.reshape(x.shape[0], -1)
        y = F.relu(ops.linear(v, c0.weight.view(c0.out_channels, -1), c0.bias))
        y = torch.sigmoid(ops.linear(y, c2.weight.view(c2.out_channels, -1), c2.bias))
        return self.bn(v * y + v)


def grad_reverse(x, lambd=1.0):
    """Model.py:37-50: the reference calls GradReverse(lambd).forward(x) as a plain function, so
    no gradient reversal ever happens; the identity is reproduced."""
    return x.view_as(x)


class DGCNN(nn.Module):
    """Model.py:54-121.  x [B,3,N,1] -> (feat [B,1024], node_fea [B,64,64,1][, None])."""

    def __init__(self):
        super().__init__()
        self.k = K
        self.input_transform_net = transform_net(6, 3)  # allocated, never used (Model.py:59, 79-86)
        self.conv1 = conv_2d(6, 64, kernel=1, bias=False, activation='leakyrelu')
        self.conv2 = conv_2d(64 * 2, 64, kernel=1, bias=False, activation='leakyrelu')
        self.conv3 = conv_2d(64 * 2, 128, kernel=1, bias=False, activation='leakyrelu')
        self.conv4 = conv_2d(128 * 2, 256, kernel=1, bias=False, activation='leakyrelu')
        self.bn5 = nn.BatchNorm1d(512)
        self.conv5 = nn.Conv1d(64 + 64 + 128 + 256, 512, kernel_size=1, bias=False)
        self.node_fea_adapt = adapt_layer_off()
        self.conv1d = nn.Conv1d(128, 64, 1)
        self.dim_redu = nn.MaxPool1d(3, stride=16)
        # Opt-in (off by default; env SUG_B200_SHARE_TRUNK=1 or ``share_trunk = True``): the SUG step runs
        # this encoder TWICE on every cloud batch (train_dg_single_gpu.py:260-261 and 297-298).  Everything
        # before the adapt layer (two kNN graphs, conv1, conv2) is a deterministic function of the input
        # and the weights -- only the FPS start of the node layer differs between the two passes -- so in
        # training mode the second pass may reuse x1 / x2 of the first.  BatchNorm side effects are replayed
        # (second momentum update with the same batch statistics, num_batches_tracked).  bench.py's headline
        # number does NOT use this (it times the four full forwards the reference runs).
        self.share_trunk = os.environ.get("SUG_B200_SHARE_TRUNK", "0") == "1"
        # FPS / ball query of the adapt layer on a side stream next to conv1 / conv2 (same work, same results)
        self.overlap_fps = os.environ.get("SUG_B200_OVERLAP_FPS", "1") == "1"
        self._trunk = {}  # key -> (x1, x2, batch statistics of conv1 / conv2); at most two batches (source, target)

    def _trunk_key(self, x):
        ps = (self.conv1.conv[0].weight, self.conv1.conv[1].weight, self.conv1.conv[1].bias,
              self.conv2.conv[0].weight, self.conv2.conv[1].weight, self.conv2.conv[1].bias)
        return (x.data_ptr(), x._version, tuple(x.shape), torch.is_grad_enabled(), tuple(p._version for p in ps))

    def _trunk_fwd(self, x_loc, x0):
        """x1, x2 of Model.py:88-94 (and nothing else)."""
        k = self.k
        if not (self.share_trunk and self.training):
            x1 = self.conv1.edgeconv(x0, ops.knn_cm(x_loc, k))
            return x1, self.conv2.edgeconv(x1, ops.knn_pm(x1, k))
        key = self._trunk_key(x_loc)
        hit = self._trunk.pop(key, None)  # one reuse per batch: the pair of passes of one step
        if hit is not None:
            x1, x2, stats = hit
            with torch.no_grad():
                for bn, (s_m, s_v) in zip((self.conv1.conv[1], self.conv2.conv[1]), stats):
                    m = bn.momentum
                    bn.running_mean.mul_(1 - m).add_(s_m, alpha=m)  # the same batch statistics once more
                    bn.running_var.mul_(1 - m).add_(s_v, alpha=m)
                    bn.num_batches_tracked.add_(1)
            return x1, x2
        if len(self._trunk) >= 2:
            self._trunk.clear()
        bns = (self.conv1.conv[1], self.conv2.conv[1])
        before = [(bn.running_mean.clone(), bn.running_var.clone()) for bn in bns]
        x1 = self.conv1.edgeconv(x0, ops.knn_cm(x_loc, k))
        x2 = self.conv2.edgeconv(x1, ops.knn_pm(x1, k))
        with torch.no_grad():  # r1 = (1-m) r0 + m s  =>  s, the batch statistic this pass folded in
            stats = [((bn.running_mean - (1 - bn.momentum) * rm0) / bn.momentum,
                      (bn.running_var - (1 - bn.momentum) * rv0) / bn.momentum) for bn, (rm0, rv0) in zip(bns, before)]
        self._trunk[key] = (x1, x2, stats)
        return x1, x2

    def _tail(self, x_cat):
        bn = ops.bn_tick(self.bn5, self.training)
        return ops.mlp_bn_act_pool(x_cat, self.conv5.weight, None, bn.weight, bn.bias, bn.running_mean,
                                   bn.running_var, self.training, 0.2, ops.POOL_MAX_AVG, bn.eps, bn.momentum)

    def forward(self, x, node=False):
        x_loc = x.squeeze(-1)  # [B,3,N]
        B = x.size(0)
        k = self.k
        x0 = x_loc.transpose(1, 2).contiguous()  # point-major [B,N,3]
        # FPS + ball query of the node layer need the cloud only: start them now, next to conv1 / conv2
        pre = self.node_fea_adapt.prefetch_indices(x_loc) if self.overlap_fps else None
        if self.share_trunk and self.training:
            x1, x2 = self._trunk_fwd(x_loc, x0)
            cat = None
        else:
            # x1, x2', x3, x4 are written straight into the channel slices of the concatenated tensor of
            # Model.py:111 (no torch.cat copy; the backward hands out slices of its gradient)
            cat = torch.empty(B, x0.shape[1], 512, dtype=torch.float32, device=x0.device)
            x1 = self.conv1.edgeconv(x0, ops.knn_cm(x_loc, k), out=cat[:, :, 0:64])
            x2 = self.conv2.edgeconv(x1, ops.knn_pm(x1, k))
        x_, node_pm, _ = self.node_fea_adapt.forward_pm(x2, x_loc, pre=pre)  # [B,N,128], [B,64 nodes,64]
        node_fea = node_pm.transpose(1, 2).unsqueeze(3)                       # reference layout [B,64,64,1]
        if cat is None:
            x2 = ops.linear(x_, self.conv1d.weight, self.conv1d.bias)         # Conv1d(128,64,1) on point-major rows
            x3 = self.conv3.edgeconv(x2, ops.knn_pm(x2, k))
            x4 = self.conv4.edgeconv(x3, ops.knn_pm(x3, k))
            feat = self._tail(torch.cat((x1, x2, x3, x4), dim=2))
        else:
            x2 = ops.linear(x_, self.conv1d.weight, self.conv1d.bias, out=cat[:, :, 64:128])
            x3 = self.conv3.edgeconv(x2, ops.knn_pm(x2, k), out=cat[:, :, 128:256])
            x4 = self.conv4.edgeconv(x3, ops.knn_pm(x3, k), out=cat[:, :, 256:512])
            feat = self._tail(ops.join_slices(cat, x1, x2, x3, x4))
        if node:
            return feat, node_fea, None
        return feat, node_fea


class Pointnet_g(nn.Module):
    """Model.py:235-283.  x [B,3,N,1] -> (feat [B,1024], node_fea [B,64,64,1][, node_off [B,3,64]])."""

    def __init__(self):
        super().__init__()
        self.trans_net1 = transform_net(3, 3)
        self.trans_net2 = transform_net(64, 64)
        self.conv1 = conv_2d(3, 64, 1)
        self.conv2 = conv_2d(64, 64, 1)
        self.conv3 = adapt_layer_off()
        self.conv4 = conv_2d(128, 128, 1)
        self.conv5 = conv_2d(128, 1024, 1)
        self.bn1 = nn.BatchNorm1d(1024)

    def forward(self, x, node=False):
        x_loc = x.squeeze(-1)
        transform = self.trans_net1(x)
        x = torch.bmm(x.transpose(2, 1).squeeze(-1), transform).unsqueeze(3).transpose(2, 1)
        x = self.conv2(self.conv1(x))
        transform = self.trans_net2(x)
        x = torch.bmm(x.transpose(2, 1).squeeze(-1), transform).unsqueeze(3).transpose(2, 1)
        x, node_fea, node_off = self.conv3(x, x_loc)
        x = self.conv4(x)
        x = self.conv5.pool_max(x.squeeze(3).transpose(1, 2))  # conv5 + BN + ReLU + max over N, fused
        x = self.bn1(x)
        if node:
            return x, node_fea, node_off
        return x, node_fea


class Pointnet_c(nn.Module):
    """Model.py:412-449 classifier head (LayerNorm MLP + Dropout2d)."""

    def __init__(self, num_class=10, dgcnn_flag=False, PTran_flag=False):
        super().__init__()
        activate, bias = ('leakyrelu', True) if dgcnn_flag else ('relu', False)
        self.mlp1 = fc_layer(1024, 512, bn=True, activation=activate, bias=bias)
        self.dropout1 = nn.Dropout2d(p=0.4)
        self.mlp2 = fc_layer(512, 256, bn=True, activation=activate, bias=True)
        self.dropout2 = nn.Dropout2d(p=0.4)
        self.mlp3 = nn.Linear(256, num_class)
        self.PTran = PTran_flag

    def forward(self, x, adapt=False):
        if not self.PTran:
            x = self.dropout1(self.mlp1(x))
        x = self.mlp2(x)
        mid_feature = x
        x = self.dropout2(x)
        x = ops.linear(x, self.mlp3.weight, self.mlp3.bias)
        if adapt is False or adapt == False:  # noqa: E712  (reference compares with ==)
            return x
        return x, mid_feature


class Net_MDA(nn.Module):
    """Model.py:452-520."""

    def __init__(self, model_name='Pointnet'):
        super().__init__()
        self.dgcnn_flag = False
        self.PTran_flag = False
        if model_name == 'Pointnet':
            self.g = Pointnet_g()
        elif model_name == 'DGCNN':
            self.g = DGCNN()
            self.dgcnn_flag = True
        elif model_name in ('Pointnet2', 'PTran', 'KPConv'):
            raise NotImplementedError(f"backbone {model_name!r} is outside the accelerated hot path "
                                      "(SURVEY.md §2); use the reference implementation for it")
        else:
            raise NotImplementedError("Unsupported model name")
        self.attention_s = CALayer(64 * 64)
        self.attention_t = CALayer(64 * 64)
        self.c1 = Pointnet_c(dgcnn_flag=self.dgcnn_flag, PTran_flag=self.PTran_flag)
        self.c2 = Pointnet_c(dgcnn_flag=self.dgcnn_flag, PTran_flag=self.PTran_flag)

    def forward(self, x, constant=1, adaptation=False, node_vis=False, mid_feat=False, node_adaptation_s=False,
                node_adaptation_t=False, semantic_adaption=False):
        x, feat_ori, node_idx = self.g(x, node=True)
        batch_size = feat_ori.size(0)
        if node_vis:
            return node_idx
        if mid_feat:
            return x, feat_ori
        if node_adaptation_s:
            feat_node = feat_ori.contiguous().view(batch_size, -1)
            return self.attention_s(feat_node.unsqueeze(2).unsqueeze(3))
        elif node_adaptation_t:
            feat_node = feat_ori.contiguous().view(batch_size, -1)
            return self.attention_t(feat_node.unsqueeze(2).unsqueeze(3))
        if adaptation:
            x = grad_reverse(x, constant)
        if not semantic_adaption:
            return self.c1(x, adapt=semantic_adaption), self.c2(x, adapt=semantic_adaption)
        y1, sem_feature1 = self.c1(x, adapt=semantic_adaption)
        y2, sem_feature2 = self.c2(x, adapt=semantic_adaption)
        return y1, y2, sem_feature1, sem_feature2
