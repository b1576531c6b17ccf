"""Drop-in for the reference's ``model/model_pointnet.py``: ``DGCNN`` (lines 93-161: DGCNN without the
node layer, any N, with the classifier head -- the model of BASELINE.json's config 1 and the only DGCNN
usable at LiDAR scale, config 5), ``Pointnet_cls`` (lines 5-55: the stand-alone PointNet classifier that
``train_dg_single_gpu.py:8`` and ``dataset_splitter.py:6`` import) and ``Pointnet2_cls`` (lines 58-90:
PointNet++, another model family -- out of scope; constructing it raises)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .Model import K, Pointnet_c
from .model_utils import conv_2d, fc_layer, transform_net


class Pointnet_cls(nn.Module):
    """model_pointnet.py:5-55.  x [B,3,N,1] -> logits [B,num_class] (``adapt=True``: also the 1024-d feature)."""

    def __init__(self, num_class=10):
        super().__init__()
        self.trans_net1 = transform_net(3, 3)
        self.trans_net2 = transform_net(64, 64)
        self.conv1 = conv_2d(3, 64, 1)
        self.conv2 = conv_2d(64, 64, 1)
        self.conv3 = conv_2d(64, 64, 1)
        self.conv4 = conv_2d(64, 128, 1)
        self.conv5 = conv_2d(128, 1024, 1)
        self.mlp1 = fc_layer(1024, 512)
        self.dropout1 = nn.Dropout2d(p=0.7)
        self.mlp2 = fc_layer(512, 256)
        self.dropout2 = nn.Dropout2d(p=0.7)
        self.mlp3 = nn.Linear(256, num_class)

    def forward(self, x, adapt=False):
        transform = self.trans_net1(x)
        x = torch.bmm(x.transpose(2, 1).squeeze(-1), transform).unsqueeze(3).transpose(2, 1)
        x = self.conv2(self.conv1(x))
        transform = self.trans_net2(x)
        x = torch.bmm(x.transpose(2, 1).squeeze(-1), transform).unsqueeze(3).transpose(2, 1)
        x = self.conv4(self.conv3(x))
        x = self.conv5.pool_max(x.squeeze(3).transpose(1, 2))  # conv5 + BN + ReLU + max over N (lines 42-43), fused
        mid_feature = x
        x = self.dropout1(self.mlp1(x))
        x = self.dropout2(self.mlp2(x))
        x = ops.linear(x, self.mlp3.weight, self.mlp3.bias)
        if adapt is False or adapt == False:  # noqa: E712  (the reference compares with ==)
            return x
        return x, mid_feature


class Pointnet2_cls(nn.Module):
    """model_pointnet.py:58-90 (PointNet++ set abstraction): a different model family, outside the accelerated
    hot path.  The name exists so that ``from model.model_pointnet import *`` keeps
    working; use the reference implementation for it."""

    def __init__(self, num_class=10, normal_channel=False):
        super().__init__()
        raise NotImplementedError("Pointnet2_cls (PointNet++) is outside the accelerated hot path (SURVEY.md §2); "
                                  "use the reference implementation for it")


class DGCNN(nn.Module):
    def __init__(self):
        super().__init__()
        self.k = K
        self.input_transform_net = transform_net(6, 3)
        self.conv1 = conv_2d(6, 64, kernel=1, bias=False, activation='leakyrelu')
        self.conv2 = conv_2d(64 * 2, 64, kernel=1, bias=False, activation='leakyrelu')
        self.conv3 = conv_2d(64 * 2, 128, kernel=1, bias=False, activation='leakyrelu')
        self.conv4 = conv_2d(128 * 2, 256, kernel=1, bias=False, activation='leakyrelu')
        self.bn5 = nn.BatchNorm1d(512)
        self.conv5 = nn.Conv1d(64 + 64 + 128 + 256, 512, kernel_size=1, bias=False)
        self.classifier = Pointnet_c(dgcnn_flag=True)

    def forward(self, x, node=False):
        x_loc = x.squeeze(-1)
        k = self.k
        x0 = x_loc.transpose(1, 2).contiguous()
        x1 = self.conv1.edgeconv(x0, ops.knn_cm(x_loc, k))
        x2 = self.conv2.edgeconv(x1, ops.knn_pm(x1, k))
        x3 = self.conv3.edgeconv(x2, ops.knn_pm(x2, k))
        x4 = self.conv4.edgeconv(x3, ops.knn_pm(x3, k))
        bn = ops.bn_tick(self.bn5, self.training)
        feat = ops.mlp_bn_act_pool(torch.cat((x1, x2, x3, x4), dim=2), self.conv5.weight, None, bn.weight, bn.bias,
                                   bn.running_mean, bn.running_var, self.training, 0.2, ops.POOL_MAX_AVG, bn.eps,
                                   bn.momentum)
        return self.classifier(feat)
