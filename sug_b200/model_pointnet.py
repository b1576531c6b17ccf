"""Drop-in for ``model_pointnet.DGCNN`` of the reference (model/model_pointnet.py:93-161): DGCNN
without the node layer, any N, with the classifier head.  This is the model of BASELINE.json's
config 1 (CPU-runnable case) and the only DGCNN usable at LiDAR scale (config 5)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .Model import K, Pointnet_c
from .model_utils import conv_2d, transform_net


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
        bn = self.bn5
        if self.training and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
        feat = ops.mlp_bn_act_pool(torch.cat((x1, x2, x3, x4), dim=2), self.conv5.weight, None, bn.weight, bn.bias,
                                   bn.running_mean, bn.running_var, self.training, 0.2, ops.POOL_MAX_AVG, bn.eps,
                                   bn.momentum)
        return self.classifier(feat)
