"""Harness only (BASELINE.json configs[4]): the consumer of the D-TSDF volumes, i.e. the 3-D DenseNet of the
reference's training loop, rebuilt from its hyper-parameters so that it can travel to the GPU box
(`/root/reference` does not exist there).  It is cuDNN / cuBLAS work and NOT part of the product (SURVEY.md
section 2 rows 8-10: out of scope, allowed "as harness for config 5"); tools/overlap_demo.py uses it to show what
the voxeliser adds to a real training step.

Architecture (3D_CNN/network.py:53-129), input [B, 3, 32, 32, 32]:
    conv0 3 -> 16 (3x3x3, pad 1, no bias)
    four dense blocks of (2, 4, 4, 2) layers, growth 32, bottleneck 4 x growth:
        layer = BN - ReLU - conv 1x1x1 -> 128 - BN - ReLU - conv 3x3x3 -> 32, output concatenated to its input
    between blocks: BN - ReLU - conv 1x1x1 halving the channels - average pool 2   (32 -> 16 -> 8 -> 4)
    final BN (170 channels), flatten 170 x 4^3 -> 4096 -> 1024 -> 63, ReLU after the first two, dropout
    (p = 0.5, ALWAYS on: the reference calls F.dropout without passing `training`, :118,120) after each.
Training step (3D_CNN/train.py:104-108,248-256): SGD(momentum 0.9, weight decay 5e-4 in the reference's
defaults), loss = MSE(est, normalised joints) * PCA_SZ.

Module names equal the reference's, so a state_dict of one loads into the other
(tests/test_consumer_network.py checks names, shapes and the forward pass against the unmodified file where
the reference tree is present).
"""
from collections import OrderedDict

import torch
from torch import nn
import torch.nn.functional as F

GROWTH, BLOCKS, STEM, BOTTLENECK = 32, (2, 4, 4, 2), 16, 4
FLAT = 170 * 4 * 4 * 4


class _Concat(nn.Sequential):
    """x -> cat([x, body(x)], channel axis)."""

    def forward(self, x):
        return torch.cat([x, super().forward(x)], 1)


def _bn_relu_conv(prefix_names, cin, cout, k):
    n_norm, n_relu, n_conv = prefix_names
    return [(n_norm, nn.BatchNorm3d(cin)), (n_relu, nn.ReLU(inplace=True)),
            (n_conv, nn.Conv3d(cin, cout, kernel_size=k, stride=1, padding=k // 2, bias=False))]


def _features():
    mods = OrderedDict(conv0=nn.Conv3d(3, STEM, kernel_size=3, stride=1, padding=1, bias=False))
    ch = STEM
    for b, n_layers in enumerate(BLOCKS, start=1):
        block = nn.Sequential()
        for i in range(1, n_layers + 1):
            layer = _Concat(OrderedDict(
                _bn_relu_conv(("norm1", "relu1", "conv1"), ch, BOTTLENECK * GROWTH, 1) +
                _bn_relu_conv(("norm2", "relu2", "conv2"), BOTTLENECK * GROWTH, GROWTH, 3)))
            block.add_module("dense_layer%d" % i, layer)
            ch += GROWTH
        mods["dense_block%d" % b] = block
        if b != len(BLOCKS):
            mods["transition%d" % b] = nn.Sequential(OrderedDict(
                _bn_relu_conv(("norm", "relu", "conv"), ch, ch // 2, 1) + [("pool", nn.AvgPool3d(2, 2))]))
            ch //= 2
    mods["norm5"] = nn.BatchNorm3d(ch)
    return nn.Sequential(mods), ch


class DenseNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.features, ch = _features()
        assert ch * 4 ** 3 == FLAT, ch
        self.FC1 = nn.Sequential(nn.Linear(FLAT, 4096), nn.ReLU(inplace=True))
        self.FC2 = nn.Sequential(nn.Linear(4096, 1024), nn.ReLU(inplace=True))
        self.FC3 = nn.Linear(1024, 63)

    def forward(self, x):
        h = self.features(x).view(-1, FLAT)
        h = F.dropout(self.FC1(h))          # training=True whatever the mode: as the reference
        h = F.dropout(self.FC2(h))
        return self.FC3(h)


def make_training_step(net, lr=0.01, momentum=0.9, weight_decay=5e-4, loss_scale=63.0):
    """One optimiser step as the reference's non-PCA branch does it (3D_CNN/train.py:248-256)."""
    opt = torch.optim.SGD(net.parameters(), lr=lr, momentum=momentum, weight_decay=weight_decay)
    crit = nn.MSELoss()

    def step(tsdf, target):
        opt.zero_grad()
        loss = crit(net(tsdf), target) * loss_scale
        loss.backward()
        opt.step()
        return loss
    return step
