"""Cost-to-go network with the reference's parameter layout, so the shipped `model_state_dict.pt` files load.

Mirror of utils/pytorch_models.py:5-86 (`ResnetModel`): one-hot -> fc1(h1)+BN+ReLU -> fc2(res)+BN+ReLU ->
N x [fc, BN, ReLU, fc, BN, +skip, ReLU] -> fc_out.  Parameter names (`fc1`, `bn1`, `fc2`, `bn2`,
`blocks.<i>.<0..3>`, `fc_out`) are part of the checkpoint format and therefore identical.
This module is the fp32 PyTorch reference of the network; the inference path the search uses is
deepcubea_b200/nnet/folded.py (BN folded, tensor-core GEMMs).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn


class ResnetModel(nn.Module):
    def __init__(self, state_dim: int, one_hot_depth: int, h1_dim: int, resnet_dim: int, num_resnet_blocks: int,
                 out_dim: int, batch_norm: bool):
        super().__init__()
        self.state_dim, self.one_hot_depth = state_dim, one_hot_depth
        self.num_resnet_blocks, self.batch_norm = num_resnet_blocks, batch_norm
        in_dim = state_dim * one_hot_depth if one_hot_depth > 0 else state_dim
        self.fc1 = nn.Linear(in_dim, h1_dim)
        if batch_norm:
            self.bn1 = nn.BatchNorm1d(h1_dim)
        self.fc2 = nn.Linear(h1_dim, resnet_dim)
        if batch_norm:
            self.bn2 = nn.BatchNorm1d(resnet_dim)
        self.blocks = nn.ModuleList()
        for _ in range(num_resnet_blocks):
            layers = [nn.Linear(resnet_dim, resnet_dim)]
            if batch_norm:
                layers.append(nn.BatchNorm1d(resnet_dim))
            layers.append(nn.Linear(resnet_dim, resnet_dim))
            if batch_norm:
                layers.append(nn.BatchNorm1d(resnet_dim))
            self.blocks.append(nn.ModuleList(layers))
        self.fc_out = nn.Linear(resnet_dim, out_dim)

    def _encode(self, states_nnet: torch.Tensor) -> torch.Tensor:
        if self.one_hot_depth > 0:
            return F.one_hot(states_nnet.long(), self.one_hot_depth).float().flatten(1)
        return states_nnet.float()

    def forward(self, states_nnet: torch.Tensor) -> torch.Tensor:
        bn = self.batch_norm
        x = self.fc1(self._encode(states_nnet))
        x = F.relu(self.bn1(x) if bn else x)
        x = self.fc2(x)
        x = F.relu(self.bn2(x) if bn else x)
        for blk in self.blocks:
            skip = x
            if bn:
                x = F.relu(blk[1](blk[0](x)))
                x = blk[3](blk[2](x))
            else:
                x = blk[1](F.relu(blk[0](x)))
            x = F.relu(x + skip)
        return self.fc_out(x)
