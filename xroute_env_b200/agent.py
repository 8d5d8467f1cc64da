"""Batched front end of the reference agents' ``RepresentationNetwork``.

The reference encodes an observation with a Python loop: one forward pass of a small 3-D
convolutional tower per remaining net, batch size 1, on tensors rebuilt from numpy every call
(``/root/reference/baseline/baseline_utils.py:231-379``; 0.98 s per ``select_action`` on a T1-7x7
region).  Once the environment runs on the GPU that loop is the bottleneck, so this module applies
the *same* network -- same sub-module names and parameter shapes, reference checkpoints load with
``load_state_dict`` -- to every (environment, net) block of a ``VecGame`` observation batch at once,
reading the 7-channel blocks straight from the library-owned buffer (DLPack view, no copy to the host).

In eval mode (BatchNorm on running statistics) the outputs equal the reference's per-net results
(``tests/test_agent_frontend.py``); in train mode BatchNorm sees the whole batch instead of one net.
Plain PyTorch: this is agent-side plumbing, not part of the hot path.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class ResidualBlock(nn.Module):
    """conv-bn-relu-conv-bn + skip, size preserving (``baseline_utils.py:208-229``)."""

    def __init__(self, num_channels: int):
        super().__init__()
        self.conv1 = nn.Conv3d(num_channels, num_channels, 3, 1, 1)
        self.bn1 = nn.BatchNorm3d(num_channels)
        self.conv2 = nn.Conv3d(num_channels, num_channels, 3, 1, 1)
        self.bn2 = nn.BatchNorm3d(num_channels)

    def forward(self, x):
        out = F.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return F.relu(out + x)


def _align(x: torch.Tensor, conv: nn.Conv3d, standard) -> torch.Tensor:
    """The reference's ``clip`` (``baseline_utils.py:127-205``): one strided convolution that brings every
    spatial size to at most ``standard`` (stride = ceil(excess / standard) + 1 per axis), zero-padded at the
    far end up to exactly ``standard``."""
    shape = x.shape[-3:]
    stride = [math.ceil(max(0, shape[i] - standard[i]) / standard[i]) + 1 for i in range(3)]
    y = F.conv3d(x, conv.weight, conv.bias, stride=stride, padding=conv.padding)
    pad = []
    for i in (2, 1, 0):
        pad += [0, max(0, standard[i] - y.shape[-3 + i])]
    return F.pad(y, pad)


class BatchedRepresentationNetwork(nn.Module):
    """Same parameters as the reference ``RepresentationNetwork``; batched forward.

    ``forward(obs, n_remaining)``: ``obs`` float32 ``[N, 2+7*max_nets, Z, Y, X]`` (``VecGame.obs_batch()``),
    ``n_remaining`` int ``[N]``.  Returns ``(obstacle_rep [N, 64], net_rep [N, max_nets, 64], valid [N, max_nets])``;
    ``net_rep[e, r]`` is the encoding of the net of rank ``r`` (id ``obs[e, 1].flatten()[r]``) and is zero where
    ``valid`` is false."""

    net_input_channels = 7
    standard_net_shape = (3, 64, 64)

    def __init__(self):
        super().__init__()
        c, s = self.net_input_channels, self.standard_net_shape
        self.net_conv1 = ResidualBlock(c)
        self.net_align_conv1 = nn.Conv3d(c, c, 5, 1, 1)
        self.net_conv2 = ResidualBlock(c)
        self.net_align_conv2 = nn.Conv3d(c, 1, (s[0], s[1], 3), 1, (0, 0, 1))
        self.ob_conv1 = ResidualBlock(1)
        self.ob_align_conv1 = nn.Conv3d(1, c, 5, 1, 1)
        self.ob_conv2 = ResidualBlock(c)
        self.ob_align_conv2 = nn.Conv3d(c, 1, (s[0], s[1], 3), 1, (0, 0, 1))

    def encode_nets(self, blocks: torch.Tensor) -> torch.Tensor:
        """``[M, 7, Z, Y, X]`` -> ``[M, 64]``"""
        x = self.net_conv1(blocks)
        x = _align(x, self.net_align_conv1, self.standard_net_shape)
        x = self.net_conv2(x)
        return self.net_align_conv2(x).flatten(1)

    def encode_obstacles(self, grids: torch.Tensor) -> torch.Tensor:
        """``[N, 1, Z, Y, X]`` -> ``[N, 64]``"""
        x = self.ob_conv1(grids)
        x = _align(x, self.ob_align_conv1, self.standard_net_shape)
        x = self.ob_conv2(x)
        return self.ob_align_conv2(x).flatten(1)

    def forward(self, obs: torch.Tensor, n_remaining: torch.Tensor, chunk: int = 2048):
        N, C = obs.shape[:2]
        max_nets = (C - 2) // 7
        ob = self.encode_obstacles(obs[:, 0:1])
        valid = torch.arange(max_nets, device=obs.device)[None, :] < n_remaining.to(obs.device)[:, None]
        idx = valid.nonzero()
        nets = obs[:, 2:2 + 7 * max_nets].unflatten(1, (max_nets, 7))               # view of the env buffer
        rep = obs.new_zeros((N, max_nets, ob.shape[1]))
        for s in range(0, idx.shape[0], chunk):
            e, r = idx[s:s + chunk, 0], idx[s:s + chunk, 1]
            rep[e, r] = self.encode_nets(nets[e, r]).to(rep.dtype)      # (under autocast the tower returns bf16)
        return ob, rep, valid
