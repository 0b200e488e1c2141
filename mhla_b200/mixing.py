"""Host-side construction of the block-mixing matrices W (rows A0 / B0 of SURVEY.md 8a).

``BlockDistanceConv`` / ``BlockDistanceConv3D`` keep the reference's module names, constructor arguments and
``state_dict`` key (``conv.weight`` of shape [M, M, 1, 1]) - mhla_dit/mhla/mhla.py:10-138 and
mhla_videogen/diffusion/model/wan/mhla_utils.py:9-125 - so checkpoints load unchanged and the trainers'
``piece_attn.conv.weight`` clamps keep working.  The weights are built vectorised (the reference uses an O(M^2)
Python loop).
"""
from __future__ import annotations

import math
from typing import Sequence

import torch
from torch import nn


def block_distance_matrix(blocks_layout: Sequence[int], transform: str = "linear", local_thres: float = 1.5,
                          exp_sigma: float = 3.0) -> torch.Tensor:
    """W[M, M] from Euclidean distances between block centres on a 2-D / 3-D block grid (raster order)."""
    axes = [torch.arange(int(n), dtype=torch.float32) + 0.5 for n in blocks_layout]
    centres = torch.stack([g.reshape(-1) for g in torch.meshgrid(*axes, indexing="ij")], dim=-1)
    dist = torch.linalg.vector_norm(centres[:, None, :] - centres[None, :, :], ord=2, dim=-1)
    if transform == "linear":
        mat = 1.0 - dist / dist.max()
    elif transform == "cos":
        mat = torch.cos(dist / dist.max() * math.pi / 4)
    elif transform == "exp":
        mat = torch.exp(-dist / exp_sigma)
    elif transform == "gaussian":
        sigma = dist.max() / 3
        return torch.exp(-(dist ** 2) / (2 * sigma ** 2))          # not normalised in the reference
    elif transform == "local":
        mat = (dist <= local_thres).float()
    else:
        raise ValueError(f"Unknown transform: {transform}")
    return mat / mat.sum(dim=0, keepdim=True)


class _BlockMix(nn.Module):
    def _make(self, layout, transform, local_thres, exp_sigma):
        self.total_blocks = int(math.prod(layout))
        self.conv = nn.Conv2d(self.total_blocks, self.total_blocks, kernel_size=1, bias=False)
        with torch.no_grad():
            w = block_distance_matrix(layout, transform, local_thres, exp_sigma)
            self.conv.weight.data = w.unsqueeze(-1).unsqueeze(-1)

    def forward(self, x):
        """Plain 1x1 conv over the block axis (kept for API parity; the fused operator consumes the weight directly)."""
        return self.conv(x)

    def get_weight_matrix(self):
        return self.conv.weight.data.squeeze(-1).squeeze(-1)


class BlockDistanceConv(_BlockMix):
    def __init__(self, num_patches_per_side=16, patch_group_size=16, transform="linear", local_thres=1.5, exp_sigma=3):
        super().__init__()
        self.num_patches_per_side = num_patches_per_side
        self.patch_group_size = patch_group_size
        self.transform = transform
        self.local_thres = local_thres
        self.exp_sigma = exp_sigma
        self.blocks_per_side = num_patches_per_side // int(math.sqrt(patch_group_size))
        self._make((self.blocks_per_side, self.blocks_per_side), transform, local_thres, exp_sigma)


class BlockDistanceConv3D(_BlockMix):
    def __init__(self, blocks_layout=(4, 4, 4), transform="linear", local_thres=1.5, exp_sigma=3):
        super().__init__()
        self.blocks_layout = tuple(int(x) for x in blocks_layout)
        self.transform = transform
        self.local_thres = local_thres
        self.exp_sigma = exp_sigma
        self._make(self.blocks_layout, transform, local_thres, exp_sigma)
