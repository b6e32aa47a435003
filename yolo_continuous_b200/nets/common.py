"""ImplicitA / ImplicitM parameter holders (reference nets/common.py:416-439).

Inside the B200 heads these are never executed as separate passes: ImplicitA is folded into
the conv bias and ImplicitM is the epilogue scale of the head kernel (csrc/yc_head_*.cu).
They remain nn.Modules so that `state_dict` keys (`ia.{i}.implicit`, `im.{i}.implicit`)
and initialisation (both N(mean=0, std=.02), reference nets/common.py:417,423,430,436)
match the reference.
"""
import torch
from torch import nn


class _Implicit(nn.Module):
    _fill = 0.0

    def __init__(self, channel, mean=0., std=.02):
        super().__init__()
        self.channel, self.mean, self.std = channel, mean, std
        self.implicit = nn.Parameter(torch.full((1, channel, 1, 1), self._fill))
        nn.init.normal_(self.implicit, mean=mean, std=std)


class ImplicitA(_Implicit):
    """x + implicit (reference nets/common.py:425-426)."""

    def forward(self, x):
        return self.implicit + x


class ImplicitM(_Implicit):
    """x * implicit (reference nets/common.py:438-439)."""
    _fill = 1.0

    def forward(self, x):
        return self.implicit * x
