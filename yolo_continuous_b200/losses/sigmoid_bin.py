"""SigmoidBin parameter holder (reference losses/sigmoid_bin.py:10-63, forward only).

IBin's fused kernel consumes `bins`; this module exists so that IBin's state_dict carries the
same buffers as the reference (`w_bin_sigmoid.bins`, `w_bin_sigmoid.BCE_bins.pos_weight`, ...).
`training_loss` (losses/sigmoid_bin.py:65-96) is training-only and out of scope.
"""
import torch
from torch import nn


class SigmoidBin(nn.Module):
    def __init__(self, bin_count=10, min=0.0, max=1.0, reg_scale=2.0, use_loss_regression=True,
                 use_fw_regression=True, bce_weight=1.0, smooth_eps=0.0):
        super().__init__()
        self.bin_count = bin_count
        self.length = bin_count + 1
        self.min, self.max = min, max
        self.scale = float(max - min)
        self.shift = self.scale / 2.0
        self.use_loss_regression = use_loss_regression
        self.use_fw_regression = use_fw_regression
        self.reg_scale = reg_scale
        self.BCE_weight = bce_weight
        self.step = self.scale / self.bin_count
        start = min + (self.scale / 2.0) / self.bin_count
        # same values as torch.range(start, end + 0.0001, step).float(), losses/sigmoid_bin.py:33-38
        bins = (start + self.step * torch.arange(bin_count, dtype=torch.float64)).float()
        self.register_buffer('bins', bins)
        self.cp = 1.0 - 0.5 * smooth_eps
        self.cn = 0.5 * smooth_eps
        self.BCE_bins = nn.BCEWithLogitsLoss(pos_weight=torch.Tensor([bce_weight]))
        self.MSE = nn.MSELoss()

    def get_length(self):
        return self.length

    def forward(self, pred):
        """Stand-alone decode of already-sigmoided values (not on the hot path; IBin fuses it)."""
        reg = (pred[..., 0] * self.reg_scale - self.reg_scale / 2.0) * self.step
        idx = pred[..., 1:1 + self.bin_count].argmax(-1)
        out = self.bins[idx] + reg if self.use_fw_regression else self.bins[idx]
        return out.clamp(min=self.min, max=self.max)
