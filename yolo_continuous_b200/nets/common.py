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


# ---- RepConv re-parameterisation (reference nets/common.py:541-614), SURVEY.md section 8(f) rank 4 -----------------
def _fold_conv_bn(weight, bn):
    """conv (no bias) followed by BatchNorm in eval -> (weight, bias) of one conv: w * gamma / std, beta - mean * gamma /
    std, std = sqrt(var + eps) (RepConv.fuse_conv_bn, nets/common.py:541-563), same operation order."""
    std = (bn.running_var + bn.eps).sqrt()
    bias = bn.bias - bn.running_mean * bn.weight / std
    t = (bn.weight / std).reshape(-1, 1, 1, 1)
    return weight * t, bias


@torch.no_grad()
def repconv_equivalent(rep):
    """Weight [c2, c1/g, 3, 3] and bias [c2] of the single 3x3 convolution a RepConv block collapses to:
    3x3 conv+BN  +  1x1 conv+BN zero-padded to 3x3  +  identity BN written as a 1x1 identity conv + BN
    (RepConv.fuse_repvgg_block, nets/common.py:565-614).  `rep` is a reference-style RepConv (attributes rbr_dense,
    rbr_1x1 as Sequential(conv, bn), rbr_identity a BatchNorm2d or None, in_channels / out_channels / groups); it is
    not modified.  Runs on whatever device the parameters live on (load-time tensor algebra, a few MB)."""
    w3, b3 = _fold_conv_bn(rep.rbr_dense[0].weight, rep.rbr_dense[1])
    w1, b1 = _fold_conv_bn(rep.rbr_1x1[0].weight, rep.rbr_1x1[1])
    w1 = torch.nn.functional.pad(w1, [1, 1, 1, 1])
    ident = getattr(rep, "rbr_identity", None)
    if isinstance(ident, (nn.BatchNorm2d, nn.SyncBatchNorm)):
        # the reference builds the identity as a Conv2d(c1, c2, 1, groups=g) weight filled with a diagonal
        wi = torch.zeros((rep.out_channels, rep.in_channels // rep.groups), dtype=w3.dtype, device=w3.device)
        wi.fill_diagonal_(1.0)
        wi, bi = _fold_conv_bn(wi.unsqueeze(2).unsqueeze(3), ident)
        wi = torch.nn.functional.pad(wi, [1, 1, 1, 1])
    else:
        wi, bi = torch.zeros_like(w1), torch.zeros_like(b1)
    return w3 + w1 + wi, b3 + b1 + bi


@torch.no_grad()
def fuse_repvgg_block(rep):
    """In-place form with the reference's side effects (nets/common.py:565-614): installs `rbr_reparam`, sets
    deploy = True and drops the three training branches."""
    if getattr(rep, "deploy", False):
        return rep
    w, b = repconv_equivalent(rep)
    c = rep.rbr_dense[0]
    fused = nn.Conv2d(c.in_channels, c.out_channels, c.kernel_size, c.stride, c.padding, c.dilation, c.groups, bias=True,
                      padding_mode=c.padding_mode).to(device=w.device, dtype=w.dtype)
    fused.weight, fused.bias = nn.Parameter(w), nn.Parameter(b)
    rep.rbr_reparam = fused
    rep.deploy = True
    rep.rbr_identity = rep.rbr_1x1 = rep.rbr_dense = None
    return rep
