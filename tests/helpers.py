"""Shared helpers for the parity tests (tolerance definition, fixture loading)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def head_params(fx, kind):
    """state_dict arrays of a head fixture -> oracle/product parameter dict."""
    nl = fx["sd__anchors"].shape[0]
    p = {
        "anchors": fx["sd__anchor_grid"].reshape(nl, -1, 2),
        "w": [fx[f"sd__m__{i}__weight"][:, :, 0, 0] for i in range(nl)],
        "b": [fx[f"sd__m__{i}__bias"] for i in range(nl)],
        "ia": [fx[f"sd__ia__{i}__implicit"].reshape(-1) for i in range(nl)],
        "im": [fx[f"sd__im__{i}__implicit"].reshape(-1) for i in range(nl)],
    }
    if kind == "iaux":
        p["w2"] = [fx[f"sd__m2__{i}__weight"][:, :, 0, 0] for i in range(nl)]
        p["b2"] = [fx[f"sd__m2__{i}__bias"] for i in range(nl)]
    if kind == "ibin":
        p["bins_w"] = fx["sd__w_bin_sigmoid__bins"]
        p["bins_h"] = fx["sd__h_bin_sigmoid__bins"]
        p["bin_count"] = int(p["bins_w"].shape[0])
    return p


def box_scale(anchors, strides, shapes, na, no_out):
    """Per-row, per-column scale `s` of the parity rule |a-b| <= rtol*max(|ref|, s)
    (BASELINE.md section 4): s = stride for xy, anchor for wh, 1 for scores."""
    rows = []
    for i, (ny, nx) in enumerate(shapes):
        for a in range(na):
            s = np.ones((ny * nx, no_out), np.float32)
            s[:, 0:2] = strides[i]
            s[:, 2] = anchors[i][a][0]
            s[:, 3] = anchors[i][a][1]
            rows.append(s)
    return np.concatenate(rows, 0)[None]


def assert_close_scaled(got, ref, scale, rtol, what=""):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    tol = rtol * np.maximum(np.abs(ref), scale)
    err = np.abs(got - ref)
    bad = err > tol
    if bad.any():
        i = np.unravel_index(np.argmax(err / tol), err.shape)
        raise AssertionError(f"{what}: {bad.sum()} of {bad.size} outside rtol={rtol}; worst at {i}: "
                             f"got {got[i]!r} ref {ref[i]!r} err {err[i]:.3e} tol {tol[i]:.3e}")


def make_rep(c1, c2, stride):
    """A RepConv block in the reference's layout (nets/common.py:440-472), rebuilt for the tests (the reference module
    itself cannot travel to the GPU box)."""
    import torch
    from torch import nn

    class Rep(nn.Module):
        def __init__(self):
            super().__init__()
            self.deploy, self.groups, self.in_channels, self.out_channels = False, 1, c1, c2
            self.act = nn.SiLU()
            self.rbr_identity = nn.BatchNorm2d(c1) if c2 == c1 and stride == 1 else None
            self.rbr_dense = nn.Sequential(nn.Conv2d(c1, c2, 3, stride, 1, bias=False), nn.BatchNorm2d(c2))
            self.rbr_1x1 = nn.Sequential(nn.Conv2d(c1, c2, 1, stride, 0, bias=False), nn.BatchNorm2d(c2))

        def forward(self, x):
            if hasattr(self, "rbr_reparam"):
                return self.act(self.rbr_reparam(x))
            return self.act(self.rbr_dense(x) + self.rbr_1x1(x) + (0 if self.rbr_identity is None else self.rbr_identity(x)))

    return Rep().eval()
