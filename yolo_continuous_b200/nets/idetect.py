"""IDetect head (drop-in for reference nets/idetect.py) on the B200 kernels."""
from .. import _lib
from ._head import HeadBase


class IDetect(HeadBase):
    """YOLOv7 lead head.  forward(x: list of [bs,ch_i,H_i,W_i]) ->
    train: list of [bs,na,ny,nx,no] raw maps; eval: (z [bs, sum(na*H*W), no], that list).
    Reference: nets/idetect.py:11-45.  The input list is updated in place like the reference's
    `x[i] = ...` (nets/idetect.py:31-34).
    """

    def __init__(self, nc=80, anchors=(), ch=()):
        super().__init__()
        self._init_common(nc, anchors, nc + 5)
        self._make_lead(ch)

    def forward(self, x):
        self.training |= self.export
        nl = self.nl
        if self.training:
            _, raws = self._run(x[:nl], self.m, self.ia, self.im, _lib.YC_HEAD_RAW, False, True)
            for i in range(nl):
                x[i] = raws[i]
            return x
        z, raws = self._run(x[:nl], self.m, self.ia, self.im, _lib.YC_HEAD_IDETECT, True, self.return_raw)
        for i in range(nl):
            self._update_grid_cache(i, x[i].shape[2], x[i].shape[3], z.device)
        if self.return_raw:
            for i in range(nl):
                x[i] = raws[i]
        return z, x
