"""IAuxDetect head (drop-in for reference nets/iaux_detect.py) on the B200 kernels."""
from torch import nn

from .. import _lib
from ._head import HeadBase


class IAuxDetect(HeadBase):
    """Lead head on x[:nl] plus auxiliary 1x1 convs `m2` on x[nl:2nl].

    Reference: nets/iaux_detect.py:11-49.  In eval mode the reference still runs the aux convs
    and writes them into the caller's list before discarding them (`return (cat(z,1), x[:nl])`,
    nets/iaux_detect.py:37-38,49); set `compute_aux_in_eval=False` to skip that dead work.
    """
    compute_aux_in_eval = True

    def __init__(self, nc=80, anchors=(), ch=()):
        super().__init__()
        self._init_common(nc, anchors, nc + 5)
        self._make_lead(ch[:self.nl])
        self.m2 = nn.ModuleList(nn.Conv2d(c, self.no * self.na, 1) for c in ch[self.nl:])

    def _aux(self, x):
        nl = self.nl
        _, aux = self._run(x[nl:2 * nl], self.m2, None, None, _lib.YC_HEAD_RAW, False, True)
        for i in range(nl):
            x[i + nl] = aux[i]

    def forward(self, x):
        self.training |= self.export
        nl = self.nl
        if self.training:
            _, raws = self._run(x[:nl], self.m, self.ia, self.im, _lib.YC_HEAD_RAW, False, True)
            self._aux(x)
            for i in range(nl):
                x[i] = raws[i]
            return x
        z, raws = self._run(x[:nl], self.m, self.ia, self.im, _lib.YC_HEAD_IDETECT, True, self.return_raw)
        if self.compute_aux_in_eval:
            self._aux(x)
        for i in range(nl):
            self._update_grid_cache(i, x[i].shape[2], x[i].shape[3], z.device)
            if self.return_raw:
                x[i] = raws[i]
        return z, x[:nl]
