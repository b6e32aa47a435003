"""IBin head (drop-in for reference nets/ibin.py) on the B200 kernels."""
from .. import _lib
from ..losses.sigmoid_bin import SigmoidBin
from ._head import HeadBase


class IBin(HeadBase):
    """Bin-regression head: per anchor [x, y | w: 1 reg + 21 bins | h: 1 reg + 21 bins | obj | cls].

    Reference: nets/ibin.py:12-74; width/height decode per losses/sigmoid_bin.py:49-63.
    Eval output z has nc+5 columns, the raw maps keep all `no` columns.
    """

    def __init__(self, nc=80, anchors=(), ch=(), bin_count=21):
        super().__init__()
        self.bin_count = bin_count
        self.w_bin_sigmoid = SigmoidBin(bin_count=bin_count, min=0.0, max=4.0)
        self.h_bin_sigmoid = SigmoidBin(bin_count=bin_count, min=0.0, max=4.0)
        no = nc + 3 + self.w_bin_sigmoid.get_length() + self.h_bin_sigmoid.get_length()
        self._init_common(nc, anchors, no)
        self._make_lead(ch)

    def forward(self, x):
        self.w_bin_sigmoid.use_fw_regression = True
        self.h_bin_sigmoid.use_fw_regression = True
        self.training |= self.export
        nl = self.nl
        if self.training:
            _, raws = self._run(x[:nl], self.m, self.ia, self.im, _lib.YC_HEAD_RAW, False, True)
            for i in range(nl):
                x[i] = raws[i]
            return x
        # the exact-fp32 path decodes from the raw maps, so it always needs them; the tcgen05 path (bf16 maps whose
        # levels meet the TMA alignment rules) decodes in its epilogue
        import torch
        tc = all(t.dtype == torch.bfloat16 and (t.shape[2] * t.shape[3]) % 8 == 0 and t.shape[1] % 8 == 0
                 and t.data_ptr() % 16 == 0 for t in x[:nl])
        want_raw = self.return_raw or not tc or self.head_path == _lib.YC_PATH_GENERIC
        z, raws = self._run(x[:nl], self.m, self.ia, self.im, _lib.YC_HEAD_IBIN, True, want_raw,
                            no_out=self.nc + 5, bins=self.w_bin_sigmoid.bins, bin_count=self.bin_count)
        for i in range(nl):
            self._update_grid_cache(i, x[i].shape[2], x[i].shape[3], z.device)
            if self.return_raw:     # return_raw = False skips the 127-column raw maps (60 % of the output traffic)
                x[i] = raws[i]
        return z, x
