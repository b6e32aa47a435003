"""B200-native detection post-backbone path for xin-pu/yolo-continuous.

Drop-in replacements (same names, signatures and state_dict keys as the reference) for
nets/idetect.py, nets/iaux_detect.py, nets/ibin.py, utils/bbox.py and the decode / NMS
functions of detect.py, backed by hand-written sm_100a kernels behind a C ABI
(include/yc_b200.h, csrc/libyc_b200.so).  There is no CPU or eager fallback: importing
this package without the built library raises.
"""
from . import _lib  # noqa: F401  (fails loudly when csrc/libyc_b200.so is missing)

__all__ = ["_lib"]
