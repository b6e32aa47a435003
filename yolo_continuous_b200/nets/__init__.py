from .common import ImplicitA, ImplicitM  # noqa: F401
from .detect import Detect  # noqa: F401
from .iaux_detect import IAuxDetect  # noqa: F401
from .ibin import IBin  # noqa: F401
from .idetect import IDetect  # noqa: F401
