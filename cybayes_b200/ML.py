"""Mirror of the reference module ``ML`` (ML.pyx): single-rate matML / cache_matML taking the state dict."""
import numpy as np  # noqa: F401

from . import config  # noqa: F401
from .likelihood import cache_matML_single as cache_matML  # noqa: F401
from .likelihood import matML_single as matML  # noqa: F401
