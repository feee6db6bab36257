"""Mirror of the reference module ``ML_gamma`` (ML_gamma.pyx): matML, matML_cython, cache_matML,
plus the batched ``score_proposals`` extension."""
import numpy as np  # noqa: F401

from . import config  # noqa: F401
from .likelihood import cache_matML, matML, matML_cython, score_proposals  # noqa: F401
