"""Mirror of the reference module ``utils`` (utils.pyx): Phylip readers and leaf encoding."""
import numpy as np  # noqa: F401

from .alignment import readBinaryPhy, readMultiPhy, readPhy, sites2Mat  # noqa: F401
