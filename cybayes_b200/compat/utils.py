"""Top-level alias so the reference's unchanged driver scripts (`import utils`,
`from utils import *`) resolve to cybayes_b200.utils when this directory is on PYTHONPATH."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.modules[__name__] = importlib.import_module("cybayes_b200.utils")
