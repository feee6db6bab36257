"""CPU: the C-ABI shared library loads, exports exactly what include/cybayes_b200.h declares, and
the product path fails loudly without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import REPO

HEADER = os.path.join(REPO, "include", "cybayes_b200.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from cybayes_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes binding and header disagree"
    assert lib.cb_version() >= 100


def test_binary_targets_sm_100a_only():
    so = os.path.join(REPO, "cybayes_b200", "csrc", "libcybayes_b200.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_no_gpu_means_loud_failure_not_fallback():
    from cybayes_b200 import _lib
    from cybayes_b200.engine import Engine
    lib = _lib.load()
    n = ctypes.c_int(0)
    if lib.cb_device_count(ctypes.byref(n)) == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.CyBayesB200Error, match="no CUDA device|CPU fallback"):
        Engine(np.zeros((3, 8), dtype=np.uint8), 2, 4)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, "cybayes_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "pruning_oracle" not in src and "golden_io" not in src and "oracle/" not in src.replace(
                    "oracle/_ref", "").replace("oracle/build_ref", ""), f
