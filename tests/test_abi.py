"""CPU tests (no GPU): the C-ABI libraries load, export every symbol include/*.h declares, and refuse to run
without a device instead of falling back to the CPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADERS = {"b200gso.h": "libb200gso.so", "b200enum.h": "libb200enum.so", "b200bkz.h": "libb200bkz.so",
           "b200hh.h": "libb200hh.so"}


@pytest.mark.parametrize("hdr,libname", sorted(HEADERS.items()))
def test_exports_every_declared_symbol(hdr, libname):
    hp = os.path.join(ROOT, "include", hdr)
    lp = os.path.join(ROOT, "fplll_b200", "lib", libname)
    if not os.path.exists(hp):
        pytest.skip(hdr + " not part of this build yet")
    assert os.path.exists(lp), "%s missing: build() did not produce it" % lp
    lib = C.CDLL(lp)
    prefix = hdr.split(".")[0]
    names = set(re.findall(r"\b(%s_[a-z_0-9]+)\s*\(" % prefix, open(hp).read()))
    assert names, "no declarations found"
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, "declared in %s but not exported: %s" % (hdr, missing)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np
    import fplll_b200 as fb
    with pytest.raises(fb.B200Error) as e:
        fb.MatGSO(np.eye(4, dtype=np.int64))
    assert "no CUDA device" in str(e.value) or "-2" in str(e.value)


def test_product_does_not_import_oracle():
    """The shipped package must never reach into oracle/ (the checker)."""
    pkg = os.path.join(ROOT, "fplll_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b|oracle/|liboracle", txt, re.M), os.path.join(dp, f)
