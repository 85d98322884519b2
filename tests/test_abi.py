"""The C-ABI library loads and exports every symbol include/gpb200.h declares
(no compute calls: this runs without a GPU)."""
import os
import re

import pytest

from conftest import ROOT


def _declared():
    txt = open(os.path.join(ROOT, "include", "gpb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gpb_[a-zA-Z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from gaussian_processes_b200 import _lib
    names = _declared()
    assert len(names) >= 45
    for n in names:
        assert hasattr(_lib.lib, n), "libgpb200.so does not export %s" % n


def test_binding_covers_header():
    from gaussian_processes_b200 import _lib
    missing = set(_declared()) - set(_lib.EXPORTS)
    assert not missing, "ctypes prototypes missing for %s" % sorted(missing)


def test_constants():
    from gaussian_processes_b200 import _lib
    assert _lib.lib.gpb_version() >= 100
    assert abs(_lib.lib.gpb_min_log() - (-705.6238298100243)) < 1e-12


def test_header_cites_reference_lines():
    txt = open(os.path.join(ROOT, "include", "gpb200.h")).read()
    for cite in ("gaussian_c.pyx:18", "periodic_c.pyx:18", "gp_c.pyx:17-31", "gp/gp.py:294", "gp/gp.py:332-334"):
        assert cite in txt


def test_no_cpu_fallback():
    """Without a GPU every compute entry point must fail loudly, never fall back."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import gaussian_processes_b200 as gpb
    from gaussian_processes_b200._lib import GpbError
    k = gpb.GaussianKernel(1.0, 0.5)
    x = np.linspace(0, 1, 5)
    with pytest.raises(GpbError):
        k(x, x)
    gp = gpb.GP(k, x, np.sin(x), s=1.0)
    with pytest.raises(GpbError):
        gp.log_lh


def test_product_never_imports_oracle():
    """Only tests/, smoke() and bench.py's baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "gaussian_processes_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower().replace("# oracle", ""), "%s mentions the oracle" % f
