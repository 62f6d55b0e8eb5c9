"""CPU: the C-ABI library loads and exports every symbol include/pylians_b200.h declares.
No compute call is made (there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "pylians_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:int|void|size_t|int64_t|char)\s*\*?\s*(\w+)\s*\(", src, flags=re.M)
    return sorted(set(names))


def test_header_declares_expected_entry_points():
    names = header_symbols()
    for must in ("pylb_ma", "pylb_pk_bin", "pylb_fft_r2c", "NGP", "CIC", "TSC", "PCS", "pylb_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from pylians_b200 import _lib
    if not os.path.exists(_lib.SO_PATH):
        from pylians_b200 import build
        build.build()
    lib = ctypes.CDLL(_lib.SO_PATH)
    for name in header_symbols():
        assert hasattr(lib, name), "library does not export %s" % name
    # and the python binding table covers the header exactly
    assert sorted(_lib.SIGNATURES) == header_symbols()


def test_layout_matches_reference_frequencies():
    """pylb_pk_get_layout is host-only code: bin counts must equal frequencies() (Pk_library.pyx:59-64)."""
    from pylians_b200 import Pk_library as PKL
    from oracle import pylians_oracle as O
    for dims in (8, 15, 16, 20, 64, 128, 512, 1024, 2048):
        L = PKL.get_layout(dims, 2)
        kF, kN, kpar, kper, kmax = O.frequencies(1000.0, dims)
        assert (L.kmax_par, L.kmax_per, L.kmax) == (kpar, kper, kmax)
        assert L.B2 == (kpar + 1) * (kper + 1) and L.X == 1
        assert PKL.frequencies(1000.0, dims) == (kF, kN, kpar, kper, kmax)


def test_no_cpu_fallback():
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import MAS_library as MASL
    import Pk_library as PKL
    with pytest.raises(RuntimeError):
        MASL.MA(np.zeros((4, 3), np.float32), np.zeros((8, 8, 8), np.float32), 1.0, "CIC")
    with pytest.raises(RuntimeError):
        PKL.Pk(np.zeros((8, 8, 8), np.float32), 1.0, 2, "CIC", 1)


def test_argument_errors_match_reference_before_any_gpu_work():
    import numpy as np
    import MAS_library as MASL
    pos = np.zeros((4, 3), np.float32)
    with pytest.raises(SystemExit):
        MASL.MA(pos, np.zeros((8, 8), np.float32), 1.0, "CIC")          # dimension mismatch
    with pytest.raises(SystemExit):
        MASL.MA(pos, np.zeros((8, 8, 8), np.float32), 1.0, "XYZ")       # bad scheme
    with pytest.raises(ValueError):
        MASL.MA(pos.astype(np.float64), np.zeros((8, 8, 8), np.float32), 1.0, "CIC")
