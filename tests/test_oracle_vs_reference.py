"""CPU: the oracle against the compiled, unmodified reference (oracle/_ref) on fresh seeded inputs.
Skipped where oracle/_ref was never built (it is built in the container that has /root/reference
and shipped to the GPU box as a binary)."""
import contextlib
import io

import numpy as np
import pytest

from oracle import pylians_oracle as O, ref_loader
import parity

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="oracle/_ref not built")


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load()


@pytest.mark.parametrize("mas", ["NGP", "CIC", "TSC", "PCS"])
def test_ma_random(ref, mas):
    MASL, _ = ref
    rng = np.random.default_rng(42)
    dims, box = 40, 250.0
    pos = (rng.random((30000, 3)) * box).astype(np.float32)
    W = rng.random(30000).astype(np.float32)
    for w in (None, W):
        a = np.zeros((dims,) * 3, np.float32); b = np.zeros((dims,) * 3, np.float32)
        MASL.MA(pos, a, box, mas, W=w); O.MA(pos, b, box, mas, W=w)
        parity.assert_grid_close(b, a, mas)


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_pk_random(ref, axis):
    MASL, PKL = ref
    rng = np.random.default_rng(axis)
    dims, box = 24, 500.0
    pos = (rng.random((50000, 3)) * box).astype(np.float32)
    d = np.zeros((dims,) * 3, np.float32); MASL.MA(pos, d, box, "PCS")
    d /= d.mean(dtype=np.float64); d -= 1
    parity.check_pk(O.Pk(d, box, axis, "PCS", 2), quiet(PKL.Pk, d, box, axis, "PCS", 2))


def test_xpk_random(ref):
    MASL, PKL = ref
    rng = np.random.default_rng(9)
    dims, box = 18, 300.0
    fs = []
    for mas in ("CIC", "TSC"):
        pos = (rng.random((20000, 3)) * box).astype(np.float32)
        d = np.zeros((dims,) * 3, np.float32); MASL.MA(pos, d, box, mas)
        d /= d.mean(dtype=np.float64); d -= 1; fs.append(d)
    parity.check_xpk(O.XPk(fs, box, 1, ["CIC", "TSC"], 1), quiet(PKL.XPk, fs, box, 1, ["CIC", "TSC"], 1))
