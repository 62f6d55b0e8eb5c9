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


@pytest.mark.skipif(not ref_loader.extras_available(), reason="oracle/_ref consumers not built")
def test_consumers_random():
    """smoothing_library, void_library.gaussian_smoothing, Bk, Xi, correct_MAS of the compiled reference on a fresh
    field (odd-ish grid 22, sizes the golden files do not hold)."""
    ex = ref_loader.load_extras()
    MASL, PKL = ref_loader.load()
    rng = np.random.default_rng(123)
    dims, box = 22, 400.0
    pos = (rng.random((60000, 3)) * box).astype(np.float32)
    d = np.zeros((dims,) * 3, np.float32); MASL.MA(pos, d, box, "TSC")
    d /= d.mean(dtype=np.float64); d -= 1

    def close(a, b, what, rtol=1e-5):
        a, b = np.asarray(a), np.asarray(b)
        assert a.shape == b.shape and a.dtype == b.dtype, what
        assert np.abs(a.astype(np.complex128) - b).max() <= rtol * np.abs(b).max(), what

    for name, R in (("Top-Hat", 45.0), ("Gaussian", 30.0)):
        W_ref = np.asarray(ex["smoothing_library"].FT_filter(box, R, dims, name, 2))
        close(O.FT_filter(box, R, dims, name), W_ref, "FT_filter " + name)
        close(O.field_smoothing(d, W_ref), np.asarray(ex["smoothing_library"].field_smoothing(d, W_ref, 2)), "field_smoothing " + name)
    close(O.gaussian_smoothing(d, box, 37.0), np.asarray(ex["void_library"].gaussian_smoothing(d, box, 37.0, 2)), "gaussian_smoothing")
    close(O.correct_MAS(d, box, "TSC"), np.asarray(quiet(PKL.correct_MAS, d, box, "TSC", 1)), "correct_MAS")
    parity.check_xi(O.Xi(d, box, "TSC", 1), quiet(PKL.Xi, d, box, "TSC", 1, 1))
    kF = 2 * np.pi / box
    theta = np.array([0.4, 1.5, 2.6])
    r = quiet(ex["bispectrum_library"].Bk, d, box, 2.5 * kF, 3.5 * kF, theta, "TSC", 1)
    o = O.Bk(d, box, 2.5 * kF, 3.5 * kF, theta, "TSC")
    np.testing.assert_allclose(o.Pk, r.Pk, rtol=1e-5)
    np.testing.assert_allclose(o.B, r.B, rtol=1e-5, atol=1e-5 * np.abs(r.B).max())
    np.testing.assert_allclose(o.Q, r.Q, rtol=2e-5, atol=2e-5 * np.abs(r.Q).max())
