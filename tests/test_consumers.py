"""FFT consumers either side of the path: smoothing_library.FT_filter / field_smoothing,
void_library.gaussian_smoothing, PKL.Bk.  The oracle's restatements and the CUDA path are both checked against the
outputs of the compiled, unmodified reference (tests/golden/consumers.npz, made by tests/golden/make_golden.py)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import parity                                           # noqa: E402
from oracle import pylians_oracle as O                  # noqa: E402

# Filters and smoothed fields: float32 results of two float32 transforms; compare at 1e-5 of the largest value
# (single elements can sit arbitrarily close to zero).  Bk: ratios of sums over the whole grid, 1e-5 relative.
FIELD_RTOL = 1e-5
BK_RTOL = 1e-5


@pytest.fixture(scope="module")
def gc(golden_dir):
    return np.load(os.path.join(golden_dir, "consumers.npz"))


def _close_field(a, b, what, rtol=FIELD_RTOL):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (what, a.shape, b.shape, a.dtype, b.dtype)
    err = np.abs(a.astype(np.complex128) - b.astype(np.complex128)).max()
    assert err <= rtol * np.abs(b).max(), "%s: worst error %.3e vs scale %.3e" % (what, err, np.abs(b).max())


def _check_all(SL, VL, Bk, gc, to_in=lambda x: x):
    box, theta = float(gc["box"]), gc["theta"]
    kF = 2.0 * np.pi / box
    for dims in (16, 18):
        d = gc["delta_%d" % dims]
        for name, R in (("Top-Hat", 31.0), ("Gaussian", 17.5)):
            W_k = SL.FT_filter(box, R, dims, name, 2)
            _close_field(W_k, gc["filter_%d_%s" % (dims, name)], "FT_filter %s %d" % (name, dims))
            assert abs(W_k[0, 0, 0] - 1.0) < 1e-6                       # unit-sum filter
            sm = SL.field_smoothing(to_in(d), to_in(gc["filter_%d_%s" % (dims, name)]), 2)
            sm = sm.cpu().numpy() if hasattr(sm, "cpu") else sm
            _close_field(sm, gc["smooth_%d_%s" % (dims, name)], "field_smoothing %s %d" % (name, dims))
        vs = VL.gaussian_smoothing(to_in(d), box, 23.0, 2)
        vs = vs.cpu().numpy() if hasattr(vs, "cpu") else vs
        _close_field(vs, gc["void_smooth_%d" % dims], "gaussian_smoothing %d" % dims)
        b = Bk(to_in(d), box, 3.0 * kF, 4.2 * kF, theta, "CIC", 1)
        np.testing.assert_array_equal(b.k, gc["bk_%d_k" % dims])
        np.testing.assert_allclose(b.Pk, gc["bk_%d_Pk" % dims], rtol=BK_RTOL, atol=0)
        scale = np.abs(gc["bk_%d_B" % dims]).max()
        np.testing.assert_allclose(b.B, gc["bk_%d_B" % dims], rtol=BK_RTOL, atol=BK_RTOL * scale)
        qscale = np.abs(gc["bk_%d_Q" % dims]).max()
        np.testing.assert_allclose(b.Q, gc["bk_%d_Q" % dims], rtol=2 * BK_RTOL, atol=2 * BK_RTOL * qscale)


def test_oracle_consumers_match_reference_golden(gc):
    class _O(object):
        FT_filter = staticmethod(O.FT_filter)
        field_smoothing = staticmethod(O.field_smoothing)
        gaussian_smoothing = staticmethod(O.gaussian_smoothing)
    _check_all(_O, _O, O.Bk, gc)


def test_filter_errors_and_f2():
    from pylians_b200 import bispectrum_library as BL
    with pytest.raises(Exception, match="not implemented"):
        O.FT_filter(100.0, 5.0, 8, "Box", 1)
    # F2 kernel, known values: parallel vectors of equal length -> 5/7 + 1 + 2/7 = 2; perpendicular -> 5/7
    assert abs(BL.F2(np.array([0, 0, 1.0]), np.array([0, 0, 1.0])) - 2.0) < 1e-14
    assert abs(BL.F2(np.array([0, 0, 1.0]), np.array([0, 2.0, 0])) - 5.0 / 7.0) < 1e-14
    th, B = BL.Bispectrum_theory(np.logspace(-3, 1, 50), np.logspace(-3, 1, 50) ** -1.5, 0.1, 0.2)
    assert th.shape == B.shape == (50,) and np.all(np.isfinite(B[1:-1]))


@pytest.mark.gpu
def test_gpu_consumers_match_reference_golden(gc):
    import torch
    import pylians_b200
    import Pk_library as PKL
    import smoothing_library as SL
    import void_library as VL
    pylians_b200.set_verbose(False)
    _check_all(SL, VL, PKL.Bk, gc)                                       # numpy in -> numpy out
    _check_all(SL, VL, PKL.Bk, gc, to_in=lambda x: torch.from_numpy(x).cuda())     # CUDA tensors in -> CUDA tensors out
    with pytest.raises(Exception, match="not implemented"):
        SL.FT_filter(100.0, 5.0, 8, "Box", 1)
    with pytest.raises(Exception, match="different grids"):
        SL.field_smoothing(gc["delta_16"], gc["filter_18_Gaussian"], 1)


@pytest.mark.gpu
def test_gpu_consumers_against_oracle_at_64(gc):
    """A size the golden file does not hold: 64^3, TSC field, Gaussian filter and bispectrum vs the oracle."""
    import pylians_b200
    import MAS_library as MASL
    import Pk_library as PKL
    import smoothing_library as SL
    import void_library as VL
    pylians_b200.set_verbose(False)
    dims, box = 64, 500.0
    rng = np.random.default_rng(77)
    pos = (rng.random((3 * dims ** 3, 3)) * box).astype(np.float32)
    d = np.zeros((dims,) * 3, np.float32)
    MASL.MA(pos, d, box, "TSC")
    MASL.overdensity(d)
    for name, R in (("Top-Hat", 20.0), ("Gaussian", 12.0)):
        W_ref = O.FT_filter(box, R, dims, name)
        _close_field(SL.FT_filter(box, R, dims, name, 1), W_ref, "FT_filter 64 " + name)
        _close_field(SL.field_smoothing(d, W_ref, 1), O.field_smoothing(d, W_ref), "field_smoothing 64 " + name)
    _close_field(VL.gaussian_smoothing(d, box, 15.0), O.gaussian_smoothing(d, box, 15.0), "gaussian_smoothing 64")
    kF = 2 * np.pi / box
    theta = np.linspace(0.1, 3.0, 5)
    g, r = PKL.Bk(d, box, 5 * kF, 8 * kF, theta, "TSC", 1), O.Bk(d, box, 5 * kF, 8 * kF, theta, "TSC")
    np.testing.assert_allclose(g.Pk, r.Pk, rtol=BK_RTOL)
    np.testing.assert_allclose(g.B, r.B, rtol=BK_RTOL, atol=BK_RTOL * np.abs(r.B).max())
