"""CPU: the oracle (oracle/pyl_oracle.c) against reference-generated golden vectors and analytic
known answers.  This is what pins the oracle on machines without
/root/reference."""
import os

import numpy as np
import pytest

from oracle import pylians_oracle as O
import parity


@pytest.fixture(scope="module")
def gma(golden_dir):
    return np.load(os.path.join(golden_dir, "ma.npz"))


@pytest.mark.parametrize("mas", ["NGP", "CIC", "TSC", "PCS"])
@pytest.mark.parametrize("weighted", [False, True])
def test_ma3d_golden(gma, mas, weighted):
    box, dims = float(gma["box"]), int(gma["dims"])
    g = np.zeros((dims,) * 3, np.float32)
    O.MA(gma["pos"], g, box, mas, W=gma["W"] if weighted else None)
    ref = gma["grid_%s%s" % (mas, "W" if weighted else "")]
    if mas == "NGP" and not weighted:
        parity.assert_exact(g, ref, "NGP grid")
    parity.assert_grid_close(g, ref, mas)


def test_ma_accumulates_in_place(gma):
    box, dims = float(gma["box"]), int(gma["dims"])
    g = np.full((dims,) * 3, 0.5, np.float32)
    O.MA(gma["pos"][:1000], g, box, "CIC")
    O.MA(gma["pos"][1000:2000], g, box, "TSC", W=gma["W"][1000:2000])
    parity.assert_grid_close(g, gma["grid_accum"], "accumulate")


@pytest.mark.parametrize("mas", ["NGP", "CIC", "TSC", "PCS"])
@pytest.mark.parametrize("weighted", [False, True])
def test_ma2d_golden(gma, mas, weighted):
    box, dims = float(gma["box"]), int(gma["dims2"])
    g = np.zeros((dims,) * 2, np.float32)
    O.MA(np.ascontiguousarray(gma["pos"][:, :2]), g, box, mas, W=gma["W"] if weighted else None)
    parity.assert_grid_close(g, gma["grid2d_%s%s" % (mas, "W" if weighted else "")], "2d " + mas)


def test_ma2d_norenorm(gma):
    box, dims = float(gma["box"]), int(gma["dims2"])
    g = np.zeros((dims,) * 2, np.float32)
    O.MA(np.ascontiguousarray(gma["pos"][:, :2]), g, box, "TSC", renormalize_2D=False)
    parity.assert_grid_close(g, gma["grid2d_TSC_norenorm"], "2d TSC no renorm")


def test_ma_fp64_grid(gma):
    box, dims = float(gma["box"]), int(gma["dims"])
    for mas, key in (("NGP", "grid_NGPW_d"), ("CIC", "grid_CICW_d")):
        g = np.zeros((dims,) * 3, np.float64)
        O.MA(gma["pos"], g, box, mas, W=gma["W"])
        np.testing.assert_allclose(g, gma[key], rtol=1e-6, atol=1e-6)


def test_ma_matches_openmp_entry(gma):
    box, dims = float(gma["box"]), int(gma["dims"])
    g = np.zeros((dims,) * 3, np.float32)
    O.MA(gma["pos"], g, box, "PCS", W=gma["W"])
    parity.assert_grid_close(g, gma["grid_PCSWc3D"], "PCSWc3D")


@pytest.mark.parametrize("mas", ["NGP", "CIC", "TSC", "PCS"])
@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("ndim", [2, 3])
def test_masc_shims_golden(gma, mas, weighted, ndim):
    """The sixteen MAS_c shims (Test/test_MAS.py:88-238).  Their 2-D kernels add every contribution once (MAS_c.c:21-34),
    which equals MA()'s renormalised 2-D grid up to the rounding of S equal fp32 additions."""
    box = float(gma["box"])
    dims = int(gma["dims"] if ndim == 3 else gma["dims2"])
    pos = gma["pos"] if ndim == 3 else np.ascontiguousarray(gma["pos"][:, :2])
    g = np.zeros((dims,) * ndim, np.float32)
    O.MA(pos, g, box, mas, W=gma["W"] if weighted else None)
    parity.assert_grid_close(g, gma["c%dD_%s%s" % (ndim, mas, "W" if weighted else "")], "%s c%dD" % (mas, ndim))


def test_ma_fortran_order_pos(gma):
    box, dims = float(gma["box"]), int(gma["dims"])
    g = np.zeros((dims,) * 3, np.float32)
    O.MA(np.asfortranarray(gma["pos"]), g, box, "CIC")
    parity.assert_grid_close(g, gma["grid_CIC"], "F-order pos")


def test_ma_errors():
    pos = np.zeros((4, 3), np.float32)
    with pytest.raises(SystemExit):
        O.MA(pos, np.zeros((8, 8), np.float32), 1.0, "CIC")
    with pytest.raises(SystemExit):
        O.MA(pos, np.zeros((8, 8, 8), np.float32), 1.0, "XYZ")
    with pytest.raises(ValueError):
        O.MA(pos.astype(np.float64), np.zeros((8, 8, 8), np.float32), 1.0, "CIC")


# reference's own tests: Test/test_MAS.py:10-84 (mass conservation, 1000 particles, 64^3, unit box)
@pytest.mark.parametrize("mas,places", [("NGP", 20), ("CIC", 8), ("TSC", 8), ("PCS", 8)])
@pytest.mark.parametrize("weighted", [False, True])
def test_reference_unit_tests_mass_conservation(mas, places, weighted):
    particles, BoxSize, dims, seed = 1000, 1.0, 64, 1
    np.random.seed(seed)
    pos = np.random.random((particles, 3)).astype(np.float32)
    delta = np.zeros((dims, dims, dims), dtype=np.float32)
    W = np.ones(particles, dtype=np.float32) * 3.0 if weighted else None
    O.MA(pos, delta, BoxSize, mas, W=W)
    suma = np.sum(delta, dtype=np.float64)
    norm = 3.0 * particles if weighted else particles
    assert round(abs(suma / norm - 1.0), places) == 0


def test_single_particle_profiles():
    # a particle exactly on a grid point: TSC (1/8,3/4,1/8), PCS (1/6,2/3,1/6) per axis
    dims, box = 8, 8.0
    pos = np.array([[3.0, 3.0, 3.0]], np.float32)
    g = np.zeros((dims,) * 3, np.float32); O.MA(pos, g, box, "TSC")
    np.testing.assert_allclose(g[2:5, 3, 3], [0.125 * 0.75 * 0.75, 0.75 ** 3, 0.125 * 0.75 * 0.75], rtol=1e-6)
    g = np.zeros((dims,) * 3, np.float32); O.MA(pos, g, box, "PCS")
    np.testing.assert_allclose(g[2:5, 3, 3], [(1 / 6) * (2 / 3) ** 2, (2 / 3) ** 3, (1 / 6) * (2 / 3) ** 2], rtol=1e-6)
    assert g[5, 3, 3] == 0.0
    g = np.zeros((dims,) * 3, np.float32); O.MA(pos, g, box, "CIC")
    assert g[3, 3, 3] == 1.0 and g.sum() == 1.0


# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def gpk(golden_dir):
    return np.load(os.path.join(golden_dir, "pk.npz"))


@pytest.mark.parametrize("dims", [16, 20])
@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("mas", ["TSC", "None"])
def test_pk_golden(gpk, dims, axis, mas):
    p = O.Pk(gpk["delta_%d" % dims], float(gpk["box"]), axis, mas, 1)
    ref = {n: gpk["pk_%d_a%d_%s_%s" % (dims, axis, mas, n)] for n in
           ["k3D", "Pk", "Nmodes3D", "Pkphase", "k1D", "Pk1D", "Nmodes1D", "kpar", "kper", "Pk2D", "Nmodes2D"]}
    parity.check_pk(p, ref)


def test_pk_keep_deltak(gpk):
    p = O.Pk(gpk["delta_16"], float(gpk["box"]), 2, "TSC", 1, keep_deltak=True)
    ref = gpk["pk_16_deltak"]
    np.testing.assert_allclose(p.delta_k, ref, rtol=1e-4, atol=1e-4 * np.abs(ref).max())


def test_xpk_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "xpk.npz"))
    names = ["k3D", "Pk", "XPk", "Nmodes3D", "k1D", "Pk1D", "PkX1D", "Nmodes1D", "kpar", "kper", "Pk2D", "PkX2D", "Nmodes2D"]
    fs = [g["delta%d" % i] for i in range(3)]
    x = O.XPk(fs[:2], float(g["box"]), 2, ["CIC", "PCS"], 1)
    parity.check_xpk(x, {n: g["x2_a2_" + n] for n in names})
    x = O.XPk(fs, float(g["box"]), 0, ["CIC", "PCS", "None"], 1)
    parity.check_xpk(x, {n: g["x3_a0_" + n] for n in names})


def test_xpk_auto_equals_pk(gpk):
    d = gpk["delta_16"]
    p = O.Pk(d, 1000.0, 2, "TSC", 1)
    x = O.XPk([d, d], 1000.0, 2, ["TSC", "TSC"], 1)
    assert np.array_equal(x.Pk[:, :, 0], p.Pk)
    np.testing.assert_allclose(x.XPk[:, 0, 0], p.Pk[:, 0], rtol=1e-14)


def test_mode_count_invariant():
    # Pk_library.pyx:90-102: sum(Nmodes3D incl. DC) == (N^3-8)/2+8 for even N
    for dims in (8, 12, 16):
        d = np.random.default_rng(0).standard_normal((dims,) * 3).astype(np.float32)
        p = O.Pk(d, 100.0, 2, "None", 1)
        assert p.Nmodes3D.sum() + 1 == (dims ** 3 - 8) // 2 + 8
        kF = 2 * np.pi / 100.0
        i = np.arange(len(p.k3D))
        assert np.all(p.k3D >= (i + 1) * kF - 1e-12) and np.all(p.k3D < (i + 2) * kF)


def test_plane_wave_known_answer():
    # delta = A cos(2 pi 5 z / N): one independent mode at kz=5
    N, L, A = 32, 1000.0, 2.0
    z = np.arange(N)
    d = np.broadcast_to(A * np.cos(2 * np.pi * 5 * z / N), (N, N, N)).astype(np.float32).copy()
    p = O.Pk(d, L, 2, "None", 1)
    b = 4  # k_index 5 -> bin 4 after dropping DC
    want = (A * N ** 3 / 2) ** 2 * (L / N ** 2) ** 3
    np.testing.assert_allclose(p.Pk[b, 0] * p.Nmodes3D[b], want, rtol=1e-5)
    np.testing.assert_allclose(p.Pk[b, 1] / p.Pk[b, 0], 5.0, rtol=1e-5)
    np.testing.assert_allclose(p.Pk[b, 2] / p.Pk[b, 0], 9.0, rtol=1e-5)
    p = O.Pk(d, L, 0, "None", 1)
    np.testing.assert_allclose(p.Pk[b, 1] / p.Pk[b, 0], -2.5, rtol=1e-5)
    np.testing.assert_allclose(p.Pk[b, 2] / p.Pk[b, 0], 3.375, rtol=1e-5)


def test_white_noise_level():
    N, L = 32, 1000.0
    d = np.random.default_rng(2).standard_normal((N, N, N)).astype(np.float32)
    p = O.Pk(d, L, 2, "None", 1)
    w = p.Nmodes3D
    assert abs(np.sum(p.Pk[:, 0] * w) / np.sum(w) / (L ** 3 / N ** 3) - 1.0) < 0.02


def test_rsd_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "rsd.npz"))
    for axis in (0, 1, 2):
        a = g["pos"].copy()
        O.pos_redshift_space(a, g["vel"], float(g["box"]), float(g["hubble"]), float(g["redshift"]), axis)
        assert np.array_equal(a, g["rsd_a%d" % axis])
