"""Siblings of Pk/XPk that share the FFT and the mode loop: Pk_plane, XPk_plane, Pk_theta,
correct_MAS, Xi.  CPU: the numpy oracle against the reference's golden vectors.  GPU: the CUDA path against the
golden vectors and against the oracle on other seeded inputs."""
import os

import numpy as np
import pytest

from oracle import pylians_oracle as O
import parity

GRIDS = (16, 18)


@pytest.fixture(scope="module")
def gs(golden_dir):
    return np.load(os.path.join(golden_dir, "siblings.npz"))


def _gold(gs, prefix, names):
    return {n: gs["%s_%s" % (prefix, n)] for n in names}


def _run_all(impl, gs, grid):
    """Every sibling on the golden inputs with implementation module `impl`; returns dict of results."""
    box = float(gs["box"])
    a, b = gs["img_a_%d" % grid], gs["img_b_%d" % grid]
    v = [gs["vel%d_%d" % (i, grid)] for i in range(3)]
    d = gs["delta_%d" % grid]
    return {"plane": impl.Pk_plane(a, box, "CIC", 1),
            "xplane": impl.XPk_plane(a, b, box, "CIC", "TSC", 1),
            "theta": impl.Pk_theta(v[0], v[1], v[2], box, 2, "PCS", 1),
            "correct": np.asarray(impl.correct_MAS(d, box, "CIC", 1)),
            "xi": [impl.Xi(d, box, "CIC", axis, 1) for axis in (0, 1, 2)]}


def _check_against_golden(res, gs, grid):
    parity.check_plane(res["plane"], _gold(gs, "plane_%d" % grid, ("k", "Nmodes", "Pk")))
    parity.check_xplane(res["xplane"], _gold(gs, "xplane_%d" % grid, ("k", "Nmodes", "Pk", "XPk", "r")))
    parity.check_theta(res["theta"], [gs["theta_%d_%s" % (grid, n)] for n in ("k", "Pk", "Nmodes")])
    ref = gs["correct_%d" % grid]
    np.testing.assert_allclose(res["correct"], ref, rtol=0, atol=1e-5 * float(np.abs(ref).max()))
    for axis in (0, 1, 2):
        parity.check_xi(res["xi"][axis], _gold(gs, "xi_%d_a%d" % (grid, axis), ("r3D", "xi", "Nmodes3D")))


@pytest.mark.parametrize("grid", GRIDS)
def test_oracle_matches_reference_golden(gs, grid):
    _check_against_golden(_run_all(O, gs, grid), gs, grid)


@pytest.mark.gpu
@pytest.mark.parametrize("grid", GRIDS)
def test_gpu_matches_reference_golden(gs, grid):
    import pylians_b200
    import Pk_library as PKL
    pylians_b200.set_verbose(False)
    _check_against_golden(_run_all(PKL, gs, grid), gs, grid)


@pytest.mark.gpu
@pytest.mark.parametrize("grid", [24, 33, 64])
def test_gpu_matches_oracle(grid):
    """Fresh seeded inputs, including an odd grid (no Nyquist planes) and device-tensor inputs."""
    import torch
    import pylians_b200
    import Pk_library as PKL
    pylians_b200.set_verbose(False)
    rng = np.random.default_rng(grid)
    box = 250.0
    a = rng.standard_normal((grid, grid)).astype(np.float32)
    b = (a[::-1] + 0.3 * rng.standard_normal((grid, grid))).astype(np.float32)
    parity.check_plane(PKL.Pk_plane(a, box, "PCS", 1), O.Pk_plane(a, box, "PCS"))
    parity.check_xplane(PKL.XPk_plane(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), box, "NGP", "None", 1),
                        O.XPk_plane(a, b, box, "NGP", "None"))
    v = [rng.standard_normal((grid,) * 3).astype(np.float32) for _ in range(3)]
    parity.check_theta(PKL.Pk_theta(v[0], v[1], v[2], box, 2, "CIC", 1), O.Pk_theta(v[0], v[1], v[2], box, 2, "CIC"))
    d = (rng.standard_normal((grid,) * 3) * 0.3).astype(np.float32)
    ref = O.correct_MAS(d, box, "TSC")
    got = PKL.correct_MAS(torch.from_numpy(d).cuda(), box, "TSC", 1)
    assert isinstance(got, torch.Tensor) and got.is_cuda
    np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=0, atol=1e-5 * float(np.abs(ref).max()))
    for axis in (0, 2):
        parity.check_xi(PKL.Xi(d, box, "TSC", axis, 1), O.Xi(d, box, "TSC", axis))
    # the helper transforms
    np.testing.assert_allclose(PKL.FFT2Dr_f(a, 1), np.fft.rfftn(a.astype(np.float64)), rtol=0, atol=2e-5 * grid)
    dk = np.fft.rfftn(d.astype(np.float64)).astype(np.complex64)
    np.testing.assert_allclose(PKL.IFFT3Dr_f(dk, 1), d, rtol=0, atol=1e-5)       # pyfftw normalises the inverse
    ak = np.fft.rfftn(a.astype(np.float64)).astype(np.complex64)
    np.testing.assert_allclose(PKL.IFFT2Dr_f(ak, 1), a, rtol=0, atol=1e-5)


@pytest.mark.gpu
def test_xi_is_the_transform_of_pk():
    """Known answer: a single plane wave delta = A cos(2 pi m x / L) has xi(r) = (A^2/2) cos(2 pi m r_x / L)
    (MAS 'None').  The reference divides the inverse transform of |delta_k|^2 (delta_k from an unnormalised forward
    transform) by dims^3 once (:2139-2143); together with the 1/dims^3 pyfftw applies to every inverse transform
    (FFTW.__call__, normalise_idft=True) that is exactly xi.  correct_MAS with 'None' is the identity."""
    import pylians_b200
    import Pk_library as PKL
    pylians_b200.set_verbose(False)
    grid, box, m, A = 32, 100.0, 3, 0.25
    x = np.arange(grid)
    d = np.broadcast_to((A * np.cos(2 * np.pi * m * x / grid))[:, None, None], (grid,) * 3).astype(np.float32).copy()
    same = PKL.correct_MAS(d, box, "None", 1)
    np.testing.assert_allclose(same, d, rtol=0, atol=1e-6)
    xi = PKL.Xi(d, box, "None", 0, 1)
    # bin 1 (r in [1,2) cells): mean over its 26 cells of (A^2/2) cos(2 pi m r_x / grid)
    w = np.where(x > grid // 2, x - grid, x)
    rx, ry, rz = np.meshgrid(w, w, w, indexing="ij")
    r = np.sqrt(rx * rx + ry * ry + rz * rz)
    sel = (r >= 1) & (r < 2)
    want = np.mean((A * A / 2) * np.cos(2 * np.pi * m * rx[sel] / grid))
    assert abs(xi.xi[0, 0] - want) < 1e-6
    assert xi.Nmodes3D[0] == sel.sum()


def _check_ximag(impl, gs, rtol=parity.PK_RTOL):
    box = float(gs["box"])
    fs = [gs["ximag_delta%d" % i] for i in range(3)]
    for axis in (2, 0):
        ref = {n: gs["ximag_a%d_%s" % (axis, n)] for n in
               ("k3D", "Pk", "XPk", "Nmodes3D", "k1D", "Pk1D", "PkX1D", "Nmodes1D", "kpar", "kper", "Pk2D", "PkX2D", "Nmodes2D")}
        parity.check_xpk(impl.XPk_imag(fs, box, axis, ["CIC", "TSC", "PCS"], 1), ref, rtol)


def test_oracle_ximag_matches_reference_golden(gs):
    _check_ximag(O, gs)


@pytest.mark.gpu
def test_gpu_ximag_matches_reference_golden(gs):
    import pylians_b200
    import Pk_library as PKL
    pylians_b200.set_verbose(False)
    _check_ximag(PKL, gs)
