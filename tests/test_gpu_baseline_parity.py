"""GPU: parity at BASELINE.json sizes -- the CUDA path against the UNMODIFIED reference compiled from its own
sources (oracle/_ref; the C oracle where that was never built) on the SAME host-resident particles.

    particles -> MASL.MA -> delta = rho/mean - 1 -> PKL.Pk        256^3 (every scheme, uniform and Zel'dovich)
                                                                  512^3 (configs[1]: CIC uniform; PCS weighted Zel'dovich)

Contract (BASELINE.json north_star): Nmodes and k-bin edges bit-exact, density grid and P(k) within 1e-5 relative
(fp32; atomic order is not fixed).  Three comparisons per case: the grids; Pk of ONE field (the GPU's delta) through
both implementations, which isolates FFT + binning; and the whole chain end to end.
"""
import contextlib
import io
import os

import numpy as np
import pytest

import parity

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

BOX = 1000.0


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pylians_b200
    pylians_b200.set_verbose(False)


def checker():
    """(kind, deposit(pos, dims, mas, W) -> grid, Pk(delta, mas, axis) -> object) of the strongest checker present."""
    from oracle import ref_loader
    threads = os.cpu_count() or 1
    if ref_loader.available():
        RM, RP = ref_loader.load()

        def dep(pos, dims, mas, W, fast):
            d = np.zeros((dims,) * 3, np.float32)
            if fast:            # the reference's OpenMP C kernels (MAS_c.c via MAS_library.pyx:1136-1220), all host threads
                if W is None:
                    getattr(RM, mas + "c3D")(pos, d, BOX, threads)
                else:
                    getattr(RM, mas + "Wc3D")(pos, d, W, BOX, threads)
            else:               # the serial Cython kernels MA() dispatches to (MAS_library.pyx:57-112)
                RM.MA(pos, d, BOX, mas, W=W)
            return d

        def pk(delta, mas, axis):
            with contextlib.redirect_stdout(io.StringIO()):
                return RP.Pk(delta, BOX, axis, mas, threads)
        return "reference", dep, pk
    from oracle import pylians_oracle as O

    def dep(pos, dims, mas, W, fast):
        d = np.zeros((dims,) * 3, np.float32)
        O.MA(pos, d, BOX, mas, W=W)
        return d

    def pk(delta, mas, axis):
        return O.Pk(delta, BOX, axis, mas, threads)
    return "port", dep, pk


def particles(nside, data, seed):
    """Host float32 (nside^3, 3): uniform random, or a Zel'dovich-displaced lattice in lattice order (bench.py)."""
    if data == "uniform":
        rng = np.random.default_rng(seed)
        return (rng.random((nside ** 3, 3), dtype=np.float32) * np.float32(BOX)).astype(np.float32)
    import bench
    gen = torch.Generator(device="cuda"); gen.manual_seed(seed)
    pos = bench.zeldovich_particles(nside, BOX, gen, torch.device("cuda")).cpu().numpy()
    torch.cuda.empty_cache()
    return pos


def run_case(nside, mas, weighted, data, axis, fast_ref):
    import MAS_library as MASL
    import Pk_library as PKL
    kind, ref_dep, ref_pk = checker()
    pos = particles(nside, data, 7 + nside)
    W = None
    if weighted:
        W = (np.random.default_rng(5).random(nside ** 3, dtype=np.float32) + np.float32(0.5)).astype(np.float32)
    # ---- deposit ------------------------------------------------------------------------------------------
    got = np.zeros((nside,) * 3, np.float32)
    MASL.MA(pos, got, BOX, mas, W=W)                       # host arrays in, the public call of the reference
    ref = ref_dep(pos, nside, mas, W, fast_ref)
    if mas == "NGP" and not weighted:
        parity.assert_exact(got, ref, "NGP grid %d^3 %s" % (nside, data))
    parity.assert_grid_close(got, ref, "%s grid %d^3 %s vs %s" % (mas, nside, data, kind))
    # ---- overdensity (Pk_snapshot.py:88,194) ---------------------------------------------------------------
    ref /= np.mean(ref, dtype=np.float64); ref -= 1.0
    MASL.overdensity(got)
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-5 * (1.0 + np.abs(ref).max()))
    # ---- Pk of one and the same field through both implementations -------------------------------------
    mine = PKL.Pk(got, BOX, axis, mas, 1)
    few = 1e-4 if mas in ("TSC", "PCS") else None          # Nyquist-corner 2-D bins of 1-4 modes: see parity.check_pk
    parity.check_pk(mine, ref_pk(got, mas, axis), few_mode_rtol=few)
    # ---- the whole chain: the reference's spectrum of the reference's own field ---------------------------
    parity.check_pk(mine, ref_pk(ref, mas, axis), few_mode_rtol=few, corner_rtol=few)


@pytest.mark.parametrize("data", ["uniform", "zeldovich"])
@pytest.mark.parametrize("mas,weighted,axis", [("NGP", False, 2), ("CIC", False, 2), ("TSC", True, 2), ("PCS", False, 0)])
def test_parity_256(mas, weighted, axis, data):
    run_case(256, mas, weighted, data, axis, fast_ref=False)


@pytest.mark.parametrize("mas,weighted,data", [("CIC", False, "uniform"), ("PCS", True, "zeldovich")])
def test_parity_512(mas, weighted, data):
    """configs[1] (512^3 particles, CIC, 512^3 grid, Pk) exactly, and the PCS deposit of configs[3] at 512^3."""
    run_case(512, mas, weighted, data, 2, fast_ref=True)
