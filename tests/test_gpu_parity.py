"""GPU: the CUDA path, called through the reference-shaped Python API (which goes through the C ABI),
against the CPU oracle and the reference-generated golden vectors."""
import ctypes
import os

import numpy as np
import pytest

import parity

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import pylians_oracle as O  # noqa: E402  (checker only)


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pylians_b200
    pylians_b200.set_verbose(False)


@pytest.fixture(scope="module")
def MASL():
    import MAS_library
    return MAS_library


@pytest.fixture(scope="module")
def PKL():
    import Pk_library
    return Pk_library


@pytest.fixture(scope="module")
def gma(golden_dir):
    return np.load(os.path.join(golden_dir, "ma.npz"))


# ------------------------------------------------------------------------------------------------
# deposit
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mas", ["NGP", "CIC", "TSC", "PCS"])
@pytest.mark.parametrize("weighted", [False, True])
def test_ma3d_golden(MASL, gma, mas, weighted):
    box, dims = float(gma["box"]), int(gma["dims"])
    g = np.zeros((dims,) * 3, np.float32)
    MASL.MA(gma["pos"], g, box, mas, W=gma["W"] if weighted else None)
    ref = gma["grid_%s%s" % (mas, "W" if weighted else "")]
    if mas == "NGP" and not weighted:
        parity.assert_exact(g, ref, "NGP grid")
    parity.assert_grid_close(g, ref, mas)


def test_ma_accumulates_in_place(MASL, gma):
    box, dims = float(gma["box"]), int(gma["dims"])
    g = np.full((dims,) * 3, 0.5, np.float32)
    MASL.MA(gma["pos"][:1000], g, box, "CIC")
    MASL.MA(gma["pos"][1000:2000], g, box, "TSC", W=gma["W"][1000:2000])
    parity.assert_grid_close(g, gma["grid_accum"], "accumulate")


@pytest.mark.parametrize("mas", ["NGP", "CIC", "TSC", "PCS"])
@pytest.mark.parametrize("weighted", [False, True])
def test_ma2d_golden(MASL, gma, mas, weighted):
    box, dims = float(gma["box"]), int(gma["dims2"])
    g = np.zeros((dims,) * 2, np.float32)
    MASL.MA(np.ascontiguousarray(gma["pos"][:, :2]), g, box, mas, W=gma["W"] if weighted else None)
    parity.assert_grid_close(g, gma["grid2d_%s%s" % (mas, "W" if weighted else "")], "2d " + mas)


def test_ma2d_norenorm(MASL, gma):
    box, dims = float(gma["box"]), int(gma["dims2"])
    g = np.zeros((dims,) * 2, np.float32)
    MASL.MA(np.ascontiguousarray(gma["pos"][:, :2]), g, box, "TSC", renormalize_2D=False)
    parity.assert_grid_close(g, gma["grid2d_TSC_norenorm"], "2d TSC no renorm")


def test_ma_fp64_grid(MASL, gma):
    box, dims = float(gma["box"]), int(gma["dims"])
    g = np.zeros((dims,) * 3, np.float64); MASL.NGPW_d(gma["pos"], g, box, gma["W"])
    np.testing.assert_allclose(g, gma["grid_NGPW_d"], rtol=1e-6, atol=1e-6)
    g = np.zeros((dims,) * 3, np.float64); MASL.CICW_d(gma["pos"], g, box, gma["W"])
    np.testing.assert_allclose(g, gma["grid_CICW_d"], rtol=1e-6, atol=1e-6)


def test_ma_openmp_shims(MASL, gma):
    box, dims = float(gma["box"]), int(gma["dims"])
    g = np.zeros((dims,) * 3, np.float32); MASL.PCSWc3D(gma["pos"], g, gma["W"], box, 2)
    parity.assert_grid_close(g, gma["grid_PCSWc3D"], "PCSWc3D")
    g = np.zeros((dims,) * 3, np.float32); MASL.CICc3D(gma["pos"], g, box, 3)
    parity.assert_grid_close(g, gma["grid_CIC"], "CICc3D")


_REF_THREADS = {"CICc3D": 1, "TSCWc3D": 3}          # the thread counts Test/test_MAS.py passes (2 everywhere else)


@pytest.mark.parametrize("mas,places", [("NGP", 20), ("CIC", 8), ("TSC", 8), ("PCS", 8)])
@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("ndim", [2, 3])
def test_masc_shims_reference_unit_tests(MASL, mas, places, weighted, ndim):
    """Test/test_MAS.py:88-238 (test_NGPc3D ... test_PCSWc2D), restated case by case on the product's shims."""
    particles, BoxSize, dims, seed = 1000, 1.0, 64, 1
    np.random.seed(seed)
    pos = np.random.random((particles, ndim)).astype(np.float32)
    delta = np.zeros((dims,) * ndim, dtype=np.float32)
    name = "%s%sc%dD" % (mas, "W" if weighted else "", ndim)
    threads = _REF_THREADS.get(name, 2)
    if weighted:
        W = np.ones(particles, dtype=np.float32) * 3.0
        getattr(MASL, name)(pos, delta, W, BoxSize, threads)
    else:
        getattr(MASL, name)(pos, delta, BoxSize, threads)
    suma = np.sum(delta, dtype=np.float64)
    assert round(abs(suma / (3.0 * particles if weighted else particles) - 1.0), places) == 0


@pytest.mark.parametrize("mas", ["NGP", "CIC", "TSC", "PCS"])
@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("ndim", [2, 3])
def test_masc_shims_golden(MASL, gma, mas, weighted, ndim):
    """All sixteen <MAS>[W]c{2,3}D shims against the grids the reference's own MAS_c kernels produced."""
    box = float(gma["box"])
    dims = int(gma["dims"] if ndim == 3 else gma["dims2"])
    pos = gma["pos"] if ndim == 3 else np.ascontiguousarray(gma["pos"][:, :2])
    g = np.zeros((dims,) * ndim, np.float32)
    name = "%s%sc%dD" % (mas, "W" if weighted else "", ndim)
    if weighted:
        getattr(MASL, name)(pos, g, gma["W"], box, 2)
    else:
        getattr(MASL, name)(pos, g, box, 2)
    want = gma["c%dD_%s%s" % (ndim, mas, "W" if weighted else "")]
    if mas == "NGP" and not weighted:
        parity.assert_exact(g, want, name)
    parity.assert_grid_close(g, want, name)


def test_ma_fortran_order_and_device_tensors(MASL, gma):
    box, dims = float(gma["box"]), int(gma["dims"])
    g = np.zeros((dims,) * 3, np.float32)
    MASL.MA(np.asfortranarray(gma["pos"]), g, box, "CIC")
    parity.assert_grid_close(g, gma["grid_CIC"], "F-order pos")
    gt = torch.zeros((dims,) * 3, dtype=torch.float32, device="cuda")
    MASL.MA(torch.from_numpy(gma["pos"]).cuda(), gt, box, "PCS", W=torch.from_numpy(gma["W"]).cuda())
    parity.assert_grid_close(gt.cpu().numpy(), gma["grid_PCSW"], "device tensors")


def test_ma_host_chunked_streaming(MASL, gma):
    """Host particle arrays are streamed in chunks (H2D overlapped with the deposit); shrink the chunk so
    the path is exercised, including a ragged last chunk, weights and Fortran-ordered positions."""
    import pylians_b200.MAS_library as M
    box, dims = float(gma["box"]), int(gma["dims"])
    old, M.HOST_CHUNK = M.HOST_CHUNK, 1700
    oldf, M.HOST_TAPER_FLOOR = M.HOST_TAPER_FLOOR, 200          # the last full chunk is cut into 1/2, 1/4, 1/4
    try:
        g = np.zeros((dims,) * 3, np.float32); MASL.MA(gma["pos"], g, box, "TSC", W=gma["W"])
        parity.assert_grid_close(g, gma["grid_TSCW"], "chunked host TSCW")
        gt = torch.zeros((dims,) * 3, dtype=torch.float32, device="cuda")
        MASL.MA(np.asfortranarray(gma["pos"]), gt, box, "CIC")
        parity.assert_grid_close(gt.cpu().numpy(), gma["grid_CIC"], "chunked host CIC -> device grid")
    finally:
        M.HOST_CHUNK = old
        M.HOST_TAPER_FLOOR = oldf


def test_ma_errors(MASL):
    pos = np.zeros((4, 3), np.float32)
    with pytest.raises(SystemExit):
        MASL.MA(pos, np.zeros((8, 8), np.float32), 1.0, "CIC")
    with pytest.raises(SystemExit):
        MASL.MA(pos, np.zeros((8, 8, 8), np.float32), 1.0, "XYZ")
    with pytest.raises(ValueError):
        MASL.MA(pos.astype(np.float64), np.zeros((8, 8, 8), np.float32), 1.0, "CIC")
    with pytest.raises(ValueError):
        MASL.MA(pos, np.zeros((8, 8, 8), np.float64), 1.0, "CIC")
    MASL.MA(np.zeros((0, 3), np.float32), np.zeros((8, 8, 8), np.float32), 1.0, "CIC")   # empty input is fine


# the reference's own unit tests (Test/test_MAS.py:10-84): mass conservation
@pytest.mark.parametrize("mas,places", [("NGP", 20), ("CIC", 8), ("TSC", 8), ("PCS", 8)])
@pytest.mark.parametrize("weighted", [False, True])
def test_reference_unit_tests_mass_conservation(MASL, mas, places, weighted):
    particles, BoxSize, dims, seed = 1000, 1.0, 64, 1
    np.random.seed(seed)
    pos = np.random.random((particles, 3)).astype(np.float32)
    delta = np.zeros((dims, dims, dims), dtype=np.float32)
    W = np.ones(particles, dtype=np.float32) * 3.0 if weighted else None
    MASL.MA(pos, delta, BoxSize, mas, W=W)
    suma = np.sum(delta, dtype=np.float64)
    assert round(abs(suma / (3.0 * particles if weighted else particles) - 1.0), places) == 0


# algo 1: direct kernel.  algo 2 + debug path = 100 * kernel + sort: sort 0 automatic, 2 deep (second histogram sweep),
# 3 / 4 deep with 1024 / 4 lo digits (the digit widths of the largest grids); kernel 0 automatic, 1 lane per particle for
# every scheme, 2 stencil lanes where they exist (TSC, PCS)
@pytest.mark.parametrize("algo,path", [(1, 0), (2, 0), (2, 2), (2, 3), (2, 4), (2, 100), (2, 102), (2, 200), (2, 202), (2, 203),
                                       (2, 204)])
@pytest.mark.parametrize("mas", ["NGP", "CIC", "TSC", "PCS"])
@pytest.mark.parametrize("dims", [64, 80])
def test_ma_vs_oracle_both_algorithms(MASL, algo, path, mas, dims):
    """Seeded random + clustered particles, direct and tiled kernels (every sort path x every tile kernel) against the oracle."""
    from pylians_b200 import _lib
    _lib.load().pylb_ma_debug_path(path)
    rng = np.random.default_rng(100 + dims)
    box, n = 1000.0, 300000
    pos = (rng.random((n, 3)) * box).astype(np.float32)
    pos[: n // 4] = (np.float32(box) * 0.37 + rng.standard_normal((n // 4, 3)) * 45.0).astype(np.float32) % np.float32(box)
    pos[0] = 0.0; pos[1] = box; pos[2] = np.nextafter(np.float32(box), np.float32(0))
    W = (rng.random(n) + 0.5).astype(np.float32)
    old = MASL.ALGO if hasattr(MASL, "ALGO") else 0
    import pylians_b200.MAS_library as M
    M.ALGO = algo
    try:
        for w in (None, W):
            a = np.zeros((dims,) * 3, np.float32); b = np.zeros((dims,) * 3, np.float32)
            MASL.MA(pos, a, box, mas, W=w); O.MA(pos, b, box, mas, W=w)
            if mas == "NGP" and w is None:
                parity.assert_exact(a, b, "NGP")
            parity.assert_grid_close(a, b, "%s algo %d" % (mas, algo))
    finally:
        M.ALGO = old
        _lib.load().pylb_ma_debug_path(-1)


@pytest.mark.parametrize("dims,n", [(800, 6000000), (832, 3000000)])
def test_ma_deep_sort_large_grid(MASL, dims, n):
    """Grids with more than 53248 tiles take the deep sort by themselves (800^3: 62500 tiles; 832^3: 70304 tiles, more
    than 256 buckets)."""
    import pylians_b200.MAS_library as M
    rng = np.random.default_rng(dims)
    box = 1000.0
    pos = (rng.random((n, 3)) * box).astype(np.float32)
    pos[: n // 8] = (np.float32(box) * 0.5 + rng.standard_normal((n // 8, 3)) * 20.0).astype(np.float32) % np.float32(box)
    old, M.ALGO = M.ALGO, 2
    try:
        for mas in ("NGP", "PCS"):
            a = np.zeros((dims,) * 3, np.float32); MASL.MA(pos, a, box, mas)
            b = np.zeros((dims,) * 3, np.float32); O.MA(pos, b, box, mas)
            if mas == "NGP":
                parity.assert_exact(a, b, "NGP deep")
            parity.assert_grid_close(a, b, "%s deep %d" % (mas, dims))
            del a, b
    finally:
        M.ALGO = old


@pytest.mark.parametrize("mas", ["CIC", "TSC", "PCS"])
@pytest.mark.parametrize("kind", ["wide", "negative", "zero", "sparse"])
def test_ma_weights(MASL, mas, kind):
    """Weights spanning orders of magnitude, negative weights, all-zero weights and a very sparse deposit through the tiled path."""
    import pylians_b200.MAS_library as M
    rng = np.random.default_rng(5)
    dims, box = 96, 1000.0
    n = 3000 if kind == "sparse" else 400000
    pos = (rng.random((n, 3)) * box).astype(np.float32)
    W = {"wide": np.exp(rng.standard_normal(n) * 4.0), "negative": rng.standard_normal(n) + 0.25,
         "zero": np.zeros(n), "sparse": rng.random(n) + 0.5}[kind].astype(np.float32)
    old, M.ALGO = M.ALGO, 2
    try:
        a = np.full((dims,) * 3, 0.125, np.float32); MASL.MA(pos, a, box, mas, W=W)
    finally:
        M.ALGO = old
    b = np.full((dims,) * 3, 0.125, np.float32); O.MA(pos, b, box, mas, W=W)
    if kind == "negative":          # cancellations: compare against the scale of |W| deposited, not of the net field
        c = np.zeros((dims,) * 3, np.float32); O.MA(pos, c, box, mas, W=np.abs(W))
        assert np.all(np.abs(a.astype(np.float64) - b) <= 1e-5 * (c + np.mean(c)))
    else:
        parity.assert_grid_close(a, b, "%s weights %s" % (mas, kind))


@pytest.mark.parametrize("mas", ["CIC", "TSC", "PCS"])
def test_ma_and_pk_zeldovich_lattice(MASL, PKL, mas):
    """The other synthetic input of BASELINE.json: a Zel'dovich-displaced lattice in lattice order
    (spatially coherent particle order, mildly clustered), tiled deposit + Pk against the oracle."""
    import bench
    dims, box = 64, 1000.0
    gen = torch.Generator(device="cuda"); gen.manual_seed(3)
    pos_t = bench.zeldovich_particles(dims, box, gen, torch.device("cuda"))
    pos = pos_t.cpu().numpy()
    import pylians_b200.MAS_library as M
    old, M.ALGO = M.ALGO, 2
    try:
        a = torch.zeros((dims,) * 3, device="cuda"); MASL.MA(pos_t, a, box, mas)
    finally:
        M.ALGO = old
    b = np.zeros((dims,) * 3, np.float32); O.MA(pos, b, box, mas)
    parity.assert_grid_close(a.cpu().numpy(), b, "zeldovich " + mas)
    b /= np.mean(b, dtype=np.float64); b -= 1.0
    parity.check_pk(PKL.Pk(b, box, 2, mas, 1), O.Pk(b, box, 2, mas, 1))


def test_cabi_host_entry_points(gma):
    """The MAS_c.h-compatible symbols (host pointers), called the way a cgo/ctypes binding would."""
    from pylians_b200 import _lib
    lib = _lib.load()
    box, dims = float(gma["box"]), int(gma["dims"])
    pos = np.ascontiguousarray(gma["pos"]); W = np.ascontiguousarray(gma["W"])
    for name, key, w in (("CIC", "grid_CIC", None), ("PCS", "grid_PCSW", W), ("NGP", "grid_NGP", None), ("TSC", "grid_TSCW", W)):
        g = np.zeros((dims,) * 3, np.float32)
        getattr(lib, name)(pos.ctypes.data, g.ctypes.data, w.ctypes.data if w is not None else None,
                           pos.shape[0], dims, 3, ctypes.c_float(box), 4)
        assert lib.pylb_last_error() == b"", lib.pylb_last_error()
        parity.assert_grid_close(g, gma[key], "C ABI " + name)


@pytest.mark.parametrize("threads", [1, 3])
def test_cabi_host_entry_points_2d(gma, threads):
    """axes = 2 through the raw MAS_c.h ABI (MAS_c.c:21-34: one update per cell, no renormalisation), accumulating on
    top of what `number` already holds."""
    from pylians_b200 import _lib
    lib = _lib.load()
    box, dims = float(gma["box"]), int(gma["dims2"])
    pos = np.ascontiguousarray(gma["pos"][:, :2]); W = np.ascontiguousarray(gma["W"])
    for mas in ("NGP", "CIC", "TSC", "PCS"):
        for w in (None, W):
            g = np.full((dims,) * 2, 0.25, np.float32)
            getattr(lib, mas)(pos.ctypes.data, g.ctypes.data, w.ctypes.data if w is not None else None,
                              pos.shape[0], dims, 2, ctypes.c_float(box), threads)
            assert lib.pylb_last_error() == b"", lib.pylb_last_error()
            want = gma["c2D_%s%s" % (mas, "W" if w is not None else "")].astype(np.float64) + 0.25
            parity.assert_grid_close(g, want, "C ABI 2-D " + mas)


# ------------------------------------------------------------------------------------------------
# power spectra
# ------------------------------------------------------------------------------------------------
PK_NAMES = ["k3D", "Pk", "Nmodes3D", "Pkphase", "k1D", "Pk1D", "Nmodes1D", "kpar", "kper", "Pk2D", "Nmodes2D"]


@pytest.fixture(scope="module")
def gpk(golden_dir):
    return np.load(os.path.join(golden_dir, "pk.npz"))


@pytest.mark.parametrize("dims", [16, 20])
@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("mas", ["TSC", "None"])
def test_pk_golden(PKL, gpk, dims, axis, mas):
    p = PKL.Pk(gpk["delta_%d" % dims], float(gpk["box"]), axis, mas, 1)
    ref = {n: gpk["pk_%d_a%d_%s_%s" % (dims, axis, mas, n)] for n in PK_NAMES}
    parity.check_pk(p, ref)
    for n in PK_NAMES:
        assert np.asarray(getattr(p, n)).dtype == np.float64


def test_pk_keep_deltak(PKL, gpk):
    p = PKL.Pk(gpk["delta_16"], float(gpk["box"]), 2, "TSC", 1, keep_deltak=True)
    ref = gpk["pk_16_deltak"]
    assert p.delta_k.dtype == np.complex64 and p.delta_k.shape == ref.shape
    np.testing.assert_allclose(p.delta_k, ref, rtol=1e-4, atol=1e-4 * np.abs(ref).max())


def test_xpk_golden(PKL, golden_dir):
    g = np.load(os.path.join(golden_dir, "xpk.npz"))
    names = ["k3D", "Pk", "XPk", "Nmodes3D", "k1D", "Pk1D", "PkX1D", "Nmodes1D", "kpar", "kper", "Pk2D", "PkX2D", "Nmodes2D"]
    fs = [g["delta%d" % i] for i in range(3)]
    x = PKL.XPk(fs[:2], float(g["box"]), 2, ["CIC", "PCS"], 1)
    parity.check_xpk(x, {n: g["x2_a2_" + n] for n in names})
    x = PKL.XPk(fs, float(g["box"]), 0, ["CIC", "PCS", "None"], 1)
    parity.check_xpk(x, {n: g["x3_a0_" + n] for n in names})


def _field(dims, box, seed, mas="CIC"):
    rng = np.random.default_rng(seed)
    pos = (rng.random((2 * dims ** 3, 3)) * box).astype(np.float32)
    d = np.zeros((dims,) * 3, np.float32)
    O.MA(pos, d, box, mas)
    d /= np.mean(d, dtype=np.float64); d -= 1.0
    return d


# generic; ring2 (two kz per thread); one-kz ring cp.async fp32 / fp64 option / bulk fp32 / bulk fp64
@pytest.mark.parametrize("algo", [1, 2, 2 | 64, 2 | 16, 2 | 64 | 32, 2 | 16 | 32])
@pytest.mark.parametrize("dims", [48, 64, 33])
def test_pk_vs_oracle(PKL, algo, dims):
    import pylians_b200.Pk_library as P
    box = 750.0
    d = _field(dims, box, dims)
    old, P.ALGO = P.ALGO, algo
    try:
        for axis in (0, 1, 2):     # ring algos reach axes 0/1 through the real-space axis swap
            parity.check_pk(PKL.Pk(d, box, axis, "CIC", 1), O.Pk(d, box, axis, "CIC", 1))
    finally:
        P.ALGO = old


# 1 generic; 2 ring2x (two kz per thread, F = 2 / 3); 2|64 the one-kz ring kernel; 2|16 its fp64 option; 2|32 its bulk-copy loads
@pytest.mark.parametrize("algo", [1, 2, 2 | 64, 2 | 16, 2 | 32 | 64])
def test_xpk_vs_oracle(PKL, algo):
    import pylians_b200.Pk_library as P
    box, dims = 500.0, 40
    fs = [_field(dims, box, 5, "CIC"), _field(dims, box, 6, "TSC"), _field(dims, box, 7, "PCS")]
    old, P.ALGO = P.ALGO, algo
    try:
        parity.check_xpk(PKL.XPk(fs[:2], box, 2, ["CIC", "TSC"], 1), O.XPk(fs[:2], box, 2, ["CIC", "TSC"], 1))
        parity.check_xpk(PKL.XPk(fs, box, 2, ["CIC", "TSC", "PCS"], 1), O.XPk(fs, box, 2, ["CIC", "TSC", "PCS"], 1))
    finally:
        P.ALGO = old


@pytest.mark.parametrize("dims", [520, 258])
def test_ring2_segments_and_unaligned_rows(PKL, dims):
    """ring2 against the one-thread-per-mode kernel where the oracle is too slow: 520 -> two kz segments (the
    second with two pairs), 258 -> N/2+1 even, so every row starts on the same 16-byte parity; plus a field whose
    base pointer is only 8-byte aligned (the parity tables swap roles)."""
    import pylians_b200.Pk_library as P
    gen = torch.Generator(device="cuda"); gen.manual_seed(dims)
    d = torch.randn((dims,) * 3, device="cuda", dtype=torch.float32, generator=gen)
    old = P.ALGO
    try:
        P.ALGO = 1
        ref = PKL.Pk(d, 1000.0, 2, "PCS", 1)
        for algo in (2, 2 | 64):
            P.ALGO = algo
            parity.check_pk(PKL.Pk(d, 1000.0, 2, "PCS", 1), ref)
        # bins straight from a k-space field at an odd 8-byte offset
        m = dims // 2 + 1
        buf = torch.randn((dims * dims * m + 1, 2), device="cuda", dtype=torch.float32, generator=gen)
        dk = torch.view_as_complex(buf[1:]).reshape(dims, dims, m)
        assert dk.data_ptr() % 16 == 8
        _, ws, wc = P.bin_modes([dk], dims, 2, [2], True, False, algo=1)
        ws, wc = ws.cpu().numpy(), wc.cpu().numpy()
        for algo in (2, 2 | 64):
            _, gs, gc = P.bin_modes([dk], dims, 2, [2], True, False, algo=algo)
            assert np.array_equal(gc.cpu().numpy(), wc)              # mode counts: bit-exact
            np.testing.assert_allclose(gs.cpu().numpy(), ws, rtol=2e-6, atol=1e-7 * float(np.abs(ws).max()))
    finally:
        P.ALGO = old


@pytest.mark.parametrize("dims,F", [(520, 2), (264, 3)])
def test_ring2x_segments_and_unaligned_rows(PKL, dims, F):
    """The multi-field ring2x kernel (class XPk, F = 2 and 3) against the one-thread-per-mode kernel at sizes the oracle is
    too slow for: 520 -> two kz segments and several spans per level; 264 -> N/2+1 odd; through XPk (padded, aligned rows)
    and straight from dense k-space fields whose odd row pitch splits the row table by parity (8-byte aligned base)."""
    import pylians_b200.Pk_library as P
    gen = torch.Generator(device="cuda"); gen.manual_seed(dims + F)
    fields = [torch.randn((dims,) * 3, device="cuda", dtype=torch.float32, generator=gen) for _ in range(F)]
    fields[1] += 0.5 * fields[0]                                    # a real cross-correlation
    mas = ["CIC", "PCS", "None"][:F]
    old = P.ALGO
    try:
        P.ALGO = 1
        ref = PKL.XPk(fields, 1000.0, 2, mas, 1)
        for algo in (2, 2 | 64):
            P.ALGO = algo
            parity.check_xpk(PKL.XPk(fields, 1000.0, 2, mas, 1), ref)
        m = dims // 2 + 1
        bufs = [torch.randn((dims * dims * m + 1, 2), device="cuda", dtype=torch.float32, generator=gen) for _ in range(F)]
        dks = [torch.view_as_complex(b[1:]).reshape(dims, dims, m) for b in bufs]
        assert all(dk.data_ptr() % 16 == 8 for dk in dks)
        mi = [2, 4, 0][:F]
        _, ws, wc = P.bin_modes(dks, dims, 2, mi, False, False, algo=1)
        ws, wc = ws.cpu().numpy(), wc.cpu().numpy()
        for algo in (2, 2 | 64):
            _, gs, gc = P.bin_modes(dks, dims, 2, mi, False, False, algo=algo)
            assert np.array_equal(gc.cpu().numpy(), wc)              # mode counts: bit-exact
            np.testing.assert_allclose(gs.cpu().numpy(), ws, rtol=2e-6, atol=1e-7 * float(np.abs(ws).max()))
    finally:
        P.ALGO = old


def test_pk_device_tensor_input_and_untouched_delta(PKL):
    box, dims = 300.0, 32
    d = _field(dims, box, 3)
    dt = torch.from_numpy(d).cuda()
    keep = dt.clone()
    p = PKL.Pk(dt, box, 2, "CIC", 1)
    assert torch.equal(dt, keep)                       # Pk never mutates delta
    parity.check_pk(p, O.Pk(d, box, 2, "CIC", 1))


def test_swap_axes_kernel():
    from pylians_b200 import _lib
    lib = _lib.load()
    for N in (33, 64):
        x = torch.randn((N, N, N), device="cuda")
        for axis, perm in ((0, (2, 1, 0)), (1, (0, 2, 1)), (2, (0, 1, 2))):
            for pitch in (N, 2 * (N // 2 + 1)):
                out = torch.full((N, N, pitch), -7.0, device="cuda")
                _lib.check(lib.pylb_swap_axes(x.data_ptr(), out.data_ptr(), N, axis, pitch,
                                              torch.cuda.current_stream().cuda_stream), "pylb_swap_axes")
                assert torch.equal(out[:, :, :N], x.permute(*perm).contiguous())
                assert bool((out[:, :, N:] == -7.0).all())


def test_pk_axis_keep_deltak_keeps_reference_layout(PKL, gpk):
    # keep_deltak with axis != 2 must not swap axes: delta_k is returned in the reference's layout
    p = PKL.Pk(gpk["delta_16"], float(gpk["box"]), 0, "TSC", 1, keep_deltak=True)
    ref = gpk["pk_16_deltak"]          # deconvolution does not depend on the axis
    np.testing.assert_allclose(p.delta_k, ref, rtol=1e-4, atol=1e-4 * np.abs(ref).max())
    parity.check_pk(p, {n: gpk["pk_16_a0_TSC_%s" % n] for n in PK_NAMES})


@pytest.mark.parametrize("dims", [48, 33])
@pytest.mark.parametrize("axis", [0, 1])
def test_pk_axis_keep_deltak_fast_path_vs_generic(PKL, dims, axis):
    """Pk(axis=0|1, keep_deltak=True): spectra from the swapped field + delta_k from a write-back pass along z, against the
    one-thread-per-mode kernel binning along the requested axis."""
    import pylians_b200.Pk_library as P
    rng = np.random.default_rng(dims + axis)
    delta = rng.standard_normal((dims,) * 3).astype(np.float32)
    fast = PKL.Pk(delta, 1000.0, axis, "PCS", 1, keep_deltak=True)
    old, P.ALGO = P.ALGO, 1                                      # BIN_GENERIC: no swap, write-back in the generic kernel
    try:
        slow = PKL.Pk(delta, 1000.0, axis, "PCS", 1, keep_deltak=True)
    finally:
        P.ALGO = old
    assert fast.delta_k.shape == slow.delta_k.shape == (dims, dims, dims // 2 + 1)
    np.testing.assert_allclose(fast.delta_k, slow.delta_k, rtol=2e-5, atol=2e-5 * np.abs(slow.delta_k).max())
    parity.check_pk(fast, slow)


@pytest.mark.parametrize("fields,dims,axis", [(4, 48, 2), (5, 32, 0), (4, 32, 1)])
def test_xpk_more_than_three_fields_by_subsets(PKL, fields, dims, axis):
    """XPk with 4 and 5 fields (what Pk_Gadget needs for four particle types): binned three fields at a time by the ring
    kernel and assembled on the device, against the one-thread-per-mode kernel and the oracle."""
    import pylians_b200.Pk_library as P
    rng = np.random.default_rng(10 * fields + axis)
    base = rng.standard_normal((dims,) * 3).astype(np.float32)
    deltas = [(base * (0.3 + 0.2 * f) + rng.standard_normal((dims,) * 3)).astype(np.float32) for f in range(fields)]   # correlated fields
    mas = ["CIC", "PCS", "None", "TSC", "NGP"][:fields]
    fast = PKL.XPk(deltas, 1000.0, axis, mas, 1)
    old, P.ALGO = P.ALGO, 1                                      # BIN_GENERIC
    try:
        slow = PKL.XPk(deltas, 1000.0, axis, mas, 1)
    finally:
        P.ALGO = old
    assert fast.Pk.shape == slow.Pk.shape and fast.XPk.shape == slow.XPk.shape == (slow.Pk.shape[0], 3, fields * (fields - 1) // 2)
    parity.check_xpk(fast, slow)
    parity.check_xpk(fast, O.XPk(deltas, 1000.0, axis, mas, 1))


def test_pk_errors(PKL):
    with pytest.raises(ValueError):
        PKL.Pk(np.zeros((8, 8, 8), np.float64), 1.0, 2, "CIC", 1)
    with pytest.raises(SystemExit):
        PKL.XPk([np.zeros((8, 8, 8), np.float32), np.zeros((16, 16, 16), np.float32)], 1.0, 2, ["CIC", "CIC"], 1)
    with pytest.raises(TypeError):
        PKL.XPk([np.zeros((8, 8, 8), np.float32)] * 2, 1.0, 2, None, 1)


def test_plane_wave_known_answer(PKL):
    N, L, A = 64, 1000.0, 2.0
    z = np.arange(N)
    d = np.broadcast_to(A * np.cos(2 * np.pi * 5 * z / N), (N, N, N)).astype(np.float32).copy()
    p = PKL.Pk(d, L, 2, "None", 1)
    b = 4
    np.testing.assert_allclose(p.Pk[b, 0] * p.Nmodes3D[b], (A * N ** 3 / 2) ** 2 * (L / N ** 2) ** 3, rtol=1e-5)
    np.testing.assert_allclose(p.Pk[b, 1] / p.Pk[b, 0], 5.0, rtol=1e-5)
    np.testing.assert_allclose(p.Pk[b, 2] / p.Pk[b, 0], 9.0, rtol=1e-5)
    p = PKL.Pk(d, L, 0, "None", 1)
    np.testing.assert_allclose(p.Pk[b, 1] / p.Pk[b, 0], -2.5, rtol=1e-5)
    np.testing.assert_allclose(p.Pk[b, 2] / p.Pk[b, 0], 3.375, rtol=1e-5)


def test_rsd_and_overdensity(golden_dir):
    import redshift_space_library as RSL
    from pylians_b200 import _lib
    g = np.load(os.path.join(golden_dir, "rsd.npz"))
    for axis in (0, 1, 2):
        a = g["pos"].copy()
        RSL.pos_redshift_space(a, g["vel"], float(g["box"]), float(g["hubble"]), float(g["redshift"]), axis)
        assert np.array_equal(a, g["rsd_a%d" % axis])
    rng = np.random.default_rng(0)
    x = (rng.random((40, 40, 40)) * 3 + 0.1).astype(np.float32)
    want = x.copy(); want /= np.mean(want, dtype=np.float64); want -= 1.0
    t = torch.from_numpy(x).cuda(); scratch = torch.zeros(2, dtype=torch.float64, device="cuda")
    _lib.check(_lib.load().pylb_overdensity(t.data_ptr(), t.numel(), scratch.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream), "pylb_overdensity")
    np.testing.assert_allclose(t.cpu().numpy(), want, rtol=0, atol=3e-7)


# ------------------------------------------------------------------------------------------------
# full-size properties (BASELINE configs[1]: 512^3 particles, CIC, 512^3 grid)
# ------------------------------------------------------------------------------------------------
def test_full_size_properties(MASL, PKL):
    dims, box = 512, 1000.0
    gen = torch.Generator(device="cuda"); gen.manual_seed(1)
    pos = torch.rand((dims ** 3, 3), device="cuda", dtype=torch.float32, generator=gen) * box
    grid = torch.zeros((dims,) * 3, device="cuda", dtype=torch.float32)
    MASL.MA(pos, grid, box, "CIC")
    total = float(grid.sum(dtype=torch.float64))
    assert abs(total / dims ** 3 - 1.0) < 1e-6                       # mass conservation
    # linearity / algorithm independence: direct kernel on the same input gives the same grid
    import pylians_b200.MAS_library as M
    grid2 = torch.zeros_like(grid)
    old, M.ALGO = M.ALGO, 1
    try:
        MASL.MA(pos, grid2, box, "CIC")
    finally:
        M.ALGO = old
    err = (grid - grid2).abs().max().item()
    assert err <= 1e-5 * float(grid.max()), err
    del grid2, pos
    grid /= grid.mean(dtype=torch.float64).float(); grid -= 1.0
    p = PKL.Pk(grid, box, 2, "CIC", 1)
    assert p.Nmodes3D.sum() + 1 == (dims ** 3 - 8) // 2 + 8           # Pk_library.pyx:90-102
    kF = 2 * np.pi / box
    i = np.arange(len(p.k3D))
    assert np.all(p.k3D >= (i + 1) * kF * (1 - 1e-12)) and np.all(p.k3D < (i + 2) * kF)
    # Poisson shot noise: P0 ~ V/N_particles for uniform random particles (low k, where the
    # deconvolved aliased shot noise is still flat to < 1%)
    w = p.Nmodes3D
    shot = box ** 3 / dims ** 3
    assert abs(np.sum(p.Pk[5:40, 0] * w[5:40]) / np.sum(w[5:40]) / shot - 1.0) < 0.02
    # the generic (atomics) kernel agrees with the ring kernel at full size
    import pylians_b200.Pk_library as P
    old, P.ALGO = P.ALGO, 1
    try:
        q = PKL.Pk(grid, box, 2, "CIC", 1)
    finally:
        P.ALGO = old
    parity.check_pk(p, q)


# ------------------------------------------------------------------------------------------------
# full-size properties at 1024^3 (BASELINE configs[2]: TSC, multipoles along axis 2; configs[4]: two fields, XPk)
# ------------------------------------------------------------------------------------------------
def _total_power(P0, Nmodes):
    return float(np.sum(np.asarray(P0, np.float64) * np.asarray(Nmodes, np.float64)))


def test_full_size_1024_tsc_multipoles_and_xpk(MASL, PKL):
    dims, box = 1024, 1000.0
    gen = torch.Generator(device="cuda"); gen.manual_seed(3)
    fields = []
    for mas, weighted in (("TSC", False), ("CIC", True)):
        pos = torch.rand((dims ** 3, 3), device="cuda", dtype=torch.float32, generator=gen) * box
        W = (torch.rand(dims ** 3, device="cuda", dtype=torch.float32, generator=gen) + 0.5) if weighted else None
        grid = torch.zeros((dims,) * 3, device="cuda", dtype=torch.float32)
        MASL.MA(pos, grid, box, mas, W=W)
        want = float(W.sum(dtype=torch.float64)) if weighted else float(dims ** 3)
        assert abs(float(grid.sum(dtype=torch.float64)) / want - 1.0) < 1e-6      # mass conservation
        del pos, W
        MASL.overdensity(grid)
        assert abs(float(grid.mean(dtype=torch.float64))) < 1e-6
        fields.append(grid)
    # configs[2]: redshift-space-style multipoles along axis 2 of the TSC field
    p = PKL.Pk(fields[0], box, 2, "TSC", 1)
    assert p.Nmodes3D.sum() + 1 == (dims ** 3 - 8) // 2 + 8                       # Pk_library.pyx:90-102
    assert p.Nmodes1D.sum() + p.Nmodes2D.sum() > 0
    shot = box ** 3 / dims ** 3
    w = p.Nmodes3D
    assert abs(np.sum(p.Pk[5:60, 0] * w[5:60]) / np.sum(w[5:60]) / shot - 1.0) < 0.01
    # isotropic input: quadrupole and hexadecapole vanish within the scatter of (2l+1) L_l(mu) |delta_k|^2 per bin
    for ell, fac in ((1, 5.0), (2, 9.0)):
        sig = fac * shot / np.sqrt(w[20:150])                                      # k < 0.3 k_Nyquist: aliasing still isotropic
        assert np.all(np.abs(p.Pk[20:150, ell]) < 6.0 * sig), ell
    # one mode set, three binnings: sum of |delta_k|^2 over the 2-D table equals the 3-D sum plus the DC mode's bin
    tot3 = _total_power(p.Pk[:, 0], p.Nmodes3D)
    tot2 = _total_power(p.Pk2D, p.Nmodes2D)
    assert abs(tot2 / tot3 - 1.0) < 1e-7
    # the line of sight only relabels modes: the monopole and the mode counts do not depend on it
    q = PKL.Pk(fields[0], box, 0, "TSC", 1)
    parity.assert_exact(q.Nmodes3D, p.Nmodes3D, "Nmodes3D axis 0 vs 2")
    np.testing.assert_allclose(q.Pk[:, 0], p.Pk[:, 0], rtol=1e-5, atol=0)
    del q
    # configs[4]: auto and cross spectra of the two fields with per-field MAS deconvolution
    x = PKL.XPk(fields, box, 2, ["TSC", "CIC"], 1)
    parity.assert_exact(x.Nmodes3D, p.Nmodes3D, "XPk Nmodes3D")
    np.testing.assert_allclose(x.Pk[:, :, 0], p.Pk, rtol=1e-5, atol=1e-5 * shot)   # same field, XPk path vs Pk path
    assert np.all(np.abs(x.XPk[:, 0, 0]) <= np.sqrt(x.Pk[:, 0, 0] * x.Pk[:, 0, 1]) * (1 + 1e-6))   # Cauchy-Schwarz
    y = PKL.XPk(fields[::-1], box, 2, ["CIC", "TSC"], 1)                          # the cross spectrum is symmetric
    np.testing.assert_allclose(y.XPk[:, :, 0], x.XPk[:, :, 0], rtol=1e-6, atol=1e-7 * shot)
    np.testing.assert_allclose(y.Pk[:, :, 1], x.Pk[:, :, 0], rtol=1e-6, atol=1e-7 * shot)


@pytest.mark.parametrize("kernel", [100, 200])
@pytest.mark.parametrize("mas", ["NGP", "CIC", "TSC", "PCS"])
@pytest.mark.parametrize("weighted", [False, True])
def test_ma_clustered_input(MASL, mas, weighted, kernel):
    """Half of one tile's particles sit in three cells (a halo): many lanes / consecutive particles share a base cell and
    every warp of the CTA hammers the same few words -- the CAS loops of the lane-per-particle kernel and the optimistic
    CAS + repair of the stencil-lane kernel must both give the oracle's grid."""
    from oracle import pylians_oracle as O
    from pylians_b200 import _lib
    import pylians_b200.MAS_library as M
    dims, box = 64, 640.0                                   # 10 length units per cell
    rng = np.random.default_rng(31)
    uni = rng.random((150000, 3)) * box
    hot = np.concatenate([c + rng.normal(0.0, 2.0, (1500, 3)) for c in
                          (np.array([105.0, 85.0, 155.0]), np.array([104.0, 93.0, 161.0]), np.array([325.0, 325.0, 325.0]))])
    pos = np.mod(np.concatenate([uni, hot]), box).astype(np.float32)
    pos = pos[rng.permutation(len(pos))]
    pos[:300] = pos[0]                                      # and a run of identical particles
    W = (rng.random(len(pos)) + 0.5).astype(np.float32) if weighted else None
    ref = np.zeros((dims,) * 3, np.float32)
    O.MA(pos, ref, box, mas, W=W)
    old, M.ALGO = M.ALGO, 2                                 # the tiled (shared-memory) deposit, also on this small grid
    _lib.load().pylb_ma_debug_path(kernel)                  # automatic sort; 1 lane per particle, 2 stencil lanes (PCS)
    try:
        got = np.zeros((dims,) * 3, np.float32)
        MASL.MA(pos, got, box, mas, W=W)
    finally:
        M.ALGO = old
        _lib.load().pylb_ma_debug_path(-1)
    assert ref.max() > 50 * ref.mean()                      # the halo cells really are hot
    parity.assert_grid_close(got, ref, "clustered " + mas)
