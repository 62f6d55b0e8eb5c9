"""Generate golden vectors from the UNMODIFIED reference (oracle/_ref, built by oracle/build_ref.py
from /root/reference).  Run in the build container only:

    python tests/golden/make_golden.py

Writes tests/golden/*.npz (inputs + reference outputs).  The reference cannot travel to the GPU
box, these files can.  Sizes are tiny on purpose (whole directory < 3 MB).
"""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
MASL, PKL = ref_loader.load()
RSL = ref_loader.load_rsl()


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def particles(seed, n, box, ndim=3):
    rng = np.random.default_rng(seed)
    pos = (rng.random((n, ndim)) * box).astype(np.float32)
    # edge cases the reference accepts: exactly 0, exactly BoxSize (wraps to cell 0), just below
    # BoxSize, exactly on grid points and on cell mid-points.
    pos[0] = 0.0
    pos[1] = box
    pos[2] = np.nextafter(np.float32(box), np.float32(0))
    pos[3] = box / 4.0
    pos[4] = box / 4.0 + box / 48.0
    W = (rng.random(n) * 2.0 + 0.1).astype(np.float32)
    return pos, W


def gold_ma():
    out = {}
    box, dims, n = 1000.0, 24, 6000
    pos, W = particles(1, n, box)
    out["box"], out["dims"], out["pos"], out["W"] = box, dims, pos, W
    for mas in ("NGP", "CIC", "TSC", "PCS"):
        for wname, w in (("", None), ("W", W)):
            g = np.zeros((dims,) * 3, np.float32)
            MASL.MA(pos, g, box, mas, W=w)
            out["grid_%s%s" % (mas, wname)] = g
    # accumulate-in-place contract: second call adds on top of a non-zero grid
    g = np.full((dims,) * 3, 0.5, np.float32)
    MASL.MA(pos[:1000], g, box, "CIC")
    MASL.MA(pos[1000:2000], g, box, "TSC", W=W[1000:2000])
    out["grid_accum"] = g
    # 2-D
    dims2 = 32
    pos2 = np.ascontiguousarray(pos[:, :2])
    out["dims2"] = dims2
    for mas in ("NGP", "CIC", "TSC", "PCS"):
        for wname, w in (("", None), ("W", W)):
            g = np.zeros((dims2,) * 2, np.float32)
            MASL.MA(pos2, g, box, mas, W=w)
            out["grid2d_%s%s" % (mas, wname)] = g
    g = np.zeros((dims2,) * 2, np.float32)
    MASL.MA(pos2, g, box, "TSC", renormalize_2D=False)
    out["grid2d_TSC_norenorm"] = g
    # fp64 grids (NGPW_d / CICW_d)
    g = np.zeros((dims,) * 3, np.float64); MASL.NGPW_d(pos, g, box, W); out["grid_NGPW_d"] = g
    g = np.zeros((dims,) * 3, np.float64); MASL.CICW_d(pos, g, box, W); out["grid_CICW_d"] = g
    # OpenMP C entry points (MAS_c.c) -- same numbers up to summation order
    g = np.zeros((dims,) * 3, np.float32); MASL.PCSWc3D(pos, g, W, box, 2); out["grid_PCSWc3D"] = g
    # all sixteen MAS_c shims of Test/test_MAS.py:88-238 (3-D and 2-D, plain and weighted): the 2-D C kernels add
    # every contribution ONCE (n_max = 1, MAS_c.c:21-34) and nothing is renormalised afterwards
    for mas in ("NGP", "CIC", "TSC", "PCS"):
        g = np.zeros((dims,) * 3, np.float32); getattr(MASL, mas + "c3D")(pos, g, box, 2); out["c3D_%s" % mas] = g
        g = np.zeros((dims,) * 3, np.float32); getattr(MASL, mas + "Wc3D")(pos, g, W, box, 3); out["c3D_%sW" % mas] = g
        g = np.zeros((dims2,) * 2, np.float32); getattr(MASL, mas + "c2D")(pos2, g, box, 2); out["c2D_%s" % mas] = g
        g = np.zeros((dims2,) * 2, np.float32); getattr(MASL, mas + "Wc2D")(pos2, g, W, box, 1); out["c2D_%sW" % mas] = g
    np.savez_compressed(os.path.join(HERE, "ma.npz"), **out)


def pk_attrs(p, names):
    return {n: np.asarray(getattr(p, n)) for n in names}


PK_NAMES = ["k3D", "Pk", "Nmodes3D", "Pkphase", "k1D", "Pk1D", "Nmodes1D", "kpar", "kper", "Pk2D", "Nmodes2D"]
XPK_NAMES = ["k3D", "Pk", "XPk", "Nmodes3D", "k1D", "Pk1D", "PkX1D", "Nmodes1D", "kpar", "kper", "Pk2D", "PkX2D", "Nmodes2D"]


def fields(dims, box, seeds_mas):
    fs = []
    for seed, mas, weighted in seeds_mas:
        pos, W = particles(seed, 4 * dims ** 3, box)
        g = np.zeros((dims,) * 3, np.float32)
        MASL.MA(pos, g, box, mas, W=W if weighted else None)
        g /= np.mean(g, dtype=np.float64)
        g -= 1.0
        fs.append(g)
    return fs


def gold_pk():
    out = {}
    box = 1000.0
    for dims in (16, 20):
        (d,) = fields(dims, box, [(7, "TSC", False)])
        out["delta_%d" % dims] = d
        for axis in (0, 1, 2):
            for mas in ("TSC", "None"):
                p = quiet(PKL.Pk, d, box, axis, mas, 1)
                for n, v in pk_attrs(p, PK_NAMES).items():
                    out["pk_%d_a%d_%s_%s" % (dims, axis, mas, n)] = v
    p = quiet(PKL.Pk, out["delta_16"], box, 2, "TSC", 1, True)
    out["pk_16_deltak"] = np.asarray(p.delta_k)
    out["box"] = box
    np.savez_compressed(os.path.join(HERE, "pk.npz"), **out)


def gold_xpk():
    out = {}
    box, dims = 750.0, 16
    fs = fields(dims, box, [(11, "CIC", False), (12, "PCS", True), (13, "NGP", False)])
    out["box"], out["dims"] = box, dims
    for i, f in enumerate(fs):
        out["delta%d" % i] = f
    x = quiet(PKL.XPk, fs[:2], box, 2, ["CIC", "PCS"], 1)
    for n, v in pk_attrs(x, XPK_NAMES).items():
        out["x2_a2_%s" % n] = v
    x = quiet(PKL.XPk, fs, box, 0, ["CIC", "PCS", "None"], 1)
    for n, v in pk_attrs(x, XPK_NAMES).items():
        out["x3_a0_%s" % n] = v
    np.savez_compressed(os.path.join(HERE, "xpk.npz"), **out)


def gold_rsd():
    if RSL is None:
        return
    rng = np.random.default_rng(5)
    box = 100.0
    pos = (rng.random((4000, 3)) * box).astype(np.float32)
    vel = (rng.standard_normal((4000, 3)) * 2500).astype(np.float32)
    out = {"pos": pos, "vel": vel, "box": box, "hubble": 100.0, "redshift": 0.5}
    for axis in (0, 1, 2):
        a = pos.copy()
        RSL.pos_redshift_space(a, vel, box, 100.0, 0.5, axis)
        out["rsd_a%d" % axis] = a
    np.savez_compressed(os.path.join(HERE, "rsd.npz"), **out)


def gold_siblings():
    """Pk_plane / XPk_plane / Pk_theta / correct_MAS / Xi of the unmodified reference (Pk_library.pyx)."""
    out = {}
    box = 500.0
    rng = np.random.default_rng(21)
    for grid in (16, 18):                      # N/2+1 odd and even
        a = rng.standard_normal((grid, grid)).astype(np.float32)
        b = (0.5 * a + rng.standard_normal((grid, grid))).astype(np.float32)
        out["img_a_%d" % grid], out["img_b_%d" % grid] = a, b
        p = quiet(PKL.Pk_plane, a, box, "CIC", 1)
        for n in ("k", "Nmodes", "Pk"):
            out["plane_%d_%s" % (grid, n)] = np.asarray(getattr(p, n))
        x = quiet(PKL.XPk_plane, a, b, box, "CIC", "TSC", 1)
        for n in ("k", "Nmodes", "Pk", "XPk", "r"):
            out["xplane_%d_%s" % (grid, n)] = np.asarray(getattr(x, n))
        v = [rng.standard_normal((grid,) * 3).astype(np.float32) for _ in range(3)]
        for i in range(3):
            out["vel%d_%d" % (i, grid)] = v[i]
        t = quiet(PKL.Pk_theta, v[0], v[1], v[2], box, 2, "PCS", 1)
        for n, arr in zip(("k", "Pk", "Nmodes"), t):
            out["theta_%d_%s" % (grid, n)] = np.asarray(arr)
        (d,) = fields(grid, box, [(31 + grid, "CIC", False)])
        out["delta_%d" % grid] = d
        out["correct_%d" % grid] = np.asarray(quiet(PKL.correct_MAS, d, box, "CIC", 1))
        for axis in (0, 1, 2):
            xi = quiet(PKL.Xi, d, box, "CIC", axis, 1)
            for n in ("r3D", "xi", "Nmodes3D"):
                out["xi_%d_a%d_%s" % (grid, axis, n)] = np.asarray(getattr(xi, n))
    # XPk_imag: three fields, line of sight along z and along x
    fs = fields(16, box, [(41, "CIC", False), (42, "TSC", True), (43, "PCS", False)])
    for i, f in enumerate(fs):
        out["ximag_delta%d" % i] = f
    for axis in (2, 0):
        x = quiet(PKL.XPk_imag, fs, box, axis, ["CIC", "TSC", "PCS"], 1)
        for n, v in pk_attrs(x, XPK_NAMES).items():
            out["ximag_a%d_%s" % (axis, n)] = v
    out["box"] = box
    np.savez_compressed(os.path.join(HERE, "siblings.npz"), **out)


DRIVER_SNAPSHOT = dict(seed=9, counts=[3000, 5000, 0, 0, 1200, 0], box_kpc=40000.0, redshift=1.0,
                       masstable=np.array([0.0, 0.7, 0.0, 0.0, 0.0, 0.0]), nfiles=3)


def gold_drivers():
    """Snapshot drivers: the reference's density_field_gadget / Pk_comp / Pk_Gadget, compiled
    unmodified, on a synthetic 3-file format-1 snapshot that tests/gadget_writer.py regenerates from its seed."""
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import gadget_writer as GW
    ex = ref_loader.load_extras()
    MG, PS = ex["MAS_gadget"], ex["Pk_snapshot"]
    S = DRIVER_SNAPSHOT
    parts = GW.make_particles(S["seed"], S["counts"], S["box_kpc"], S["masstable"], clustered=True)
    out = {"pos_checksum": np.array([float(np.sum(parts[t][0], dtype=np.float64)) for t in (0, 1, 4)])}
    dims = 16
    with tempfile.TemporaryDirectory() as tmp:
        base = os.path.join(tmp, "snap_005")
        GW.write_snapshot(base, parts, S["masstable"], S["box_kpc"], S["redshift"], S["nfiles"], 1)
        # density fields: one species with a header mass (counts), one with a MASS block, RSD, two with MASS blocks
        out["dens_cdm_CIC"] = quiet(MG.density_field_gadget, base, [1], dims, "CIC", False, 0, False)
        out["dens_cdm_PCS_rsd1"] = quiet(MG.density_field_gadget, base, [1], dims, "PCS", True, 1, False)
        out["dens_gas_TSC"] = quiet(MG.density_field_gadget, base, [0], dims, "TSC", False, 0, False)
        out["dens_gas_stars_CIC_rsd2"] = quiet(MG.density_field_gadget, base, [0, 4], dims, "CIC", True, 2, False)
        for tag, types, rsd, axis in (("cdm", [1], False, 0), ("cdm_rs2", [1], True, 2), ("gas_cdm", [0, 1], False, 0),
                                      ("gas_cdm_stars_rs0", [0, 1, 4], True, 0)):
            folder = os.path.join(tmp, tag)
            os.makedirs(folder)
            quiet(PS.Pk_Gadget, base, dims, types, rsd, axis, 1, folder)
            for f in sorted(os.listdir(folder)):
                out["pk_%s__%s" % (tag, f)] = np.loadtxt(os.path.join(folder, f))
    np.savez_compressed(os.path.join(HERE, "drivers.npz"), **out)


def gold_consumers():
    """FFT consumers: smoothing_library, void_library.gaussian_smoothing, bispectrum_library.Bk of the
    compiled, unmodified reference."""
    ex = ref_loader.load_extras()
    SL, VL, BL = ex["smoothing_library"], ex["void_library"], ex["bispectrum_library"]
    out = {}
    box = 200.0
    for dims in (16, 18):
        (d,) = fields(dims, box, [(51 + dims, "CIC", False)])
        out["delta_%d" % dims] = d
        for name, R in (("Top-Hat", 31.0), ("Gaussian", 17.5)):
            W_k = np.asarray(SL.FT_filter(box, R, dims, name, 2))
            out["filter_%d_%s" % (dims, name)] = W_k
            out["smooth_%d_%s" % (dims, name)] = np.asarray(SL.field_smoothing(d, W_k, 2))
        out["void_smooth_%d" % dims] = np.asarray(VL.gaussian_smoothing(d, box, 23.0, 2))
        kF = 2.0 * np.pi / box
        theta = np.array([0.3, 1.1, 2.0, 2.9])
        b = quiet(BL.Bk, d, box, 3.0 * kF, 4.2 * kF, theta, "CIC", 1)
        for n in ("B", "Q", "k", "Pk"):
            out["bk_%d_%s" % (dims, n)] = np.asarray(getattr(b, n))
    out["box"], out["theta"] = box, theta
    np.savez_compressed(os.path.join(HERE, "consumers.npz"), **out)


if __name__ == "__main__":
    # python tests/golden/make_golden.py [ma pk xpk rsd siblings ...]   (default: all)
    which = sys.argv[1:] or ["ma", "pk", "xpk", "rsd", "siblings", "drivers", "consumers"]
    for name in which:
        globals()["gold_" + name]()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
