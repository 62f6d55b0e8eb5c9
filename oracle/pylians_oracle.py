"""numpy/ctypes front end of the CPU oracle (oracle/pyl_oracle.c).  TEST INFRASTRUCTURE ONLY.

Mirrors the reference's Python surface for the hot path so parity tests read like reference
usage (`MA(pos, delta, BoxSize, MAS, W)`, `Pk(delta, BoxSize, axis, MAS, threads)`,
`XPk([d1, d2], BoxSize, axis, MAS=[..], threads)`), citing /root/reference/library files.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (pylians_b200) never does.

Third-party arithmetic on the path: the reference's FFT is pyfftw -> FFTW3 (Pk_library.pyx:3,
:120-133), un-vendored and un-pinned in the reference tree.  The oracle restates it with
scipy.fft.rfftn (pocketfft) in float32 -> complex64, the same precision as FFT3Dr_f.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(HERE, "pyl_oracle.c")
_SO = os.path.join(HERE, "_build", "libpyl_oracle.so")
_LIB = None

MAS_ID = {"NGP": 0, "CIC": 1, "TSC": 2, "PCS": 3}


def build(force=False):
    """gcc -O2 build of the C restatement (no -ffast-math: the oracle keeps IEEE semantics)."""
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.check_call([cc, "-O2", "-fPIC", "-shared", "-std=gnu99", "-o", _SO, _SRC, "-lm"])
    return _SO


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = ctypes.CDLL(_SO)
        fp, dp, ip = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
        L.orc_ma_f32.argtypes = [fp, ctypes.c_long, ctypes.c_int, ctypes.c_long, ctypes.c_long, fp,
                                 ctypes.c_int, ctypes.c_float, ctypes.c_int, fp]
        L.orc_ma_f64.argtypes = [fp, ctypes.c_long, ctypes.c_int, ctypes.c_long, ctypes.c_long, dp,
                                 ctypes.c_int, ctypes.c_float, ctypes.c_int, fp]
        L.orc_pos_redshift_space.argtypes = [fp, fp, ctypes.c_long, ctypes.c_float, ctypes.c_float,
                                             ctypes.c_float, ctypes.c_int]
        L.orc_pk_loop.argtypes = [ctypes.POINTER(fp), ctypes.c_int, ctypes.c_int, ctypes.c_int, ip,
                                  ctypes.c_int, ctypes.c_int, ctypes.c_int] + [dp] * 12
        for f in (L.orc_ma_f32, L.orc_ma_f64, L.orc_pos_redshift_space, L.orc_pk_loop):
            f.restype = None
        _LIB = L
    return _LIB


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


# ------------------------------------------------------------------------------------------------
# MAS_library.pyx:57-112
# ------------------------------------------------------------------------------------------------
def MA(pos, number, BoxSize, MAS="CIC", W=None, verbose=False, renormalize_2D=True):
    coord, coord_aux = pos.shape[1], number.ndim
    if coord != coord_aux:                       # :64-66
        print("pos have %d dimensions and the density %d!!!" % (coord, coord_aux))
        sys.exit()
    if MAS not in MAS_ID:                        # :81-82
        print("option not valid!!!")
        sys.exit()
    if pos.dtype != np.float32 or (W is not None and W.dtype != np.float32):
        raise ValueError("Buffer dtype mismatch, expected 'float32_t'")
    if number.dtype not in (np.float32, np.float64) or not number.flags["C_CONTIGUOUS"]:
        raise ValueError("number must be a C-contiguous float32 (or float64) array")
    dims = number.shape[0]
    es = pos.itemsize
    ps0, ps1 = pos.strides[0] // es, pos.strides[1] // es
    Wp = None
    if W is not None:
        W = np.ascontiguousarray(W)
        Wp = _fp(W)
    L = lib()
    if number.dtype == np.float32:
        L.orc_ma_f32(_fp(pos), pos.shape[0], coord, ps0, ps1, _fp(number), dims, float(BoxSize), MAS_ID[MAS], Wp)
    else:
        L.orc_ma_f64(_fp(pos), pos.shape[0], coord, ps0, ps1, _dp(number), dims, float(BoxSize), MAS_ID[MAS], Wp)
    if coord == 2 and renormalize_2D and MAS != "NGP":   # :90-107: the WHOLE array is divided
        number /= {"CIC": 2.0, "TSC": 3.0, "PCS": 4.0}[MAS]


def pos_redshift_space(pos, vel, BoxSize, Hubble, redshift, axis):
    """redshift_space_library.pyx:29-43 (in place on pos)."""
    assert pos.dtype == np.float32 and vel.dtype == np.float32 and pos.flags["C_CONTIGUOUS"] and vel.flags["C_CONTIGUOUS"]
    lib().orc_pos_redshift_space(_fp(pos), _fp(vel), pos.shape[0], float(BoxSize), float(Hubble), float(redshift), int(axis))


# ------------------------------------------------------------------------------------------------
# Pk_library.pyx:59-87
# ------------------------------------------------------------------------------------------------
def frequencies(BoxSize, dims):
    kF = 2.0 * np.pi / BoxSize
    middle = dims // 2                            # python-2 integer division in the reference :60
    kN = middle * kF
    kmax_par = middle
    kmax_per = int(np.sqrt(middle ** 2 + middle ** 2))
    kmax = int(np.sqrt(middle ** 2 + middle ** 2 + middle ** 2))
    return kF, kN, kmax_par, kmax_per, kmax


def MAS_function(MAS):
    return {"NGP": 1, "CIC": 2, "TSC": 3, "PCS": 4}.get(MAS, 0)


def FFT3Dr_f(a, threads=1):
    """Pk_library.pyx:120-133 -- unnormalised forward R2C, float32 -> complex64, half-spectrum on z."""
    import scipy.fft as sf
    assert a.dtype == np.float32 and a.ndim == 3
    return np.ascontiguousarray(sf.rfftn(a, axes=(0, 1, 2), workers=threads).astype(np.complex64, copy=False))


def check_number_modes(Nmodes, dims):            # :90-102
    own = 1 if dims % 2 == 1 else 8
    indep = (dims ** 3 - own) // 2 + own
    if int(np.sum(Nmodes)) != indep:
        print("WARNING: Not all modes counted")
        sys.exit()


def _mode_loop(delta_k_list, dims, axis, mas_index, BoxSize, want_phase):
    F = len(delta_k_list)
    X = F * (F - 1) // 2
    kF, kN, kmax_par, kmax_per, kmax = frequencies(BoxSize, dims)
    B2 = (kmax_par + 1) * (kmax_per + 1)
    z = lambda *s: np.zeros(s, dtype=np.float64)
    out = dict(k3d=z(kmax + 1), n3d=z(kmax + 1), p3d=z(kmax + 1, 3, F), x3d=z(kmax + 1, 3, max(X, 1)),
               phase=z(kmax + 1), k1d=z(kmax_par + 1), n1d=z(kmax_par + 1), p1d=z(kmax_par + 1, F),
               x1d=z(kmax_par + 1, max(X, 1)), n2d=z(B2), p2d=z(B2, F), x2d=z(B2, max(X, 1)))
    fpp = ctypes.POINTER(ctypes.c_float)
    ptrs = (fpp * F)(*[dk.ctypes.data_as(fpp) for dk in delta_k_list])
    mi = np.asarray(mas_index, dtype=np.int32)
    # x-arrays are allocated with max(X,1) columns but addressed with stride X inside C; for X==0
    # they are never touched.
    lib().orc_pk_loop(ptrs, F, dims, int(axis), mi.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                      kmax_par, kmax_per, kmax, _dp(out["k3d"]), _dp(out["n3d"]), _dp(out["p3d"]),
                      _dp(out["x3d"]), _dp(out["phase"]) if want_phase else None, _dp(out["k1d"]),
                      _dp(out["n1d"]), _dp(out["p1d"]), _dp(out["x1d"]), _dp(out["n2d"]),
                      _dp(out["p2d"]), _dp(out["x2d"]))
    if X == 0:
        out["x3d"] = out["x3d"][:, :, :0]; out["x1d"] = out["x1d"][:, :0]; out["x2d"] = out["x2d"][:, :0]
    out.update(kF=kF, kN=kN, kmax_par=kmax_par, kmax_per=kmax_per, kmax=kmax, X=X, F=F)
    return out


def _kpar_kper(kmax_par, kmax_per, kF):          # :397-403
    kpar = np.zeros((kmax_par + 1) * (kmax_per + 1)); kper = np.zeros_like(kpar)
    for k_per in range(kmax_per + 1):
        sl = slice((kmax_par + 1) * k_per, (kmax_par + 1) * (k_per + 1))
        kpar[sl] = 0.5 * (2 * np.arange(kmax_par + 1) + 1) * kF
        kper[sl] = 0.5 * (k_per + k_per + 1) * kF
    return kpar, kper


class Pk(object):
    """Pk_library.pyx:266-425."""

    def __init__(self, delta, BoxSize, axis=2, MAS="CIC", threads=1, keep_deltak=False):
        if delta.dtype != np.float32:
            raise ValueError("Buffer dtype mismatch, expected 'float32_t'")
        dims = len(delta)
        delta_k = FFT3Dr_f(delta, threads)
        o = _mode_loop([delta_k.view(np.float32)], dims, axis, [MAS_function(MAS)], BoxSize, True)
        kF, kN = o["kF"], o["kN"]
        fact = (BoxSize / dims ** 2) ** 3
        # 1-D  :387-394
        k1D, N1, P1 = o["k1d"][1:].copy(), o["n1d"][1:].copy(), o["p1d"][1:, 0].copy()
        P1 = P1 * fact
        k1D = (k1D / N1) * kF
        kmaxper = np.sqrt(kN ** 2 - k1D ** 2)
        P1 = P1 * (np.pi * kmaxper ** 2 / N1) / (2.0 * np.pi) ** 2
        self.k1D, self.Pk1D, self.Nmodes1D = k1D, P1, N1
        # 2-D  :397-407 (DC bin kept; an empty bin divides by zero like the reference)
        self.kpar, self.kper = _kpar_kper(o["kmax_par"], o["kmax_per"], kF)
        if np.any(o["n2d"] == 0):
            raise ZeroDivisionError("float division")
        self.Pk2D = o["p2d"][:, 0] * fact / o["n2d"]
        self.Nmodes2D = o["n2d"]
        # 3-D  :411-421
        check_number_modes(o["n3d"], dims)
        N3 = o["n3d"][1:].copy()
        self.k3D = (o["k3d"][1:] / N3) * kF
        P3 = o["p3d"][1:, :, 0].copy()
        P3[:, 0] = (P3[:, 0] / N3) * fact
        P3[:, 1] = (P3[:, 1] * 5.0 / N3) * fact
        P3[:, 2] = (P3[:, 2] * 9.0 / N3) * fact
        self.Pk, self.Nmodes3D = P3, N3
        self.Pkphase = (o["phase"][1:] / N3) * fact
        if keep_deltak:
            self.delta_k = delta_k


class XPk(object):
    """Pk_library.pyx:534-798."""

    def __init__(self, delta, BoxSize, axis=2, MAS=None, threads=1):
        dims = len(delta[0])
        for d in delta[1:]:
            if len(d) != dims:                   # :563-565
                print("Fields have different grid sizes!!!")
                sys.exit()
        dks = [FFT3Dr_f(d, threads) for d in delta]
        o = _mode_loop([dk.view(np.float32) for dk in dks], dims, axis, [MAS_function(m) for m in MAS], BoxSize, False)
        kF, kN = o["kF"], o["kN"]
        fact = (BoxSize / dims ** 2) ** 3
        # 1-D  :745-758
        N1 = o["n1d"][1:].copy()
        k1D = (o["k1d"][1:] / N1) * kF
        kmaxper = np.sqrt(kN ** 2 - k1D ** 2)
        s1 = (np.pi * kmaxper ** 2 / N1) / (2.0 * np.pi) ** 2
        self.k1D, self.Nmodes1D = k1D, N1
        self.Pk1D = o["p1d"][1:] * fact * s1[:, None]
        self.PkX1D = o["x1d"][1:] * fact * s1[:, None]
        # 2-D  :761-775
        self.kpar, self.kper = _kpar_kper(o["kmax_par"], o["kmax_per"], kF)
        if np.any(o["n2d"] == 0):
            raise ZeroDivisionError("float division")
        self.Nmodes2D = o["n2d"]
        self.Pk2D = o["p2d"] * fact / o["n2d"][:, None]
        self.PkX2D = o["x2d"] * fact / o["n2d"][:, None]
        # 3-D  :779-796
        check_number_modes(o["n3d"], dims)
        N3 = o["n3d"][1:].copy()
        self.k3D, self.Nmodes3D = (o["k3d"][1:] / N3) * kF, N3
        ell = np.array([1.0, 5.0, 9.0])[None, :, None]
        self.Pk = (o["p3d"][1:] * ell / N3[:, None, None]) * fact
        self.XPk = (o["x3d"][1:] * ell / N3[:, None, None]) * fact


class XPk_imag(XPk):
    """Pk_library.pyx:959-1225: class XPk with the imaginary cross term im_i*re_j - re_i*im_j (:1131-1132)."""

    def __init__(self, delta, BoxSize, axis=2, MAS=None, threads=1):
        lib().orc_set_cross_imag(1)
        try:
            XPk.__init__(self, delta, BoxSize, axis, MAS, threads)
        finally:
            lib().orc_set_cross_imag(0)


# --------------------------------------------------------------------------------------------------
# Siblings sharing the FFT and the mode loop, restated with vectorised numpy.  TEST INFRASTRUCTURE.
# Arithmetic notes follow the reference's C types: MAS_factor is a C float (double product rounded once), a
# complex64 times a float stays fp32 per component, |.|^2 and all sums are double.
# --------------------------------------------------------------------------------------------------
import scipy.fft as _sf


def _wavenumbers(dims):
    i = np.arange(dims)
    return np.where(i > dims // 2, i - dims, i)


def _mas_axis(dims, mas_index):
    """MAS_correction(pi*k/dims, MAS_index) for every FFT index, Pk_library.pyx:86-87."""
    x = (np.pi / dims) * _wavenumbers(dims).astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        c = np.where(x == 0.0, 1.0, (x / np.sin(x)) ** mas_index)
    return c


def _deconv(dk, mf):
    """complex64 *= C float, component-wise fp32 rounding (e.g. :355, :490)."""
    mf = mf.astype(np.float32)
    out = np.empty(dk.shape, np.complex64)
    out.real = dk.real * mf
    out.imag = dk.imag * mf
    return out


def frequencies_2D(BoxSize, dims):                   # :67-72
    kF = 2.0 * np.pi / BoxSize
    middle = dims // 2
    return kF, middle * kF, middle, middle, int(np.sqrt(middle ** 2 + middle ** 2))


def _plane_modes(grid):
    """(kx, ky) of the stored half plane and the mask of independent modes, :474-485."""
    middle = grid // 2
    kx = _wavenumbers(grid)[:, None] * np.ones((1, middle + 1), np.int64)
    ky = np.arange(middle + 1)[None, :] * np.ones((grid, 1), np.int64)
    special = (ky == 0) | ((ky == middle) & (grid % 2 == 0))
    keep = ~(special & (kx < 0))
    return kx, ky, keep


class Pk_plane(object):
    """Pk_library.pyx:440-516."""

    def __init__(self, delta, BoxSize, MAS="CIC", threads=1):
        grid = len(delta)
        kF, kN, kmax_par, kmax_per, kmax = frequencies_2D(BoxSize, grid)
        c = _mas_axis(grid, MAS_function(MAS))
        dk = _sf.rfftn(np.asarray(delta, np.float32), axes=(0, 1)).astype(np.complex64)
        kx, ky, keep = _plane_modes(grid)
        k = np.sqrt((kx * kx + ky * ky).astype(np.float64))
        dk = _deconv(dk, c[:, None] * c[None, :grid // 2 + 1])
        d2 = dk.real.astype(np.float64) ** 2 + dk.imag.astype(np.float64) ** 2
        idx = k.astype(np.int64)[keep]
        k2D = np.bincount(idx, k[keep], kmax + 1)
        Pk2D = np.bincount(idx, d2[keep], kmax + 1)
        Nmodes = np.bincount(idx, None, kmax + 1).astype(np.float64)
        self.k = (k2D[1:] / Nmodes[1:]) * kF
        self.Nmodes = Nmodes[1:]
        self.Pk = (Pk2D[1:] / Nmodes[1:]) * (BoxSize / grid ** 2) ** 2


class XPk_plane(object):
    """Pk_library.pyx:814-941."""

    def __init__(self, delta1, delta2, BoxSize, MAS1=None, MAS2=None, threads=1):
        grid = delta1.shape[0]
        kF, kN, kmax_par, kmax_per, kmax = frequencies_2D(BoxSize, grid)
        kx, ky, keep = _plane_modes(grid)
        k = np.sqrt((kx * kx + ky * ky).astype(np.float64))
        idx = k.astype(np.int64)[keep]
        dks = []
        for d, m in ((delta1, MAS1), (delta2, MAS2)):
            c = _mas_axis(grid, MAS_function(m))
            dk = _sf.rfftn(np.asarray(d, np.float32), axes=(0, 1)).astype(np.complex64)
            dks.append(_deconv(dk, c[:, None] * c[None, :grid // 2 + 1]))
        re = [d.real.astype(np.float64) for d in dks]
        im = [d.imag.astype(np.float64) for d in dks]
        Nmodes = np.bincount(idx, None, kmax + 1).astype(np.float64)[1:]
        fact = (BoxSize / grid ** 2) ** 3
        self.k = (np.bincount(idx, k[keep], kmax + 1)[1:] / Nmodes) * kF
        self.Nmodes = Nmodes
        self.Pk = np.stack([np.bincount(idx, (re[i] ** 2 + im[i] ** 2)[keep], kmax + 1)[1:] / Nmodes * fact
                            for i in range(2)], axis=1)
        self.XPk = np.bincount(idx, (re[0] * re[1] + im[0] * im[1])[keep], kmax + 1)[1:] / Nmodes * fact
        self.r = self.XPk / np.sqrt(self.Pk[:, 0] * self.Pk[:, 1])


def _cube_modes(dims):
    """(kx, ky, kz) of the stored half spectrum and the mask of independent modes, :326-330."""
    middle = dims // 2
    w = _wavenumbers(dims)
    kx = w[:, None, None] + np.zeros((1, dims, middle + 1), np.int64)
    ky = w[None, :, None] + np.zeros((dims, 1, middle + 1), np.int64)
    kz = np.arange(middle + 1)[None, None, :] + np.zeros((dims, dims, 1), np.int64)
    even = dims % 2 == 0
    plane = (kz == 0) | ((kz == middle) & even)
    drop = plane & ((kx < 0) | (((kx == 0) | ((kx == middle) & even)) & (ky < 0)))
    return kx, ky, kz, ~drop


def _mas_cube(dims, mas_index):
    c = _mas_axis(dims, mas_index)
    return (c[:, None, None] * c[None, :, None]) * c[None, None, :dims // 2 + 1]


def Pk_theta(Vx, Vy, Vz, BoxSize, axis=2, MAS="CIC", threads=1):
    """Pk_library.pyx:1245-1336."""
    dims = len(Vx)
    kF, kN, kmax_par, kmax_per, kmax = frequencies(BoxSize, dims)
    kx, ky, kz, keep = _cube_modes(dims)
    mf = _mas_cube(dims, MAS_function(MAS))
    V = [_deconv(FFT3Dr_f(np.asarray(v, np.float32)), mf) for v in (Vx, Vy, Vz)]
    f = [a.astype(np.float32) for a in (kx, ky, kz)]
    # int * float products and their sum are fp32 in the reference's C (:1308-1314)
    real = -(f[0] * V[0].imag + f[1] * V[1].imag + f[2] * V[2].imag)
    imag = f[0] * V[0].real + f[1] * V[1].real + f[2] * V[2].real
    theta2 = real.astype(np.float64) ** 2 + imag.astype(np.float64) ** 2
    kmod = np.sqrt((kx * kx + ky * ky + kz * kz).astype(np.float64))
    idx = kmod.astype(np.int64)[keep]
    k = np.bincount(idx, kmod[keep], kmax + 1)
    Pk_ = np.bincount(idx, theta2[keep], kmax + 1)
    Nmodes = np.bincount(idx, None, kmax + 1).astype(np.float64)
    check_number_modes(Nmodes, dims)
    k, Nmodes = k[1:], Nmodes[1:]
    k = (k / Nmodes) * kF
    Pk_ = Pk_[1:] * (BoxSize / dims ** 2) ** 3 * kF ** 2
    Pk_ *= (1.0 / Nmodes)
    return [k, Pk_, Nmodes]


def correct_MAS(delta, BoxSize, MAS="CIC", threads=1):
    """Pk_library.pyx:1749-1806.  The backward transform (pyfftw -> FFTW, un-vendored) is restated with pocketfft's
    irfftn, normalised by 1/dims^3 like pyfftw's FFTW.__call__ default (normalise_idft=True); it takes the real part
    after the complex passes, which fixes what the half-corrected self-conjugate planes mean."""
    dims = len(delta)
    kx, ky, kz, keep = _cube_modes(dims)
    dk = FFT3Dr_f(np.asarray(delta, np.float32))
    corrected = _deconv(dk, _mas_cube(dims, MAS_function(MAS)))
    dk = np.where(keep, corrected, dk)
    return _sf.irfftn(dk, s=(dims,) * 3, axes=(0, 1, 2)).astype(np.float32)


class Xi(object):
    """Pk_library.pyx:2035-2150."""

    def __init__(self, delta, BoxSize, MAS="CIC", axis=2, threads=1):
        BoxSize = float(np.float32(BoxSize))
        dims = delta.shape[0]
        kF, kN, kmax_par, kmax_per, kmax = frequencies(BoxSize, dims)
        dk = _deconv(FFT3Dr_f(np.asarray(delta, np.float32)), _mas_cube(dims, MAS_function(MAS)))
        p = np.zeros(dk.shape, np.complex64)
        p.real = dk.real * dk.real + dk.imag * dk.imag          # `float real, imag`, :2078-2082
        xi = _sf.irfftn(p, s=(dims,) * 3, axes=(0, 1, 2)).astype(np.float32)      # pyfftw normalises the inverse
        w = _wavenumbers(dims)
        kx = w[:, None, None] + np.zeros((1, dims, dims), np.int64)
        ky = w[None, :, None] + np.zeros((dims, 1, dims), np.int64)
        kz = w[None, None, :] + np.zeros((dims, dims, 1), np.int64)
        k = np.sqrt((kx * kx + ky * ky + kz * kz).astype(np.float64))
        kpar = (kx, ky, kz)[axis].astype(np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            mu = np.where(k == 0, 0.0, kpar / k)
        mu2 = mu * mu
        idx = k.astype(np.int64).ravel()
        x = xi.astype(np.float64).ravel()
        N = np.bincount(idx, None, kmax + 1).astype(np.float64)[1:]
        self.r3D = np.bincount(idx, k.ravel(), kmax + 1)[1:] / N * (BoxSize * 1.0 / dims)
        self.Nmodes3D = N
        norm = 1.0 / dims ** 3
        l2 = ((3.0 * mu2 - 1.0) / 2.0).ravel()
        l4 = ((35.0 * mu2 * mu2 - 30.0 * mu2 + 3.0) / 8.0).ravel()
        self.xi = np.stack([np.bincount(idx, x, kmax + 1)[1:] / N * norm,
                            np.bincount(idx, x * l2, kmax + 1)[1:] * 5.0 / N * norm,
                            np.bincount(idx, x * l4, kmax + 1)[1:] * 9.0 / N * norm], axis=1)


# ---------------------------------------------------------------------------------------------------------------
# Consumers of the FFT either side of the path: smoothing_library, void_library.gaussian_smoothing,
# bispectrum_library.Bk.  numpy restatements, pinned against the compiled reference by tests/golden/consumers.npz
# and tests/test_oracle_vs_reference.py.  The inverse transform is pyfftw's normalised one (see ref_shim/pyfftw.py).
# ---------------------------------------------------------------------------------------------------------------
def _ifft3(dk, dims):
    return _sf.irfftn(dk, s=(dims,) * 3, axes=(0, 1, 2)).astype(np.float32)


def FT_filter(BoxSize, R, dims, Filter, threads=1):
    """smoothing_library/smoothing_library.pyx:19-83."""
    if Filter not in ["Top-Hat", "Gaussian"]:
        raise Exception("Filter %s not implemented!" % Filter)
    R_grid = np.float32(np.float32(np.float32(R) * np.float32(dims)) / np.float32(BoxSize))      # C floats, :31
    R2 = np.float32(R_grid * R_grid)
    w = _wavenumbers(dims)
    d2 = (w[:, None, None] ** 2 + w[None, :, None] ** 2 + w[None, None, :] ** 2).astype(np.int64)
    if Filter == "Top-Hat":
        field = (d2.astype(np.float32) <= R2).astype(np.float32)                               # :47-50
    else:
        field = np.exp(-d2 / (2.0 * np.float64(R2))).astype(np.float32)                        # :66-67
    normalization = np.sum(field, dtype=np.float64)
    field = (field.astype(np.float64) / normalization).astype(np.float32)                      # :76-79
    return FFT3Dr_f(field)


def field_smoothing(field, filter_k, threads=1):
    """smoothing_library.pyx:89-114."""
    dims = field.shape[0]
    if dims != filter_k.shape[0]:
        raise Exception("field and filter have different grids!!!")
    fk = FFT3Dr_f(np.asarray(field, np.float32))
    return _ifft3((fk * np.asarray(filter_k, np.complex64)).astype(np.complex64), dims)


def gaussian_smoothing(delta, BoxSize, R, threads=1):
    """void_library/void_library.pyx:45-80 (a k-space top-hat, despite the name)."""
    dims = delta.shape[0]
    prefact = np.float32(float(np.float32(R)) * 2.0 * 3.141592653589793 / float(np.float32(BoxSize)))
    kx, ky, kz, _ = _cube_modes(dims)
    kR = (np.float64(prefact) * np.sqrt((kx * kx + ky * ky + kz * kz).astype(np.float64))).astype(np.float32)
    x = kR.astype(np.float64)
    kR3 = (kR * kR * kR).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        fact = np.where(np.abs(kR) < np.float32(1e-5), 1.0, 3.0 * (np.sin(x) - np.cos(x) * x) / kR3.astype(np.float64))
    fact = fact.astype(np.float32)
    fact[0, 0, 0] = 1.0                                                                        # DC mode skipped, :67-68
    return _ifft3(_deconv(FFT3Dr_f(np.asarray(delta, np.float32)), fact), dims)


class Bk(object):
    """Pk_library/bispectrum_library.pyx:32-196."""

    def __init__(self, delta, BoxSize, k1, k2, theta, MAS="CIC", threads=1):
        dims = len(delta)
        kF = frequencies(BoxSize, dims)[0]
        theta = np.asarray(theta, np.float64)
        bins = theta.shape[0]
        k3 = np.sqrt((k2 * np.sin(theta)) ** 2 + (k2 * np.cos(theta) + k1) ** 2)
        k_all = np.zeros(bins + 2)
        k_all[0], k_all[1], k_all[2:] = k1, k2, k3
        k_min, k_max = (k_all - kF) / kF, (k_all + kF) / kF
        # every stored mode is deconvolved (the skip rule is commented out, :103-108)
        dk = _deconv(FFT3Dr_f(np.asarray(delta, np.float32)), _mas_cube(dims, MAS_function(MAS)))
        kx, ky, kz, _ = _cube_modes(dims)
        k = np.sqrt((kx * kx + ky * ky + kz * kz).astype(np.float64))

        def shell(i):
            sel = (k >= k_min[i]) & (k < k_max[i])
            d = _ifft3(np.where(sel, dk, 0).astype(np.complex64), dims)
            ind = _ifft3(sel.astype(np.complex64), dims)
            P = np.sum((d * d).astype(np.float64)) / np.sum((ind * ind).astype(np.float64)) * (BoxSize / dims ** 2) ** 3
            return d, ind, P

        Pk = np.zeros(bins + 2)
        B, Q = np.zeros(bins), np.zeros(bins)
        d1, I1, Pk[0] = shell(0)
        d2, I2, Pk[1] = shell(1)
        for j in range(bins):
            d3, I3, Pk[j + 2] = shell(j + 2)
            num = np.sum(((d1 * d2) * d3).astype(np.float64))
            tri = np.sum(((I1 * I2) * I3).astype(np.float64))
            B[j] = (num / tri) * (BoxSize ** 2 / dims ** 3) ** 3
            Q[j] = B[j] / (Pk[0] * Pk[1] + Pk[0] * Pk[j + 2] + Pk[1] * Pk[j + 2])
        self.B, self.Q, self.k, self.Pk = B, Q, k_all, Pk
