"""Drop-in mirror of the reference's `Pk_library` for the FFT power-spectrum hot path.

    import Pk_library as PKL               # repo-root shim re-exports this module
    Pk  = PKL.Pk(delta, BoxSize, axis=2, MAS='CIC', threads=1)
    XPk = PKL.XPk([d1, d2], BoxSize, axis=2, MAS=['CIC', 'PCS'], threads=1)

Same constructor signatures, attribute names, dtypes, shapes and error behaviour as
library/Pk_library/Pk_library.pyx (class Pk :266-425, class XPk :534-798, frequencies :59-64,
MAS_function :75-81, check_number_modes :90-102, FFT3Dr_f :120-133).  The transform is cuFFT and the
mode loop is one fused CUDA kernel (pylians_b200/csrc/binning.cu); only the final bookkeeping on a few
KB of bins (units, (2l+1), /Nmodes, DC-bin handling, :387-421) runs here in numpy float64.
`threads` is accepted and ignored.  There is no CPU fallback.
"""
import ctypes
import sys
import time

import numpy as np
import torch

from . import _lib
from .MAS_library import _device, _is_torch, _dtype_name

VERBOSE = True          # the reference prints progress lines unconditionally; set False to silence
# binning algorithm override for tests / benchmarks: 0 auto, 1 generic (atomics), 2 ring (registers)
ALGO = _lib.BIN_AUTO
SWAP_AXES = True        # axis 0/1: swap the line of sight onto z in real space and use the ring kernel


def _say(msg):
    if VERBOSE:
        print(msg)


def frequencies(BoxSize, dims):
    """Pk_library.pyx:59-64 (python-2 integer division for `middle`)."""
    kF = 2.0 * np.pi / BoxSize
    middle = dims // 2
    kN = middle * kF
    kmax_par = middle
    kmax_per = int(np.sqrt(middle ** 2 + middle ** 2))
    kmax = int(np.sqrt(middle ** 2 + middle ** 2 + middle ** 2))
    return kF, kN, kmax_par, kmax_per, kmax


def MAS_function(MAS):
    """Pk_library.pyx:75-81: any unknown string (e.g. 'None') means no correction."""
    MAS_index = 0
    if MAS == "NGP": MAS_index = 1
    if MAS == "CIC": MAS_index = 2
    if MAS == "TSC": MAS_index = 3
    if MAS == "PCS": MAS_index = 4
    return MAS_index


def MAS_correction(x, MAS_index):
    """Pk_library.pyx:86-87."""
    return 1.0 if x == 0.0 else (x / np.sin(x)) ** MAS_index


def check_number_modes(Nmodes, dims):
    """Pk_library.pyx:90-102: abort (SystemExit) unless every independent mode was counted once."""
    own_modes = 1 if dims % 2 == 1 else 8
    repeated_modes = (dims ** 3 - own_modes) // 2
    indep_modes = repeated_modes + own_modes
    if int(np.sum(Nmodes)) != indep_modes:
        print("WARNING: Not all modes counted")
        print("Counted  %d independent modes" % (int(np.sum(Nmodes))))
        print("Expected %d independent modes" % indep_modes)
        sys.exit()


def _check_field(delta):
    if _is_torch(delta):
        ok = delta.dtype == torch.float32
    else:
        delta = np.asarray(delta) if not isinstance(delta, np.ndarray) else delta
        ok = delta.dtype == np.float32
    if not ok:
        raise ValueError("Buffer dtype mismatch, expected 'float32_t' but got '%s'" % _dtype_name(delta))
    if delta.ndim != 3:
        raise ValueError("Buffer has wrong number of dimensions (expected 3, got %d)" % delta.ndim)
    if not (delta.shape[0] == delta.shape[1] == delta.shape[2]):
        raise ValueError("delta must be a (dims,dims,dims) cube")
    return delta


def _work(nbytes, dev):
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=dev)


def _fft_field(lib, delta, dims, dev, stream, swap_axis=2, pad=False):
    """FFT3Dr_f on the device.  Returns a complex64 tensor (dims,dims,dims/2+1); `delta` is never modified.

    Host input: one strided H2D copy into the padded in-place layout, then an in-place R2C, so the
    device holds a single copy of the field.  Device input: out-of-place R2C into a fresh buffer.
    swap_axis 0|1: the real field is first axis-swapped (pylb_swap_axes) into the padded layout, so the
    requested line of sight becomes the half-spectrum axis the ring kernel is fast along.
    pad: give the k-space rows an even pitch (dims/2+2 complex when dims/2+1 is odd), so that every row starts
    on a 16-byte boundary; the result is then a strided view, not a contiguous tensor."""
    nz = dims // 2 + 1
    nzp = nz + (nz & 1) if pad else nz          # k-space row pitch in complex elements
    on_dev = _is_torch(delta) and delta.is_cuda

    def r2c(src_ptr, in_pitch, out_ptr):
        wb = lib.pylb_fft_r2c_pitched_work_bytes(dims, in_pitch, nzp)
        if wb == ctypes.c_size_t(-1).value:
            raise _lib.PylbError("cuFFT plan failed: " + lib.pylb_last_error().decode())
        work = _work(wb, dev)
        _lib.check(lib.pylb_fft_r2c_pitched(src_ptr, in_pitch, out_ptr, nzp, dims, work.data_ptr(), int(wb),
                                            stream.cuda_stream), "pylb_fft_r2c_pitched")

    def as_complex(buf):
        out = torch.view_as_complex(buf.view(dims, dims, nzp, 2))
        return out if nzp == nz else out[:, :, :nz]

    if swap_axis != 2:
        if on_dev:
            src = delta if delta.is_contiguous() else delta.contiguous()
        else:
            host = delta.contiguous() if _is_torch(delta) else np.ascontiguousarray(delta)
            src = (host if _is_torch(host) else torch.from_numpy(host)).to(dev, non_blocking=True)
        buf = torch.empty((dims, dims, 2 * nzp), dtype=torch.float32, device=dev)
        _lib.check(lib.pylb_swap_axes(src.data_ptr(), buf.data_ptr(), dims, int(swap_axis), 2 * nzp, stream.cuda_stream),
                   "pylb_swap_axes")
        r2c(buf.data_ptr(), 2 * nzp, buf.data_ptr())
        return as_complex(buf)
    if on_dev:
        src = delta if delta.is_contiguous() else delta.contiguous()
        buf = torch.empty((dims, dims, 2 * nzp), dtype=torch.float32, device=dev)
        r2c(src.data_ptr(), dims, buf.data_ptr())
        return as_complex(buf)
    host = delta.contiguous() if _is_torch(delta) else np.ascontiguousarray(delta)
    buf = torch.empty((dims, dims, 2 * nzp), dtype=torch.float32, device=dev)
    hptr = host.data_ptr() if _is_torch(host) else host.ctypes.data
    _lib.check(lib.pylb_h2d_pitched(hptr, buf.data_ptr(), dims, 2 * nzp, stream.cuda_stream), "pylb_h2d_pitched")
    r2c(buf.data_ptr(), 2 * nzp, buf.data_ptr())
    out = as_complex(buf)
    out._pylb_keepalive = host
    return out


def get_layout(dims, F):
    L = _lib.PkLayout()
    _lib.check(_lib.load().pylb_pk_get_layout(int(dims), int(F), ctypes.byref(L)), "pylb_pk_get_layout")
    return L


def bin_modes(delta_k, dims, axis, mas_index, want_phase, write_back, ks=None, sums=None, counts=None,
              accumulate=False, algo=None):
    """Run the fused binning kernel on a list of device complex64 k-space fields.

    Returns (layout, sums, counts): device float64 / int64 tensors of raw bin sums (see
    include/pylians_b200.h, pylb_pk_layout).  `ks` describes which part of k-space the tensors hold
    (default: the whole (dims,dims,dims/2+1) cube)."""
    lib = _lib.load()
    dev = delta_k[0].device
    F = len(delta_k)
    L = get_layout(dims, F)
    if ks is None:
        # the whole cube, rows along kz; strides in complex elements (a padded row pitch is fine)
        t0 = delta_k[0]
        if t0.stride(2) != 1 or any(t.stride() != t0.stride() for t in delta_k):
            raise ValueError("bin_modes: k-space fields must share strides and be contiguous along kz")
        ks = _lib.KSpace(dims, 0, dims, 0, dims, t0.stride(0), t0.stride(1))
    if sums is None:
        # one allocation for both accumulator arrays so that a single D2H copy brings them back
        raw = torch.empty(L.n_doubles + L.n_counts, dtype=torch.float64, device=dev)
        sums = raw[:L.n_doubles]
        counts = raw[L.n_doubles:].view(torch.int64)
        sums._pylb_raw = raw
    ptrs = (ctypes.c_void_p * F)(*[t.data_ptr() for t in delta_k])
    mi = (ctypes.c_int * F)(*[int(m) for m in mas_index])
    stream = torch.cuda.current_stream(dev)
    _lib.check(lib.pylb_pk_bin(ptrs, F, ctypes.byref(ks), int(axis), mi, int(want_phase), int(write_back),
                               ALGO if algo is None else algo, int(accumulate), sums.data_ptr(),
                               counts.data_ptr(), stream.cuda_stream), "pylb_pk_bin")
    return L, sums, counts


def _field_subsets(F, width=3):
    """Field subsets of `width` covering every pair (i, j), i < j < F, at least once (greedy)."""
    import itertools
    need = {(i, j) for i in range(F) for j in range(i + 1, F)}
    subsets = []
    while need:
        best = max(itertools.combinations(range(F), width), key=lambda c: len(need & set(itertools.combinations(c, 2))))
        subsets.append(best)
        need -= set(itertools.combinations(best, 2))
    return subsets


def bin_modes_by_subsets(delta_k, dims, axis, mas_index, width=3, ks=None):
    """The bins of F > 3 fields from the two- / three-field ring kernel: every auto and cross spectrum of F fields is an
    auto or cross spectrum of some subset of three of them, so the fields are binned three at a time (F = 4: three passes,
    F = 5: four) and the sums are copied into the F-field layout on the device.  The mode counts, sum |k| and the bin
    geometry do not depend on the fields.  Replaces the one-thread-per-mode kernel for XPk with more than three fields
    (Pk_Gadget on four particle types): it reads 8 B x 3 per mode and pass instead of taking a global atomic per mode."""
    F = len(delta_k)
    dev = delta_k[0].device
    L = get_layout(dims, F)
    raw = torch.zeros(L.n_doubles + L.n_counts, dtype=torch.float64, device=dev)
    sums, counts = raw[:L.n_doubles], raw[L.n_doubles:].view(torch.int64)
    sums._pylb_raw = raw
    n3, n1, B2 = L.kmax + 1, L.kmax_par + 1, L.B2
    pair = lambda i, j, nf: i * nf - i * (i + 1) // 2 + (j - i - 1)
    done_f, done_x, first = set(), set(), True
    for sub in _field_subsets(F, width):
        Ls, ss, cs = bin_modes([delta_k[f] for f in sub], dims, axis, [mas_index[f] for f in sub], False, False, ks=ks)
        w = len(sub)
        if first:
            counts[L.o_n3d:L.o_n3d + n3] = cs[Ls.o_n3d:Ls.o_n3d + n3]
            counts[L.o_n1d:L.o_n1d + n1] = cs[Ls.o_n1d:Ls.o_n1d + n1]
            counts[L.o_n2d:L.o_n2d + B2] = cs[Ls.o_n2d:Ls.o_n2d + B2]
            sums[L.o_k3d:L.o_k3d + n3] = ss[Ls.o_k3d:Ls.o_k3d + n3]
            first = False
        views = []
        for o_p, o_x, rows, mult in (("o_p3d", "o_x3d", n3 * 3, None), ("o_p1d", "o_x1d", n1, None), ("o_p2d", "o_x2d", B2, None)):
            P_full = sums[getattr(L, o_p):getattr(L, o_p) + rows * F].view(rows, F)
            X_full = sums[getattr(L, o_x):getattr(L, o_x) + rows * L.X].view(rows, L.X)
            P_sub = ss[getattr(Ls, o_p):getattr(Ls, o_p) + rows * w].view(rows, w)
            X_sub = ss[getattr(Ls, o_x):getattr(Ls, o_x) + rows * Ls.X].view(rows, Ls.X)
            views.append((P_full, X_full, P_sub, X_sub))
        for a, fa in enumerate(sub):
            if fa not in done_f:
                for P_full, _, P_sub, _ in views:
                    P_full[:, fa] = P_sub[:, a]
            for b in range(a + 1, w):
                fb = sub[b]
                if (fa, fb) not in done_x:
                    for _, X_full, _, X_sub in views:
                        X_full[:, pair(fa, fb, F)] = X_sub[:, pair(a, b, w)]
                    done_x.add((fa, fb))
        done_f.update(sub)
    return L, sums, counts


_KGRID = {}


def _kpar_kper(kmax_par, kmax_per, kF):
    """Bin centres of the 2-D table (Pk_library.pyx:397-403): pure geometry, cached per shape and handed out as
    read-only arrays (two 0.7 MB copies per call otherwise; a caller that wants to modify them copies)."""
    key = (kmax_par, kmax_per, kF)
    v = _KGRID.get(key)
    if v is None:
        i2 = np.arange((kmax_par + 1) * (kmax_per + 1))
        v = (0.5 * (2 * (i2 % (kmax_par + 1)) + 1) * kF, 0.5 * (2 * (i2 // (kmax_par + 1)) + 1) * kF)
        for a in v:
            a.flags.writeable = False
        _KGRID.clear()
        _KGRID[key] = v
    return v[0], v[1]


class _Bins(object):
    """Host view (numpy float64) of the raw sums laid out by pylb_pk_layout.

    With `fact` = (BoxSize/dims^2)^3 and device-resident sums (the normal case) the 2-D table is normalised and the
    mode counts are converted to float64 on the device first (pylb_pk_finish_tables), so the host makes exactly one
    private copy of the buffer and every attribute is a view of it."""

    def __init__(self, L, sums, counts, fact=None):
        raw = getattr(sums, "_pylb_raw", None)
        self.tables_done = False
        if raw is not None and raw.is_cuda:
            if fact is not None:
                _lib.check(_lib.load().pylb_pk_finish_tables(sums.data_ptr(), counts.data_ptr(), int(L.dims), int(L.F),
                                                             float(fact), torch.cuda.current_stream(raw.device).cuda_stream),
                           "pylb_pk_finish_tables")
                self.tables_done = True
            # single async copy into a pinned buffer of this call's own + one stream sync.  The buffer comes from torch's
            # caching host allocator (no cudaHostAlloc after the first calls) and IS the result: every attribute below is a
            # view of it, so the 24 MB of a 2048^3 bin table are not copied a second time on the host.
            host = torch.empty(raw.numel(), dtype=torch.float64, pin_memory=True)
            host.copy_(raw, non_blocking=True)
            torch.cuda.current_stream(raw.device).synchronize()
            h = host.numpy()
            if self.tables_done:
                s, c = h[:L.n_doubles], h[L.n_doubles:]
            else:
                s = h[:L.n_doubles]
                c = h[L.n_doubles:].view(np.int64).astype(np.float64)   # counts < 2^53: exact in float64
        else:
            s = sums.cpu().numpy()
            c = counts.cpu().numpy().astype(np.float64)
        F, X, n3, n1, B2 = L.F, L.X, L.kmax + 1, L.kmax_par + 1, L.B2
        self.F, self.X = F, X
        self.k3d = s[L.o_k3d:L.o_k3d + n3]
        self.p3d = s[L.o_p3d:L.o_p3d + n3 * 3 * F].reshape(n3, 3, F)
        self.x3d = s[L.o_x3d:L.o_x3d + n3 * 3 * X].reshape(n3, 3, X)
        self.phase = s[L.o_phase:L.o_phase + n3]
        self.p1d = s[L.o_p1d:L.o_p1d + n1 * F].reshape(n1, F)
        self.x1d = s[L.o_x1d:L.o_x1d + n1 * X].reshape(n1, X)
        self.p2d = s[L.o_p2d:L.o_p2d + B2 * F].reshape(B2, F)
        self.x2d = s[L.o_x2d:L.o_x2d + B2 * X].reshape(B2, X)
        self.n3d = c[L.o_n3d:L.o_n3d + n3]
        self.n1d = c[L.o_n1d:L.o_n1d + n1]
        self.n2d = c[L.o_n2d:L.o_n2d + B2]


def _finish(obj, b, dims, BoxSize, is_x):
    """Units and normalisation.  Pk: Pk_library.pyx:387-421;  XPk: :740-796."""
    kF, kN, kmax_par, kmax_per, kmax = frequencies(BoxSize, dims)
    fact = (BoxSize / dims ** 2) ** 3

    # 1-D: drop the DC bin; k1D accumulates k_par once per mode, i.e. k_par*Nmodes (:365)
    N1 = b.n1d[1:].copy()
    k1D = np.arange(1, kmax_par + 1, dtype=np.float64) * kF
    kmaxper = np.sqrt(kN ** 2 - k1D ** 2)
    s1 = (np.pi * kmaxper ** 2 / N1) / (2.0 * np.pi) ** 2
    obj.k1D, obj.Nmodes1D = k1D, N1
    P1 = b.p1d[1:] * fact * s1[:, None]
    X1 = b.x1d[1:] * fact * s1[:, None]

    # 2-D: the DC bin is kept; an empty bin is a ZeroDivisionError in the reference (cdivision False)
    if b.n2d.min() == 0:
        raise ZeroDivisionError("float division")
    obj.kpar, obj.kper = _kpar_kper(kmax_par, kmax_per, kF)
    obj.Nmodes2D = b.n2d            # already private to this call (see _Bins)
    if b.tables_done:               # normalised on the device: Pk2D = sum * (fact / Nmodes2D)
        P2, X2 = b.p2d, b.x2d
    else:
        inv2 = (fact / b.n2d)[:, None]
        P2 = b.p2d * inv2
        X2 = b.x2d * inv2

    # 3-D
    check_number_modes(b.n3d, dims)
    N3 = b.n3d[1:].copy()
    obj.k3D, obj.Nmodes3D = (b.k3d[1:] / N3) * kF, N3
    ell = np.array([1.0, 5.0, 9.0])[None, :, None]
    P3 = (b.p3d[1:] * ell / N3[:, None, None]) * fact
    X3 = (b.x3d[1:] * ell / N3[:, None, None]) * fact

    if is_x:
        obj.Pk1D, obj.PkX1D, obj.Pk2D, obj.PkX2D, obj.Pk, obj.XPk = P1, X1, P2, X2, P3, X3
    else:
        obj.Pk1D, obj.Pk = P1[:, 0].copy(), np.ascontiguousarray(P3[:, :, 0])
        obj.Pk2D = P2[:, 0] if b.tables_done else P2[:, 0].copy()     # F == 1: a contiguous view of the private copy
        obj.Pkphase = (b.phase[1:] / N3) * fact


class Pk(object):
    """1-D, 2-D and 3-D power spectrum (monopole, quadrupole, hexadecapole) of a density field.

    Attributes (all numpy float64, as in the reference): k3D, Pk[:,0..2], Nmodes3D, Pkphase,
    k1D, Pk1D, Nmodes1D, kpar, kper, Pk2D, Nmodes2D, and delta_k (complex64) when keep_deltak."""

    def __init__(self, delta, BoxSize, axis=2, MAS="CIC", threads=1, keep_deltak=False):
        start = time.time()
        _say("\nComputing power spectrum of the field...")
        delta = _check_field(delta)
        lib = _lib.load()
        dev = _device()
        dims = len(delta)
        stream = torch.cuda.current_stream(dev)
        # line of sight along x or y: swap that axis with z in real space and bin along z (same bins, same
        # mode counts; which member of each conjugate pair is kept differs, its |delta_k|^2 does not).
        # keep_deltak must return the reference's (kx,ky,kz>=0) layout: see below.
        can_swap = int(axis) in (0, 1) and (ALGO & 3) != _lib.BIN_GENERIC and SWAP_AXES
        if can_swap and keep_deltak:
            # The spectra come from the swapped field like without keep_deltak.  The array handed back is the transform of
            # the ORIGINAL field with every independent mode deconvolved: neither the MAS factor Cx*Cy*Cz nor the rule
            # which member of a conjugate pair is kept (Pk_library.pyx:326-330, in terms of the array axes) depends on the
            # line of sight, so a second transform plus the ring kernel's write-back pass along z produces it (its bins are
            # discarded).  14 ms instead of the 186 ms of the one-thread-per-mode kernel at 512^3.
            dk_s = _fft_field(lib, delta, dims, dev, stream, int(axis), pad=True)
            start2 = time.time()
            L, sums, counts = bin_modes([dk_s], dims, 2, [MAS_function(MAS)], True, False)
            del dk_s
            delta_k = _fft_field(lib, delta, dims, dev, stream, 2, pad=False)
            bin_modes([delta_k], dims, 2, [MAS_function(MAS)], False, True)
        else:
            swap = int(axis) if (can_swap and not keep_deltak) else 2
            delta_k = _fft_field(lib, delta, dims, dev, stream, swap, pad=not keep_deltak)
            start2 = time.time()
            L, sums, counts = bin_modes([delta_k], dims, 2 if swap != 2 else int(axis), [MAS_function(MAS)], True,
                                        bool(keep_deltak))
        bins = _Bins(L, sums, counts, (BoxSize / dims ** 2) ** 3)       # one D2H of the bins; synchronises the stream
        _say("Time to complete loop = %.2f" % (time.time() - start2))
        _finish(self, bins, dims, BoxSize, False)
        if keep_deltak:
            self.delta_k = delta_k if (_is_torch(delta) and delta.is_cuda) else delta_k.cpu().numpy()
        _say("Time taken = %.2f seconds" % (time.time() - start))


class XPk(object):
    """Auto- and cross-power spectra of several density fields.  Pk_library.pyx:534-798.

    Attributes: k3D, Nmodes3D, Pk[k, ell, field], XPk[k, ell, pair]; k1D, Nmodes1D, Pk1D, PkX1D;
    kpar, kper, Nmodes2D, Pk2D, PkX2D.  Pairs are ordered (0,1),(0,2),...,(1,2),..."""
    _ALGO_FLAGS = 0

    def __init__(self, delta, BoxSize, axis=2, MAS=None, threads=1):
        start = time.time()
        _say("\nComputing power spectra of the fields...")
        dims = len(delta[0])
        fields = len(delta)
        for i in range(1, fields):                      # :563-565
            if len(delta[i]) != dims:
                print("Fields have different grid sizes!!!")
                sys.exit()
        mas_index = [MAS_function(m) for m in list(MAS)]   # MAS=None -> TypeError, as in the reference (:573)
        if len(mas_index) < fields:
            raise IndexError("list index out of range")  # MAS[i] for i >= len(MAS), :577
        delta = [_check_field(d) for d in delta]
        lib = _lib.load()
        dev = _device()
        stream = torch.cuda.current_stream(dev)
        swap = int(axis) if (int(axis) in (0, 1) and (ALGO & 3) != _lib.BIN_GENERIC and SWAP_AXES
                             and not self._ALGO_FLAGS) else 2
        delta_k = [_fft_field(lib, d, dims, dev, stream, swap, pad=True) for d in delta]   # even row pitch: one aligned row table
        _say("Time FFTS = %.2f" % (time.time() - start))
        start2 = time.time()
        if fields > 3 and not self._ALGO_FLAGS and (ALGO & 3) != _lib.BIN_GENERIC and (swap != 2 or int(axis) == 2):
            L, sums, counts = bin_modes_by_subsets(delta_k, dims, 2, mas_index[:fields])      # three fields at a time
        else:
            L, sums, counts = bin_modes(delta_k, dims, 2 if swap != 2 else int(axis), mas_index[:fields], False, False,
                                        algo=ALGO | self._ALGO_FLAGS)
        bins = _Bins(L, sums, counts, (BoxSize / dims ** 2) ** 3)
        _say("Time loop = %.2f" % (time.time() - start2))
        _finish(self, bins, dims, BoxSize, True)
        _say("Time taken = %.2f seconds" % (time.time() - start))


class XPk_imag(XPk):
    """Real auto- and IMAGINARY cross-power spectra.  Pk_library.pyx:959-1225: class XPk with the cross term
    im_i*re_j - re_i*im_j (:1131-1132).  The imaginary part changes sign under k -> -k, so the line of sight is not
    moved onto z by an axis swap here (that keeps the other member of each conjugate pair)."""
    _ALGO_FLAGS = _lib.BIN_XIMAG


def FFT3Dr_f(a, threads=1):
    """Pk_library.pyx:120-133: unnormalised forward R2C, float32 -> complex64 (dims,dims,dims/2+1).
    numpy in -> numpy out; CUDA tensor in -> CUDA tensor out."""
    a = _check_field(a)
    lib = _lib.load()
    dev = _device()
    out = _fft_field(lib, a, len(a), dev, torch.cuda.current_stream(dev))
    return out if (_is_torch(a) and a.is_cuda) else out.cpu().numpy()


# --------------------------------------------------------------------------------------------------
# Siblings that share the FFT and the mode loop.  Same host conventions as Pk/XPk: numpy or
# CUDA-tensor inputs, float64 numpy spectra out, fields come back in the container they came in.
# --------------------------------------------------------------------------------------------------
def frequencies_2D(BoxSize, dims):
    """Pk_library.pyx:67-72."""
    kF = 2.0 * np.pi / BoxSize
    middle = dims // 2
    kN = middle * kF
    return kF, kN, middle, middle, int(np.sqrt(middle ** 2 + middle ** 2))


def check_number_modes_2D(Nmodes, dims):
    """Pk_library.pyx:105-118."""
    own_modes = 1 if dims % 2 == 1 else 4
    indep_modes = (dims ** 2 - own_modes) // 2 + own_modes
    if int(np.sum(Nmodes)) != indep_modes:
        print("WARNING: Not all modes counted")
        print("Counted  %d independent modes" % (int(np.sum(Nmodes))))
        print("Expected %d independent modes" % indep_modes)
        sys.exit()


def _dev_f32(a, ndim, dev):
    """float32 C-contiguous device tensor of a numpy array / tensor with `ndim` equal sides."""
    if _is_torch(a):
        if a.dtype != torch.float32:
            raise ValueError("Buffer dtype mismatch, expected 'float32_t' but got '%s'" % _dtype_name(a))
        t = a
    else:
        a = np.asarray(a)
        if a.dtype != np.float32:
            raise ValueError("Buffer dtype mismatch, expected 'float32_t' but got '%s'" % _dtype_name(a))
        t = torch.from_numpy(np.ascontiguousarray(a))
    if t.ndim != ndim:
        raise ValueError("Buffer has wrong number of dimensions (expected %d, got %d)" % (ndim, t.ndim))
    return t.to(dev, non_blocking=True).contiguous()


def _like_input(t, ref):
    return t if (_is_torch(ref) and ref.is_cuda) else t.cpu().numpy()


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def FFT2Dr_f(a, threads=1):
    """Pk_library.pyx:184-197: unnormalised forward R2C of a (grid,grid) float32 image -> complex64 (grid,grid/2+1)."""
    lib, dev = _lib.load(), _device()
    src = _dev_f32(a, 2, dev)
    n = src.shape[0]
    out = torch.empty((n, n // 2 + 1), dtype=torch.complex64, device=dev)
    _lib.check(lib.pylb_fft2d_r2c(src.data_ptr(), out.data_ptr(), n, _stream(dev)), "pylb_fft2d_r2c")
    return _like_input(out, a)


def _dev_c64(a, dev):
    t = a if _is_torch(a) else torch.from_numpy(np.ascontiguousarray(np.asarray(a)))
    if t.dtype != torch.complex64:
        raise ValueError("Buffer dtype mismatch, expected 'complex64_t' but got '%s'" % _dtype_name(a))
    return t.to(dev).contiguous().clone()          # the C2R transform destroys its input


def IFFT2Dr_f(a, threads=1):
    """Pk_library.pyx:216-229: backward C2R, complex64 (grid,grid/2+1) -> float32 (grid,grid), scaled by 1/grid^2 like
    the reference's pyfftw call (FFTW.__call__ normalises inverse transforms by default)."""
    lib, dev = _lib.load(), _device()
    src = _dev_c64(a, dev)
    n = src.shape[0]
    out = torch.empty((n, n), dtype=torch.float32, device=dev)
    _lib.check(lib.pylb_fft2d_c2r(src.data_ptr(), out.data_ptr(), n, 1, _stream(dev)), "pylb_fft2d_c2r")
    return _like_input(out, a)


def IFFT3Dr_f(a, threads=1):
    """Pk_library.pyx:152-165: backward C2R, complex64 (dims,dims,dims/2+1) -> float32 (dims,dims,dims), scaled by
    1/dims^3 like the reference's pyfftw call (FFTW.__call__ normalises inverse transforms by default)."""
    lib, dev = _lib.load(), _device()
    src = _dev_c64(a, dev)
    n = src.shape[0]
    out = torch.empty((n, n, n), dtype=torch.float32, device=dev)
    _lib.check(lib.pylb_fft_c2r(src.data_ptr(), out.data_ptr(), n, 1, _stream(dev)), "pylb_fft_c2r")
    return _like_input(out, a)


def _sums_host(sums):
    torch.cuda.current_stream(sums.device).synchronize()
    return sums.cpu().numpy()


class Pk_plane(object):
    """Power spectrum of a 2-D field (image).  Pk_library.pyx:440-516.  Attributes k, Nmodes, Pk."""

    def __init__(self, delta, BoxSize, MAS="CIC", threads=1):
        start = time.time()
        _say("\nComputing power spectrum of the field...")
        lib, dev = _lib.load(), _device()
        img = _dev_f32(delta, 2, dev)
        grid = img.shape[0]
        kF, kN, kmax_par, kmax_per, kmax = frequencies_2D(BoxSize, grid)
        dk = torch.empty((grid, grid // 2 + 1), dtype=torch.complex64, device=dev)
        _lib.check(lib.pylb_fft2d_r2c(img.data_ptr(), dk.data_ptr(), grid, _stream(dev)), "pylb_fft2d_r2c")
        sums = torch.empty((5, kmax + 1), dtype=torch.float64, device=dev)
        _lib.check(lib.pylb_plane_bin(dk.data_ptr(), None, grid, MAS_function(MAS), 0, sums.data_ptr(), _stream(dev)),
                   "pylb_plane_bin")
        s = _sums_host(sums)
        check_number_modes_2D(s[4], grid)
        Nmodes = s[4, 1:].copy()
        self.k = (s[0, 1:] / Nmodes) * kF
        self.Nmodes = Nmodes
        self.Pk = (s[1, 1:] / Nmodes) * (BoxSize / grid ** 2) ** 2          # :507
        _say("Time taken = %.2f seconds" % (time.time() - start))


class XPk_plane(object):
    """Auto- and cross-power spectrum of two images.  Pk_library.pyx:814-941.  Attributes k, Nmodes, Pk[:,2], XPk, r."""

    def __init__(self, delta1, delta2, BoxSize, MAS1=None, MAS2=None, threads=1):
        start = time.time()
        _say("\nComputing power spectra of the fields...")
        if delta1.shape[0] != delta2.shape[1]:                             # :839-840
            raise Exception("Images have different grid sizes!!!")
        lib, dev = _lib.load(), _device()
        a, b = _dev_f32(delta1, 2, dev), _dev_f32(delta2, 2, dev)
        grid = a.shape[0]
        kF, kN, kmax_par, kmax_per, kmax = frequencies_2D(BoxSize, grid)
        dk = torch.empty((2, grid, grid // 2 + 1), dtype=torch.complex64, device=dev)
        for i, t in enumerate((a, b)):
            _lib.check(lib.pylb_fft2d_r2c(t.data_ptr(), dk[i].data_ptr(), grid, _stream(dev)), "pylb_fft2d_r2c")
        sums = torch.empty((5, kmax + 1), dtype=torch.float64, device=dev)
        _lib.check(lib.pylb_plane_bin(dk[0].data_ptr(), dk[1].data_ptr(), grid, MAS_function(MAS1), MAS_function(MAS2),
                                      sums.data_ptr(), _stream(dev)), "pylb_plane_bin")
        s = _sums_host(sums)
        fact = (BoxSize / grid ** 2) ** 3                                  # :928 (sic: cubed for a 2-D field)
        Nmodes = s[4, 1:].copy()
        self.k = (s[0, 1:] / Nmodes) * kF
        self.Nmodes = Nmodes
        self.Pk = np.ascontiguousarray((s[1:3, 1:] / Nmodes).T * fact)
        self.XPk = (s[3, 1:] / Nmodes) * fact
        self.r = self.XPk / np.sqrt(self.Pk[:, 0] * self.Pk[:, 1])
        _say("Time taken = %.2f seconds" % (time.time() - start))


def _fft3(lib, field, dims, dev):
    return _fft_field(lib, _check_field(field), dims, dev, torch.cuda.current_stream(dev))


def Pk_theta(Vx, Vy, Vz, BoxSize, axis=2, MAS="CIC", threads=1):
    """Power spectrum of theta = div V.  Pk_library.pyx:1245-1336.  Returns [k, Pk, Nmodes]."""
    start = time.time()
    _say("Computing power spectrum of theta...")
    lib, dev = _lib.load(), _device()
    dims = len(Vx)
    kF, kN, kmax_par, kmax_per, kmax = frequencies(BoxSize, dims)
    vk = [_fft3(lib, v, dims, dev) for v in (Vx, Vy, Vz)]
    sums = torch.empty((3, kmax + 1), dtype=torch.float64, device=dev)
    _lib.check(lib.pylb_theta_bin(vk[0].data_ptr(), vk[1].data_ptr(), vk[2].data_ptr(), dims, MAS_function(MAS),
                                  sums.data_ptr(), _stream(dev)), "pylb_theta_bin")
    s = _sums_host(sums)
    check_number_modes(s[2], dims)
    Nmodes = s[2, 1:].copy()
    k = (s[0, 1:] / Nmodes) * kF
    Pk_ = s[1, 1:] * (BoxSize / dims ** 2) ** 3 * kF ** 2                  # :1331
    Pk_ *= (1.0 / Nmodes)
    _say("Time taken = %.2f seconds" % (time.time() - start))
    return [k, Pk_, Nmodes]


def correct_MAS(delta, BoxSize, MAS="CIC", threads=1):
    """Deconvolve the MAS window from a density field.  Pk_library.pyx:1749-1806.  Returns the corrected field, float32
    (dims,dims,dims) (the reference's IFFT3Dr_f is pyfftw's normalised inverse)."""
    start = time.time()
    _say("\nComputing power spectrum of the field...")
    lib, dev = _lib.load(), _device()
    delta = _check_field(delta)
    dims = len(delta)
    dk = _fft_field(lib, delta, dims, dev, torch.cuda.current_stream(dev))      # a fresh buffer, never `delta` itself
    _lib.check(lib.pylb_mas_correct(dk.data_ptr(), dims, MAS_function(MAS), 0, _stream(dev)), "pylb_mas_correct")
    out = torch.empty((dims, dims, dims), dtype=torch.float32, device=dev)
    _lib.check(lib.pylb_fft_c2r(dk.data_ptr(), out.data_ptr(), dims, 1, _stream(dev)), "pylb_fft_c2r")
    _say("Time taken = %.2f seconds" % (time.time() - start))
    return _like_input(out, delta)


class Xi(object):
    """Correlation function multipoles from the inverse transform of |delta_k|^2.  Pk_library.pyx:2035-2150.
    Attributes r3D, xi[:,0..2] (l = 0, 2, 4), Nmodes3D."""

    def __init__(self, delta, BoxSize, MAS="CIC", axis=2, threads=1):
        start = time.time()
        _say("\nComputing correlation function of the field...")
        lib, dev = _lib.load(), _device()
        delta = _check_field(delta)
        dims = delta.shape[0]
        BoxSize = float(np.float32(BoxSize))                               # `float BoxSize` in the signature
        kF, kN, kmax_par, kmax_per, kmax = frequencies(BoxSize, dims)
        dk = _fft_field(lib, delta, dims, dev, torch.cuda.current_stream(dev))
        _lib.check(lib.pylb_mas_correct(dk.data_ptr(), dims, MAS_function(MAS), 1, _stream(dev)), "pylb_mas_correct")
        xi = torch.empty((dims, dims, dims), dtype=torch.float32, device=dev)
        _lib.check(lib.pylb_fft_c2r(dk.data_ptr(), xi.data_ptr(), dims, 1, _stream(dev)), "pylb_fft_c2r")
        del dk
        sums = torch.empty((5, kmax + 1), dtype=torch.float64, device=dev)
        _lib.check(lib.pylb_xi_bin(xi.data_ptr(), dims, int(axis), sums.data_ptr(), _stream(dev)), "pylb_xi_bin")
        s = _sums_host(sums)
        Nmodes = s[4, 1:].copy()
        norm = 1.0 / dims ** 3
        self.r3D = (s[0, 1:] / Nmodes) * (BoxSize * 1.0 / dims)           # :2139
        self.Nmodes3D = Nmodes
        self.xi = np.ascontiguousarray(np.stack([(s[1, 1:] / Nmodes) * norm, (s[2, 1:] * 5.0 / Nmodes) * norm,
                                                 (s[3, 1:] * 9.0 / Nmodes) * norm], axis=1))
        _say("Time taken = %.2f seconds" % (time.time() - start))
