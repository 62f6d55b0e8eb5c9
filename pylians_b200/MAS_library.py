"""Drop-in mirror of the reference's `MAS_library` for the mass-assignment hot path.

    import MAS_library as MASL            # repo-root shim re-exports this module
    MASL.MA(pos, delta, BoxSize, MAS='CIC', W=None, verbose=False, renormalize_2D=True)

Same names, argument meaning and error behaviour as library/MAS_library/MAS_library.pyx:57-112
(and the MAS_c shims :1136-1220), but the work is done by the sm_100a kernels in
pylians_b200/csrc through the C ABI (include/pylians_b200.h).  Inputs may be numpy arrays (staged
host<->device here, so the in-place `number +=` contract survives) or torch CUDA tensors (zero copy).
There is no CPU fallback.
"""
import sys
import time

import numpy as np
import torch

from . import _lib

_MAS_ID = {"NGP": 0, "CIC": 1, "TSC": 2, "PCS": 3}
_SUPPORT = {"NGP": 1, "CIC": 2, "TSC": 3, "PCS": 4}

# deposit algorithm override for tests / benchmarks: 0 auto, 1 direct (global red), 2 tiled (shared memory)
ALGO = _lib.MA_AUTO


def FLOAT_type():
    """MAS_library.pyx:11-13 (typedef float FLOAT, MAS_c.h:1)."""
    return np.float32


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("pylians_b200 needs a CUDA device (B200); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _is_torch(a):
    return isinstance(a, torch.Tensor)


def _dtype_name(a):
    return str(a.dtype).replace("torch.", "")


def _require_f32(a, what, ndim):
    if (_is_torch(a) and a.dtype != torch.float32) or (not _is_torch(a) and a.dtype != np.float32):
        # Cython typed-memoryview error of the reference
        raise ValueError("Buffer dtype mismatch, expected 'float32_t' but got '%s' (%s)" % (_dtype_name(a), what))
    if a.ndim != ndim:
        raise ValueError("Buffer has wrong number of dimensions (expected %d, got %d) (%s)" % (ndim, a.ndim, what))


def _to_device(a, dev):
    """numpy / CPU tensor -> device tensor (strides preserved where torch can); CUDA tensors pass through."""
    if _is_torch(a):
        return a if a.is_cuda else a.to(dev, non_blocking=True)
    t = torch.from_numpy(a) if a.flags.writeable else torch.from_numpy(a.copy())
    return t.to(dev, non_blocking=True)


HOST_CHUNK = None         # tests: force a (small) chunk size
HOST_TAPER_FLOOR = 1 << 24  # smallest piece the last H2D chunk is cut into (tests lower it)


def _host_chunk(dims, ndim):
    """Particles per H2D chunk when `pos` lives on the host.  Every chunk is one deposit, and a deposit flushes every tile
    it touches, so a chunk should not be sparse on the grid (a quarter of a particle per cell at least); 2^26 (805 MB of
    positions) is enough to hide the fixed costs on small grids, 2^28 (3.2 GB, double-buffered) bounds the staging."""
    if HOST_CHUNK is not None:
        return int(HOST_CHUNK)
    return int(min(max(dims ** ndim // 4, 1 << 26), 1 << 28))


def _chunk_bounds(npart, chunk, taper_floor=None):
    """(lo, hi) ranges of the H2D chunks.  The copies run back to back and the deposit of a chunk is faster than its
    copy, so what the caller waits for after the LAST copy has landed is the deposit of the last chunk: the final full
    chunk is therefore cut into 1/2, 1/4, 1/4 (never below `taper_floor` particles)."""
    if taper_floor is None:
        taper_floor = HOST_TAPER_FLOOR
    bounds, lo = [], 0
    while lo < npart:
        hi = min(npart, lo + chunk)
        bounds.append((lo, hi))
        lo = hi
    if len(bounds) >= 2:
        lo, hi = bounds.pop()
        n = hi - lo
        cuts = [lo]
        if n // 4 >= taper_floor:
            cuts += [lo + n // 2, lo + n // 2 + n // 4]
        elif n // 2 >= taper_floor:
            cuts += [lo + n // 2]
        cuts.append(hi)
        bounds += [(a, b) for a, b in zip(cuts[:-1], cuts[1:])]
    return bounds


_COPY_STREAMS = {}


def _copy_stream(dev):
    s = _COPY_STREAMS.get(dev.index)
    if s is None:
        s = _COPY_STREAMS[dev.index] = torch.cuda.Stream(device=dev)
    return s


def _as_cpu_tensor(a):
    if _is_torch(a):
        return a
    return torch.from_numpy(a) if a.flags.writeable else torch.from_numpy(a.copy())


def _launch_ma(lib, d_pos, d_w, d_grid, ndim, dims, BoxSize, mas, z_repeat, grid_f64, algo, stream):
    npart = d_pos.shape[0]
    ws_bytes = lib.pylb_ma_workspace_bytes(npart, ndim, dims, mas, int(d_w is not None), int(grid_f64), algo)
    ws = torch.empty(max(int(ws_bytes), 1), dtype=torch.uint8, device=d_grid.device)
    s0, s1 = d_pos.stride()
    _lib.check(lib.pylb_ma(d_pos.data_ptr(), npart, ndim, s0, s1, d_grid.data_ptr(), int(grid_f64), dims,
                           float(BoxSize), mas, d_w.data_ptr() if d_w is not None else None, int(z_repeat),
                           algo, ws.data_ptr(), int(ws_bytes), stream.cuda_stream), "pylb_ma")


def _deposit(pos, number, BoxSize, mas, W, z_repeat, grid_f64=False, algo=None):
    """Common device path: accumulates into `number` (numpy or tensor) in place.

    Device-resident `pos`: one pylb_ma call.  Host `pos`: the particle array is streamed in chunks on a
    side stream, double-buffered, so the H2D copy of chunk i+1 overlaps the deposit of chunk i (MA only
    ever adds into the grid, so chunking changes nothing but the fp32 summation order)."""
    lib = _lib.load()
    dev = _device()
    ndim = pos.shape[1]
    dims = number.shape[0]
    npart = pos.shape[0]
    host_grid = not (_is_torch(number) and number.is_cuda)
    if _is_torch(number):
        if not number.is_contiguous():
            raise ValueError("number must be C-contiguous")
    elif not (number.flags["C_CONTIGUOUS"] and number.flags["WRITEABLE"]):
        raise ValueError("number must be a writeable C-contiguous array")
    stream = torch.cuda.current_stream(dev)
    algo = ALGO if algo is None else algo
    d_grid = _to_device(number, dev) if host_grid else number
    pos_on_dev = _is_torch(pos) and pos.is_cuda
    chunk = _host_chunk(dims, ndim)
    if pos_on_dev or npart <= chunk:
        d_pos = _to_device(pos, dev)
        d_w = None
        if W is not None:
            d_w = _to_device(W, dev)
            if not d_w.is_contiguous():
                d_w = d_w.contiguous()
        _launch_ma(lib, d_pos, d_w, d_grid, ndim, dims, BoxSize, mas, z_repeat, grid_f64, algo, stream)
        if not host_grid and (d_pos is not pos or (W is not None and d_w is not W)):
            stream.synchronize()    # host inputs were staged asynchronously: do not return before they are consumed
        return d_grid, host_grid, stream

    # ---- host particles: chunked, double-buffered H2D on a side stream ----------------------------
    h_pos = _as_cpu_tensor(pos)
    h_w = None
    if W is not None:
        h_w = _as_cpu_tensor(W) if not (_is_torch(W) and W.is_cuda) else None
        d_w_full = W if h_w is None else None
    cs = _copy_stream(dev)
    bufs = [torch.empty((chunk, ndim), dtype=torch.float32, device=dev) for _ in range(2)]
    wbufs = [torch.empty(chunk, dtype=torch.float32, device=dev) for _ in range(2)] if h_w is not None else None
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    cs.wait_stream(stream)
    bounds = _chunk_bounds(npart, chunk)
    nchunks = len(bounds)

    def issue_copy(i):
        b = i % 2
        lo, hi = bounds[i]
        with torch.cuda.stream(cs):
            if i >= 2:
                cs.wait_event(consumed[b])
            bufs[b][: hi - lo].copy_(h_pos[lo:hi], non_blocking=True)
            if wbufs is not None:
                wbufs[b][: hi - lo].copy_(h_w[lo:hi], non_blocking=True)
            copied[b].record(cs)

    issue_copy(0)
    for i in range(nchunks):
        b = i % 2
        lo, hi = bounds[i]
        if i + 1 < nchunks:
            issue_copy(i + 1)
        stream.wait_event(copied[b])
        d_w = None
        if W is not None:
            d_w = wbufs[b][: hi - lo] if wbufs is not None else d_w_full[lo:hi]
        _launch_ma(lib, bufs[b][: hi - lo], d_w, d_grid, ndim, dims, BoxSize, mas, z_repeat, grid_f64, algo, stream)
        consumed[b].record(stream)
    for t in bufs + (wbufs or []):
        t.record_stream(stream)
    if not host_grid:
        stream.synchronize()
    return d_grid, host_grid, stream


def _write_back(d_grid, number, stream):
    if _is_torch(number):
        number.copy_(d_grid)
    else:
        torch.from_numpy(number).copy_(d_grid)     # D2H on the current stream, synchronous for pageable hosts
    stream.synchronize()


def MA(pos, number, BoxSize, MAS="CIC", W=None, verbose=False, renormalize_2D=True):
    """Mass assignment of particles onto a 2-D or 3-D grid, in place.  MAS_library.pyx:57-112."""
    coord, coord_aux = pos.shape[1], number.ndim
    if coord != coord_aux:                                   # :64-66
        print("pos have %d dimensions and the density %d!!!" % (coord, coord_aux))
        sys.exit()
    if verbose:                                              # :68-70
        print("\nUsing %s mass assignment scheme%s" % (MAS, "" if W is None else " with weights"))
    start = time.perf_counter()
    if coord not in (2, 3):
        return
    if MAS not in _MAS_ID:                                   # :81-82, :108-109
        print("option not valid!!!")
        sys.exit()
    _require_f32(pos, "pos", 2)
    _require_f32(number, "number", coord)
    if W is not None:
        _require_f32(W, "W", 1)
    lib = _lib.load()
    # 2-D: the Cython kernels repeat every update 2/3/4 times along the dummy axis, then the WHOLE
    # array (including what it held before) is divided  (:84-110)
    zrep = _SUPPORT[MAS] if coord == 2 else 1
    d_grid, host_grid, stream = _deposit(pos, number, BoxSize, _MAS_ID[MAS], W, zrep)
    if coord == 2 and renormalize_2D and MAS != "NGP":
        _lib.check(lib.pylb_divide(d_grid.data_ptr(), d_grid.numel(), float(_SUPPORT[MAS]), stream.cuda_stream),
                   "pylb_divide")
    if host_grid:
        _write_back(d_grid, number, stream)
    if verbose:
        torch.cuda.synchronize()
        print("Time taken = %.3f seconds\n" % (time.perf_counter() - start))


# ---- float64-grid variants: NGPW_d :338-358, CICW_d :229-266 ------------------------------------
def _ma_d(pos, number, BoxSize, W, mas):
    _require_f32(pos, "pos", 2)
    _require_f32(W, "W", 1)
    if (_is_torch(number) and number.dtype != torch.float64) or (not _is_torch(number) and number.dtype != np.float64):
        raise ValueError("Buffer dtype mismatch, expected 'float64_t' but got '%s'" % _dtype_name(number))
    d_grid, host_grid, stream = _deposit(pos, number, BoxSize, mas, W, 1, grid_f64=True, algo=_lib.MA_DIRECT)
    if host_grid:
        _write_back(d_grid, number, stream)


def NGPW_d(pos, number, BoxSize, W):
    _ma_d(pos, number, BoxSize, W, 0)


def CICW_d(pos, number, BoxSize, W):
    _ma_d(pos, number, BoxSize, W, 1)


# ---- direct kernel entry points (no dispatch): NGP :273, CIC :123, TSC :369, PCS :463 and W variants
def _direct(mas):
    def plain(pos, number, BoxSize):
        _require_f32(pos, "pos", 2); _require_f32(number, "number", 3)
        zrep = _SUPPORT[mas] if pos.shape[1] == 2 else 1
        num3 = number
        d_grid, host_grid, stream = _deposit(pos, num3[..., 0] if pos.shape[1] == 2 else num3, BoxSize, _MAS_ID[mas], None, zrep)
        if host_grid:
            _write_back(d_grid, num3[..., 0] if pos.shape[1] == 2 else num3, stream)

    def weighted(pos, number, BoxSize, W):
        _require_f32(pos, "pos", 2); _require_f32(number, "number", 3); _require_f32(W, "W", 1)
        zrep = _SUPPORT[mas] if pos.shape[1] == 2 else 1
        num3 = number
        d_grid, host_grid, stream = _deposit(pos, num3[..., 0] if pos.shape[1] == 2 else num3, BoxSize, _MAS_ID[mas], W, zrep)
        if host_grid:
            _write_back(d_grid, num3[..., 0] if pos.shape[1] == 2 else num3, stream)
    plain.__name__, weighted.__name__ = mas, mas + "W"
    return plain, weighted


NGP, NGPW = _direct("NGP")
CIC, CICW = _direct("CIC")
TSC, TSCW = _direct("TSC")
PCS, PCSW = _direct("PCS")


# ---- MAS_c (OpenMP) shims, MAS_library.pyx:1136-1220: C-contiguous only, single update per 2-D cell,
#      no renormalisation; `threads` is accepted and ignored -----------------------------------------
def _masc(mas, ndim, weighted):
    def _check_contig(a, what):
        ok = a.is_contiguous() if _is_torch(a) else a.flags["C_CONTIGUOUS"]
        if not ok:
            raise ValueError("ndarray is not C-contiguous (%s)" % what)

    if weighted:
        def fn(pos, number, W, BoxSize, threads=1):
            _require_f32(pos, "pos", 2); _require_f32(number, "number", ndim); _require_f32(W, "W", 1)
            _check_contig(pos, "pos"); _check_contig(number, "number"); _check_contig(W, "W")
            d_grid, host_grid, stream = _deposit(pos, number, BoxSize, _MAS_ID[mas], W, 1)
            if host_grid:
                _write_back(d_grid, number, stream)
    else:
        def fn(pos, number, BoxSize, threads=1):
            _require_f32(pos, "pos", 2); _require_f32(number, "number", ndim)
            _check_contig(pos, "pos"); _check_contig(number, "number")
            d_grid, host_grid, stream = _deposit(pos, number, BoxSize, _MAS_ID[mas], None, 1)
            if host_grid:
                _write_back(d_grid, number, stream)
    fn.__name__ = "%s%sc%dD" % (mas, "W" if weighted else "", ndim)
    return fn


for _m in ("NGP", "CIC", "TSC", "PCS"):
    for _d in (2, 3):
        for _w in (False, True):
            _f = _masc(_m, _d, _w)
            globals()[_f.__name__] = _f
del _m, _d, _w, _f


def overdensity(delta):
    """In-place `delta /= mean(delta); delta -= 1` (what every caller does between MA and Pk, e.g.
    Pk_snapshot.py:88,194) as two bandwidth-bound kernels; mean in float64.  numpy or CUDA tensor."""
    _require_f32(delta, "delta", delta.ndim)
    lib = _lib.load()
    dev = _device()
    stream = torch.cuda.current_stream(dev)
    host = not (_is_torch(delta) and delta.is_cuda)
    d = _to_device(delta, dev) if host else delta
    if not d.is_contiguous():
        raise ValueError("delta must be C-contiguous")
    scratch = torch.empty(2, dtype=torch.float64, device=dev)
    _lib.check(lib.pylb_overdensity(d.data_ptr(), d.numel(), scratch.data_ptr(), stream.cuda_stream), "pylb_overdensity")
    if host:
        _write_back(d, delta, stream)
