"""ctypes binding of libpylians_b200.so (the C ABI declared in include/pylians_b200.h).

There is no CPU fallback: importing this module fails loudly when the CUDA library has not been
built (`python -m pylians_b200.build`), and every compute call fails when no CUDA device exists.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "lib", "libpylians_b200.so")

c_void_p, c_int, c_int64, c_size_t, c_float, c_double = (
    ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t, ctypes.c_float, ctypes.c_double)


class PkLayout(ctypes.Structure):
    _fields_ = [(n, c_int) for n in ("dims", "F", "X", "middle", "kmax_par", "kmax_per", "kmax")] + \
               [(n, c_int64) for n in ("B2", "o_k3d", "o_p3d", "o_x3d", "o_phase", "o_p1d", "o_x1d", "o_p2d",
                                        "o_x2d", "n_doubles", "o_n3d", "o_n1d", "o_n2d", "n_counts")]


class KSpace(ctypes.Structure):
    _fields_ = [(n, c_int) for n in ("dims", "x0", "nx", "y0", "ny")] + \
               [("stride_x", c_int64), ("stride_y", c_int64)]


# every symbol include/pylians_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "pylb_version": (c_int, []),
    "pylb_last_error": (ctypes.c_char_p, []),
    "pylb_launch_count": (c_int64, []),
    "pylb_timing_enable": (None, [c_int]),
    "pylb_timing_collect": (c_int, [c_int, ctypes.POINTER(c_double), ctypes.POINTER(c_int)]),
    "NGP": (None, [c_void_p, c_void_p, c_void_p, ctypes.c_long, c_int, c_int, c_float, c_int]),
    "CIC": (None, [c_void_p, c_void_p, c_void_p, ctypes.c_long, c_int, c_int, c_float, c_int]),
    "TSC": (None, [c_void_p, c_void_p, c_void_p, ctypes.c_long, c_int, c_int, c_float, c_int]),
    "PCS": (None, [c_void_p, c_void_p, c_void_p, ctypes.c_long, c_int, c_int, c_float, c_int]),
    "pylb_ma_workspace_bytes": (c_size_t, [c_int64, c_int, c_int, c_int, c_int, c_int, c_int]),
    "pylb_ma": (c_int, [c_void_p, c_int64, c_int, c_int64, c_int64, c_void_p, c_int, c_int, c_float, c_int,
                        c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "pylb_ma_window_workspace_bytes": (c_size_t, [c_int64, c_int, c_int, c_int, c_int]),
    "pylb_ma_window": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int, c_int, c_int, c_float, c_int,
                               c_void_p, c_int64, c_int, c_void_p, c_size_t, c_void_p]),
    "pylb_partition_xslab": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int, c_float, c_int,
                                     c_int, c_void_p, c_void_p, c_void_p]),
    "pylb_add_f32": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "pylb_ma_debug_path": (None, [c_int]),
    "pylb_divide": (c_int, [c_void_p, c_int64, c_float, c_void_p]),
    "pylb_h2d_padded": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "pylb_h2d_pitched": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p]),
    "pylb_overdensity": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "pylb_grid_sum": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "pylb_overdensity_apply": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "pylb_pos_redshift_space": (c_int, [c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_int, c_void_p]),
    "pylb_swap_axes": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p]),
    "pylb_fft_r2c_work_bytes": (c_size_t, [c_int, c_int]),
    "pylb_fft_r2c": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "pylb_fft_r2c_pitched_work_bytes": (c_size_t, [c_int, c_int64, c_int64]),
    "pylb_fft_r2c_pitched": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_size_t, c_void_p]),
    "pylb_fft_c2r": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "pylb_fft2d_r2c": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "pylb_fft2d_c2r": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "pylb_mas_correct": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p]),
    "pylb_theta_bin": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "pylb_plane_bin": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pylb_xi_bin": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "pylb_fft_slab_yz_work_bytes": (c_size_t, [c_int, c_int, c_int64]),
    "pylb_fft_slab_yz": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p, c_size_t, c_void_p]),
    "pylb_fft_slab_x_work_bytes": (c_size_t, [c_int, c_int, c_int64]),
    "pylb_fft_slab_x": (c_int, [c_void_p, c_int, c_int, c_int64, c_void_p, c_size_t, c_void_p]),
    "pylb_slab_pack": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_void_p]),
    "pylb_pk_finish_tables": (c_int, [c_void_p, c_void_p, c_int, c_int, c_double, c_void_p]),
    "pylb_scale_f32": (c_int, [c_void_p, c_int64, c_float, c_void_p]),
    "pylb_overdensity_mean": (c_int, [c_void_p, c_int64, c_float, c_void_p]),
    "pylb_axpy_f32": (c_int, [c_void_p, c_void_p, c_float, c_int64, c_void_p]),
    "pylb_filter_real": (c_int, [c_void_p, c_int, c_float, c_int, c_void_p, c_void_p]),
    "pylb_cmul_c64": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "pylb_tophat_k": (c_int, [c_void_p, c_int, c_float, c_void_p]),
    "pylb_bk_shell": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_double, c_double, c_void_p]),
    "pylb_prod_sum": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "pylb_pk_get_layout": (c_int, [c_int, c_int, ctypes.POINTER(PkLayout)]),
    "pylb_pk_bin": (c_int, [ctypes.POINTER(c_void_p), c_int, ctypes.POINTER(KSpace), c_int, ctypes.POINTER(c_int),
                            c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
}

MA_AUTO, MA_DIRECT, MA_TILED = 0, 1, 2
BIN_AUTO, BIN_GENERIC, BIN_RING, BIN_PRECISE, BIN_BULK, BIN_RING1, BIN_XIMAG = 0, 1, 2, 16, 32, 64, 256

_lib = None


class PylbError(RuntimeError):
    pass


def load():
    """Load the shared library once; raise ImportError with build instructions if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            "pylians_b200: CUDA library %s not found. Build it with `python -m pylians_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback." % SO_PATH)
    lib = ctypes.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header and library out of sync
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().pylb_last_error()
        raise PylbError("%s failed: %s" % (what, msg.decode("utf-8", "replace") if msg else "unknown error"))


def launch_count():
    return int(load().pylb_launch_count())


T_RING, T_TILE, T_DIRECT, T_GENERIC, T_BIN, T_FFT, T_SORT = 0, 1, 2, 3, 4, 5, 6


def timing_enable(on):
    load().pylb_timing_enable(int(bool(on)))


def timing_collect(which):
    """(total_ms, launches) of kernel `which` since the last collect (CUDA events on its stream)."""
    ms, n = c_double(0), c_int(0)
    check(load().pylb_timing_collect(int(which), ctypes.byref(ms), ctypes.byref(n)), "pylb_timing_collect")
    return ms.value, n.value
