"""Load the compiled, unmodified reference modules from oracle/_ref.  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / --impl reference legs may
import this.  Returns (MASL, PKL) or raises ImportError when oracle/_ref was never built.
The modules are loaded under private names so they can never shadow the product's
`MAS_library` / `Pk_library` drop-in modules.
"""
import glob
import importlib.util
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
_CACHE = {}


def available():
    d = os.path.join(HERE, "_ref")
    return bool(glob.glob(os.path.join(d, "MAS_library*.so"))) and bool(glob.glob(os.path.join(d, "Pk_library*.so")))


def _load(name):
    cands = glob.glob(os.path.join(HERE, "_ref", name + "*.so"))
    if not cands:
        raise ImportError("oracle/_ref/%s*.so not built (run python oracle/build_ref.py where /root/reference exists)" % name)
    # the extension's init symbol is PyInit_<name>; register it under that name only while loading
    spec = importlib.util.spec_from_file_location(name, cands[0])
    mod = importlib.util.module_from_spec(spec)
    saved = sys.modules.get(name)
    sys.modules[name] = mod
    try:
        spec.loader.exec_module(mod)
    finally:
        if saved is not None:
            sys.modules[name] = saved
        else:
            sys.modules.pop(name, None)
    return mod


def load():
    if "mods" in _CACHE:
        return _CACHE["mods"]
    if not hasattr(time, "clock"):          # MAS_library.pyx:71,112 call time.clock()
        time.clock = time.perf_counter
    shim = os.path.join(HERE, "ref_shim")
    had_pyfftw = sys.modules.get("pyfftw")
    sys.path.insert(0, shim)
    try:
        masl = _load("MAS_library")
        pkl = _load("Pk_library")
    finally:
        sys.path.remove(shim)
        if had_pyfftw is None:
            # keep the stand-in importable only through the reference module's own reference
            sys.modules.pop("pyfftw", None)
    _CACHE["mods"] = (masl, pkl)
    return masl, pkl


def load_rsl():
    """redshift_space_library (reference), or None when that module was not built."""
    if "rsl" not in _CACHE:
        try:
            _CACHE["rsl"] = _load("redshift_space_library")
        except ImportError:
            _CACHE["rsl"] = None
    return _CACHE["rsl"]


def load_extras():
    """The compiled reference callers / consumers of the hot path as a dict:
    readsnap, readgadget, MAS_gadget, Pk_snapshot, units_library, smoothing_library, bispectrum_library, void_library.
    They import their siblings by bare name (`import MAS_library as MASL`, `import readsnap`, `import h5py` ...), so
    the reference modules are registered under those names only while loading; afterwards sys.modules is restored
    and the product's drop-in modules of the same names are never shadowed."""
    if "extras" in _CACHE:
        return _CACHE["extras"]
    masl, pkl = load()
    rsl = load_rsl()
    shim = os.path.join(HERE, "ref_shim")
    names = ["MAS_library", "Pk_library", "redshift_space_library", "h5py", "pyfftw", "units_library", "readsnap",
             "readgadget", "MAS_gadget", "Pk_snapshot", "smoothing_library", "bispectrum_library", "void_library"]
    saved = {n: sys.modules.get(n) for n in names}
    out = {}
    sys.path.insert(0, shim)
    try:
        sys.modules["MAS_library"], sys.modules["Pk_library"] = masl, pkl
        sys.modules["redshift_space_library"] = rsl
        sys.modules.pop("h5py", None); sys.modules.pop("pyfftw", None)
        for n in ("units_library", "readsnap", "readgadget", "MAS_gadget", "Pk_snapshot", "smoothing_library",
                  "bispectrum_library", "void_library"):
            spec = importlib.util.spec_from_file_location(n, glob.glob(os.path.join(HERE, "_ref", n + "*.so"))[0])
            mod = importlib.util.module_from_spec(spec)
            sys.modules[n] = mod
            spec.loader.exec_module(mod)
            out[n] = mod
    finally:
        sys.path.remove(shim)
        for n, m in saved.items():
            if m is not None:
                sys.modules[n] = m
            else:
                sys.modules.pop(n, None)
    _CACHE["extras"] = out
    return out


def extras_available():
    return all(glob.glob(os.path.join(HERE, "_ref", n + "*.so")) for n in
               ("readsnap", "readgadget", "MAS_gadget", "Pk_snapshot", "smoothing_library", "bispectrum_library",
                "void_library"))
