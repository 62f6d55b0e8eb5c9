"""pylians_b200 -- B200-native (sm_100a) implementation of Pylians' density-field -> power-spectrum
hot path: MAS_library.MA (NGP/CIC/TSC/PCS, weights) and Pk_library.Pk / XPk, behind the reference's
own call signatures, plus the callers (Gadget readers, snapshot drivers) and FFT consumers (smoothing, bispectrum)
either side of it.  CUDA kernels + C ABI live in csrc/ (built by `python -m pylians_b200.build`);
the modules here are the thin host-side mirror of the reference's Python interface.
"""
from . import _lib  # noqa: F401

__all__ = ["MAS_library", "Pk_library", "redshift_space_library", "dist", "set_verbose",
           # callers / data formats / FFT consumers either side of the path
           "readgadget", "readsnap", "MAS_gadget", "Pk_snapshot", "units_library", "smoothing_library", "void_library",
           "bispectrum_library"]


def set_verbose(flag):
    """The reference prints progress lines from Pk/XPk unconditionally; silence them with False."""
    from . import Pk_library
    Pk_library.VERBOSE = bool(flag)
