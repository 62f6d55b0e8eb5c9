"""`import void_library as VL` -- drop-in name of the reference package (library/void_library); only the FFT
consumer `gaussian_smoothing` is provided."""
from pylians_b200.void_library import gaussian_smoothing  # noqa: F401
