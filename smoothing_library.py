"""`import smoothing_library as SL` -- drop-in name of the reference package (library/smoothing_library)."""
from pylians_b200.smoothing_library import FT_filter, field_smoothing  # noqa: F401
