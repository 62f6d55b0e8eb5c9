"""`import units_library as UL` -- drop-in name of the reference module (library/units_library.py)."""
from pylians_b200.units_library import units  # noqa: F401
