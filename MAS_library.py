"""`import MAS_library as MASL` -- drop-in name of the reference module (library/MAS_library/__init__.py:1-3)."""
from pylians_b200.MAS_library import *  # noqa: F401,F403
from pylians_b200.MAS_library import MA, FLOAT_type  # noqa: F401
