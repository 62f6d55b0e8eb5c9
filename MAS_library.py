"""`import MAS_library as MASL` -- drop-in name of the reference package (library/MAS_library/__init__.py:1-3:
`from MAS_library import *; from MAS_gadget import *`)."""
from pylians_b200.MAS_library import *  # noqa: F401,F403
from pylians_b200.MAS_library import MA, FLOAT_type  # noqa: F401
from pylians_b200.MAS_gadget import density_field_gadget, density_field_gadget_device  # noqa: F401
