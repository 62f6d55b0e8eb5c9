"""`import redshift_space_library as RSL` -- drop-in name of the reference module."""
from pylians_b200.redshift_space_library import pos_redshift_space  # noqa: F401
