"""`import readgadget` -- drop-in name of the reference module (library/readgadget.py)."""
from pylians_b200.readgadget import fname_format, header, read_field, read_block, SnapFile, subfiles  # noqa: F401
