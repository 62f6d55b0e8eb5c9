"""`import readsnap` -- drop-in name of the reference module (library/readsnap.py)."""
from pylians_b200.readsnap import (snapshot_header, find_block, read_block, list_format2_blocks,  # noqa: F401
                                   read_gadget_header, block_table)
