"""Placeholder for `h5py`, which the reference's readgadget.py imports at module level (readgadget.py:4) and this
image does not have.  TEST INFRASTRUCTURE.  Only the binary (format 1/2) snapshot paths are exercised; any HDF5
access raises."""


class File(object):
    def __init__(self, *a, **k):
        raise ImportError("h5py is not installed in this image; HDF5 snapshots cannot be read")
