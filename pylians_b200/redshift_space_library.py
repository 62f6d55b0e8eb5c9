"""Drop-in mirror of library/redshift_space_library.pyx:29-43 (`pos_redshift_space`), on the GPU."""
import numpy as np
import torch

from . import _lib
from .MAS_library import _device, _is_torch, _require_f32


def pos_redshift_space(pos, vel, BoxSize, Hubble, redshift, axis):
    """s = r + (1+z)/H(z) * v along `axis`, wrapped with the reference's rule; `pos` is modified in place."""
    _require_f32(pos, "pos", 2)
    _require_f32(vel, "vel", 2)
    for a, n in ((pos, "pos"), (vel, "vel")):
        if not (a.is_contiguous() if _is_torch(a) else a.flags["C_CONTIGUOUS"]):
            raise ValueError("ndarray is not C-contiguous (%s)" % n)     # float[:,::1] in the reference
    if pos.shape[1] != 3 or vel.shape != pos.shape:
        raise ValueError("pos and vel must both be (N,3)")
    lib = _lib.load()
    dev = _device()
    stream = torch.cuda.current_stream(dev)
    host = not (_is_torch(pos) and pos.is_cuda)
    d_pos = torch.from_numpy(pos).to(dev) if not _is_torch(pos) else (pos if pos.is_cuda else pos.to(dev))
    d_vel = torch.from_numpy(np.ascontiguousarray(vel)).to(dev) if not _is_torch(vel) else (vel if vel.is_cuda else vel.to(dev))
    _lib.check(lib.pylb_pos_redshift_space(d_pos.data_ptr(), d_vel.data_ptr(), pos.shape[0], float(BoxSize),
                                           float(Hubble), float(redshift), int(axis), stream.cuda_stream),
               "pylb_pos_redshift_space")
    if host:
        (torch.from_numpy(pos) if not _is_torch(pos) else pos).copy_(d_pos)
    stream.synchronize()
