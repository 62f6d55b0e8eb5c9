"""Density field of a Gadget snapshot on the GPU: mirror of library/MAS_library/MAS_gadget.py:7-83.

    delta = MASL.density_field_gadget(snapshot_fname, ptypes, dims, MAS='CIC', do_RSD=False, axis=0, verbose=True)

Same arguments and return value (a float32 numpy (dims,dims,dims) array of mass, or of particle counts for a
single species with a header mass).  Built for the GPU: the grid lives in HBM for the whole call and is copied
back once; every (sub-file, species) block is read from disk straight into pinned host memory
(readgadget.SnapFile.read_into), copied to the device asynchronously, converted (kpc/h -> Mpc/h, internal ->
peculiar velocity, redshift-space shift, 1e10 Msun/h -> Msun/h) by the library's elementwise kernels with the
reference's fp32 arithmetic, and deposited with pylb_ma, which only ever adds into the grid -- so streaming
sub-file by sub-file gives the same field as reading the snapshot whole, up to fp32 summation order.
`StreamedSnapshot` is the shared engine; Pk_snapshot.py uses it too.
"""
import math
import time

import numpy as np
import torch

from . import _lib, readgadget
from . import MAS_library as MASL
from .MAS_library import _device


class StreamedSnapshot(object):
    """Streams the species of a binary Gadget snapshot, sub-file by sub-file, to the device.

    for pos, vel, mass, n in s.blocks(ptype, want_vel, want_mass): device tensors of one (sub-file, species) block:
    pos float32 (n,3) in Mpc/h, vel float32 (n,3) in km/s (peculiar) or None, mass float32 (n,) in Msun/h or a
    python float (header mass) or None.  Two pinned staging buffers alternate, so the disk read of block i+1 overlaps
    the H2D copy and the kernels of block i."""

    def __init__(self, snapshot_fname):
        self.name = snapshot_fname
        self.head = readgadget.header(snapshot_fname)
        if self.head.format == "hdf5":
            raise ImportError("HDF5 snapshots need h5py, which is not installed; use the binary formats")
        self.files = readgadget.subfiles(snapshot_fname)
        self.dev = _device()
        self.lib = _lib.load()
        self.stream = torch.cuda.current_stream(self.dev)
        self._pinned = {}
        self._turn = 0

    def count(self, ptype):
        return sum(int(sf.npart[ptype]) for _, sf in self.files)

    def _stage(self, kind, shape):
        """One of two alternating pinned buffers per kind; reused only after its last H2D copy has finished."""
        slot = (kind, self._turn % 2)
        n = int(np.prod(shape))
        buf = self._pinned.get(slot)
        if buf is None or buf[0].numel() < n:
            buf = self._pinned[slot] = (torch.empty(max(n, 1), dtype=torch.float32, pin_memory=True), torch.cuda.Event())
        else:
            buf[1].synchronize()
        return buf[0][:n].view(shape), buf[1]

    def _upload(self, sf, block, ptype, shape, kind):
        host, done = self._stage(kind, shape)
        sf.read_into(block, ptype, host.numpy())
        d = host.to(self.dev, non_blocking=True)
        done.record(self.stream)
        return d

    def blocks(self, ptype, want_vel=False, want_mass=False):
        lib, st = self.lib, self.stream.cuda_stream
        for _, sf in self.files:
            n = int(sf.npart[ptype])
            if n == 0:
                continue
            self._turn += 1
            pos = self._upload(sf, "POS ", ptype, (n, 3), "pos")
            _lib.check(lib.pylb_divide(pos.data_ptr(), 3 * n, 1e3, st), "pylb_divide")            # kpc/h -> Mpc/h
            vel = None
            if want_vel:
                vel = self._upload(sf, "VEL ", ptype, (n, 3), "vel")
                if sf.redshift != 0:                                                            # readsnap.py:375-376
                    _lib.check(lib.pylb_scale_f32(vel.data_ptr(), 3 * n, math.sqrt(sf.time), st), "pylb_scale_f32")
            mass = None
            if want_mass:
                if sf.massarr[ptype] != 0:
                    mass = float(np.float32(sf.massarr[ptype] * 1e10))
                else:
                    mass = self._upload(sf, "MASS", ptype, (n,), "mass")
                    _lib.check(lib.pylb_scale_f32(mass.data_ptr(), n, 1e10, st), "pylb_scale_f32")  # -> Msun/h
            yield pos, vel, mass, n

    def to_redshift_space(self, pos, vel, BoxSize, axis):
        h = self.head
        _lib.check(self.lib.pylb_pos_redshift_space(pos.data_ptr(), vel.data_ptr(), pos.shape[0], float(BoxSize),
                                                    float(h.Hubble), float(h.redshift), int(axis),
                                                    self.stream.cuda_stream), "pylb_pos_redshift_space")

    def sum_f64(self, t):
        """float64 sum of a float32 device tensor (np.sum(mass, dtype=np.float64)) on the device."""
        acc = torch.zeros(1, dtype=torch.float64, device=self.dev)
        _lib.check(self.lib.pylb_grid_sum(t.data_ptr(), t.numel(), acc.data_ptr(), self.stream.cuda_stream),
                   "pylb_grid_sum")
        return acc


def density_field_gadget_device(snapshot_fname, ptypes, dims, MAS="CIC", do_RSD=False, axis=0, verbose=True):
    """density_field_gadget with the result left in HBM: returns (CUDA float32 (dims,dims,dims) tensor, the part of the
    deposited total known on the host (counts, header masses), the part summed on the device (MASS blocks; a 1-element
    float64 tensor, only filled when `verbose`))."""
    snap = StreamedSnapshot(snapshot_fname)
    head = snap.head
    BoxSize = head.boxsize / 1e3                              # Mpc/h
    if list(ptypes) == [-1]:
        ptypes = [0, 1, 2, 3, 4, 5]
    single_component = len(ptypes) == 1
    density = torch.zeros((dims, dims, dims), dtype=torch.float32, device=snap.dev)
    num, num_dev = 0.0, torch.zeros(1, dtype=torch.float64, device=snap.dev)
    # the reference loops files outermost and species inside (MAS_gadget.py:37-79); addition order is the only
    # thing that differs here (species outermost), and MA is order-independent up to fp32 rounding
    for ptype in ptypes:
        if snap.count(ptype) == 0:
            continue
        header_mass = head.massarr[ptype] != 0
        for pos, vel, mass, n in snap.blocks(ptype, want_vel=do_RSD, want_mass=not (header_mass and single_component)):
            if do_RSD:
                snap.to_redshift_space(pos, vel, BoxSize, axis)
            if header_mass and single_component:              # :60-63: plain counts
                MASL.MA(pos, density, BoxSize, MAS)
                num += n
            else:
                if not torch.is_tensor(mass):                 # :64-67: np.ones(n, float32)*Masses[ptype]
                    num += n * float(np.float32(mass))
                    mass = torch.full((n,), mass, dtype=torch.float32, device=snap.dev)
                elif verbose:
                    num_dev += snap.sum_f64(mass)
                MASL.MA(pos, density, BoxSize, MAS, W=mass)
    return density, num, num_dev


def density_field_gadget(snapshot_fname, ptypes, dims, MAS="CIC", do_RSD=False, axis=0, verbose=True):
    """MAS_gadget.py:7-83."""
    start = time.time()
    if verbose:
        print("\nComputing density field of particles %s" % (ptypes,))
    density, num, num_dev = density_field_gadget_device(snapshot_fname, ptypes, dims, MAS, do_RSD, axis, verbose)
    out = density.cpu().numpy()
    if verbose:
        print("%.8e should be equal to\n%.8e" % (np.sum(out, dtype=np.float64), num + float(num_dev.item())))
        print("Time taken = %.2f seconds" % (time.time() - start))
    return out
