"""Gadget snapshot reader: the data format on the input side of the MA -> Pk path.

Mirrors the interface of library/readgadget.py (`fname_format` :7-19, `header` :23-63, `read_field` :67-103,
`read_block` :108-154) and the parts of library/readsnap.py it delegates to for format-1 / format-2 binary files
(`snapshot_header` :26-93, `find_block` :97-156, `read_block` :161-379): same names, arguments, units and
return types, so `readgadget.read_block(snap, "POS ", [1])/1e3` keeps working.

Built differently from the reference: every file is indexed ONCE (`SnapFile`: record table from the Fortran record
markers, either byte order, format-2 labels), blocks are read with `readinto` straight into a caller-supplied
buffer (the snapshot drivers pass pinned host tensors, so a block goes disk -> pinned memory -> HBM with no
intermediate copy), and particle totals are taken from the per-file counts, so snapshots with more than 2^32
particles of a type (2048^3, the north-star size; header `nall` is 32-bit) need no manual override; records of
4 GiB or more, whose Fortran markers wrap modulo 2^32, are indexed by trying the wrapped lengths.

This is host-side I/O: no arithmetic beyond the reference's unit conversions.  HDF5 snapshots need `h5py`, which
this image does not have; those paths raise ImportError instead of guessing.
"""
import math
import os

import numpy as np

try:                                     # library/readgadget.py:4 imports it unconditionally
    import h5py
except ImportError:                      # not in this image; binary snapshots still work
    h5py = None

_HEADER_FIELDS = [("npart", "i4", (6,)), ("massarr", "f8", (6,)), ("time", "f8"), ("redshift", "f8"),
                  ("sfr", "i4"), ("feedback", "i4"), ("nall", "u4", (6,)), ("cooling", "i4"), ("filenum", "i4"),
                  ("boxsize", "f8"), ("omega_m", "f8"), ("omega_l", "f8"), ("hubble", "f8"),
                  ("stellarage", "i4"), ("metals", "i4"), ("nall_hw", "u4", (6,))]
_BLOCK_ORDINAL = {"POS ": 2, "VEL ": 3, "ID  ": 4, "MASS": 5}      # format-1 record order, readsnap.py:214-243
_VECTOR = np.dtype((np.float32, 3))


def _header_dtype(order):
    return np.dtype([(f[0], order + f[1]) + tuple(f[2:]) for f in _HEADER_FIELDS])


def fname_format(snapshot):
    """readgadget.py:7-19: resolve `snapshot` to an existing file and its container format."""
    if os.path.exists(snapshot):
        return (snapshot, "hdf5") if snapshot[-4:] == "hdf5" else (snapshot, "binary")
    if os.path.exists(snapshot + ".0"):
        return snapshot + ".0", "binary"
    if os.path.exists(snapshot + ".hdf5"):
        return snapshot + ".hdf5", "hdf5"
    if os.path.exists(snapshot + ".0.hdf5"):
        return snapshot + ".0.hdf5", "hdf5"
    raise Exception("File not found!")


class SnapFile(object):
    """One binary Gadget file, indexed once: byte order, format (1|2), header, and the table of data records."""
    _MARKER_MOD = 1 << 32            # Fortran record markers are 32-bit: lengths wrap modulo this (a test shrinks it)

    def __init__(self, path):
        self.path = path
        size = os.path.getsize(path)
        with open(path, "rb") as f:
            first = f.read(4)
            if len(first) < 4:
                raise IOError("incorrect file format encountered when reading header of %s" % path)
            for order in ("<", ">"):
                v = int(np.frombuffer(first, order + "u4")[0])
                if v in (8, 256):
                    self.order, self.format = order, (2 if v == 8 else 1)
                    break
            else:                                            # readsnap.py:51-52
                raise IOError("incorrect file format encountered when reading header of %s" % path)
            self.swap = int(self.order == ">")
            u4 = np.dtype(self.order + "u4")
            # walk the Fortran records: [u4 n][n bytes][u4 n]; format 2 puts a 16-byte label record before each
            self.records = []                                # (label or None, payload offset, payload bytes)
            pos, label = 0, None
            while pos + 8 <= size:
                f.seek(pos)
                n = int(np.frombuffer(f.read(4), u4)[0])
                if self.format == 2 and label is None and n == 8:
                    label = f.read(4).decode("ascii", "replace")
                    pos += 16
                    continue
                # a record of 4 GiB or more (the POS block of one 2048^3 file is 103 GB) carries its length modulo 2^32
                # in both markers: the true length is the first n + k 2^32 whose trailing marker repeats the leading one
                low, ok = n, False
                while pos + 8 + n <= size:
                    f.seek(pos + 4 + n)
                    tail = f.read(4)
                    if len(tail) == 4 and int(np.frombuffer(tail, u4)[0]) == low:
                        ok = True
                        break
                    n += self._MARKER_MOD
                if not ok:                                                     # readsnap.py:146-148
                    raise IOError("something wrong: record markers of %s disagree at byte %d" % (path, pos))
                self.records.append((label, pos + 4, n))
                label = None
                pos += n + 8
            if not self.records or self.records[0][2] != 256:
                raise IOError("incorrect file format encountered when reading header of %s" % path)
            f.seek(self.records[0][1])
            hd = _header_dtype(self.order)
            h = np.frombuffer(f.read(hd.itemsize), hd)[0]
        native = lambda a: np.ascontiguousarray(a).astype(a.dtype.newbyteorder("="))    # noqa: E731
        self.npart = native(h["npart"])
        self.massarr = native(h["massarr"])
        self.nall = native(h["nall"])
        self.nall_hw = native(h["nall_hw"])
        for name in ("time", "redshift", "boxsize", "omega_m", "omega_l", "hubble"):
            setattr(self, name, np.float64(h[name]))
        for name in ("sfr", "feedback", "cooling", "filenum"):
            setattr(self, name, np.int32(h[name]))

    # ---- where a block lives -------------------------------------------------------------------
    def _record(self, block, ordinal=None):
        """(payload offset, payload bytes) of a block: by label in format 2, by record number (1 = header) in format 1."""
        if self.format == 2:
            for label, off, n in self.records:
                if label == block:
                    return off, n
        else:
            k = (_BLOCK_ORDINAL[block] if ordinal is None else ordinal) - 1
            if 0 <= k < len(self.records):
                return self.records[k][1], self.records[k][2]
        raise IOError("Error: block not found (%r in %s)" % (block, self.path))        # readsnap.py:153-155

    def _types_in_block(self, block):
        # POS/VEL/ID hold every species; MASS only those without a header mass (readsnap.py:217-236)
        return self.massarr == 0 if block == "MASS" else np.ones(6, bool)

    def item_dtype(self, block):
        if block in ("POS ", "VEL "):
            return np.dtype((self.order + "f4", 3))
        if block == "MASS":
            return np.dtype(self.order + "f4")
        if block == "ID  ":                                  # 64-bit IDs when the record is twice as long (:349-352)
            _, n = self._record(block)
            count = int(self.npart.sum())
            return np.dtype(self.order + ("u8" if count and n == 8 * count else "u4"))
        raise Exception("block not implemented in readgadget!")

    def span(self, block, ptype):
        """(byte offset, particle count, on-disk item dtype) of species `ptype` inside `block`."""
        present = self._types_in_block(block)
        if not present[ptype]:
            raise IOError("Error: no data for specified particle type %d in the block %s" % (ptype, block))
        dt = self.item_dtype(block)
        off, nbytes = self._record(block)
        count = int(self.npart[present].sum())
        if dt.itemsize * count != nbytes:                    # readsnap.py:354-356
            raise IOError("something wrong with blocksize! expected = %d actual = %d" % (dt.itemsize * count, nbytes))
        before = int(self.npart[:ptype][present[:ptype]].sum())
        return off + before * dt.itemsize, int(self.npart[ptype]), dt

    def read_into(self, block, ptype, out):
        """Read species `ptype` of `block` into `out` (a writable C-contiguous numpy array of the right size:
        float32 (n,3) for POS/VEL, float32 (n,) for MASS, uint32/uint64 (n,) for ID), native byte order."""
        off, count, dt = self.span(block, ptype)
        flat = out.reshape(-1).view(np.uint8)
        if flat.size != count * dt.itemsize:
            raise ValueError("read_into: buffer holds %d bytes, block needs %d" % (flat.size, count * dt.itemsize))
        with open(self.path, "rb", buffering=0) as f:
            f.seek(off)
            got = f.readinto(memoryview(flat))
            while got < flat.size:                           # short reads on large files
                more = f.readinto(memoryview(flat[got:]))
                if not more:
                    raise IOError("unexpected end of file in %s" % self.path)
                got += more
        if self.swap:
            out.byteswap(True)
        return out

    def out_dtype(self, block):
        """Native-order dtype of the arrays handed back for `block`."""
        dt = self.item_dtype(block)
        return _VECTOR if dt.shape == (3,) else dt.newbyteorder("=")

    def read(self, block, ptype):
        _, count, _ = self.span(block, ptype)
        out = np.empty(count, dtype=self.out_dtype(block))
        return self.read_into(block, ptype, out)


_INDEX = {}


def _snapfile(path):
    """Index cache keyed by (path, mtime, size): a driver touches each sub-file several times (header, POS, VEL, MASS)."""
    st = os.stat(path)
    key = (os.path.abspath(path), st.st_mtime_ns, st.st_size)
    sf = _INDEX.get(key)
    if sf is None:
        if len(_INDEX) > 4096:
            _INDEX.clear()
        sf = _INDEX[key] = SnapFile(path)
    return sf


def _need_h5py():
    if h5py is None:
        raise ImportError("h5py is not installed: HDF5 Gadget snapshots cannot be read in this environment")
    return h5py


class header(object):
    """readgadget.py:23-63.  Attributes: time, redshift, boxsize, filenum, omega_m, omega_l, hubble, massarr, npart,
    nall, cooling, format, Hubble [km/s/(Mpc/h)].  Extra: nall64 (nall with the header's high words)."""

    def __init__(self, snapshot):
        filename, fformat = fname_format(snapshot)
        if fformat == "hdf5":
            with _need_h5py().File(filename, "r") as f:
                a = f["Header"].attrs
                self.time, self.redshift, self.boxsize = a[u"Time"], a[u"Redshift"], a[u"BoxSize"]
                self.filenum, self.omega_m, self.omega_l = a[u"NumFilesPerSnapshot"], a[u"Omega0"], a[u"OmegaLambda"]
                self.hubble, self.massarr = a[u"HubbleParam"], a[u"MassTable"]
                self.npart, self.nall, self.cooling = a[u"NumPart_ThisFile"], a[u"NumPart_Total"], a[u"Flag_Cooling"]
                self.nall64 = np.asarray(self.nall, np.int64)
                if u"NumPart_Total_HighWord" in a:
                    self.nall64 = self.nall64 + (np.asarray(a[u"NumPart_Total_HighWord"], np.int64) << 32)
            self.format = "hdf5"
        else:
            sf = _snapfile(filename)
            for name in ("time", "redshift", "boxsize", "filenum", "omega_m", "omega_l", "hubble", "massarr", "npart",
                         "nall", "cooling", "format"):
                setattr(self, name, getattr(sf, name))
            self.nall64 = sf.nall.astype(np.int64) + (sf.nall_hw.astype(np.int64) << 32)
        self.Hubble = 100.0 * np.sqrt(self.omega_m * (1.0 + self.redshift) ** 3 + self.omega_l)


def _physical_velocity(array, time, redshift):
    # readsnap.py:375-376: u = v_internal*sqrt(a), skipped at z = 0 (a = 1); in place, so the product stays float32
    if redshift != 0:
        array *= math.sqrt(time)
    return array


def _header_mass(count, mass):
    # readsnap.py:237-243 / readgadget.py:86-88: np.ones(n, float32)*massarr.  float32 under the numpy the reference
    # was written for (value-based casting); numpy >= 2 would promote to float64, which MASL.MA then rejects.
    return np.full(count, np.float32(mass), dtype=np.float32)


def read_field(snapshot, block, ptype):
    """readgadget.py:67-103: one block of ONE file (the file `snapshot` resolves to) for one particle type."""
    filename, fformat = fname_format(snapshot)
    if fformat == "binary":
        sf = _snapfile(filename)
        if block == "MASS" and ptype >= 0 and sf.massarr[ptype] > 0:
            return _header_mass(int(sf.npart[ptype]), sf.massarr[ptype])
        array = sf.read(block, ptype)
        return _physical_velocity(array, sf.time, sf.redshift) if block == "VEL " else array
    head = header(filename)
    names = {"POS ": "Coordinates", "MASS": "Masses", "ID  ": "ParticleIDs", "VEL ": "Velocities"}
    if block not in names:
        raise Exception("block not implemented in readgadget!")
    key = "PartType%d/%s" % (ptype, names[block])
    with _need_h5py().File(filename, "r") as f:
        if key not in f:
            if head.massarr[ptype] * 1e10 != 0.0:            # readgadget.py:85-89 (any missing block, sic)
                return _header_mass(int(head.npart[ptype]), head.massarr[ptype] * 1e10)
            raise Exception("Problem reading the block %s" % block)
        array = f[key][:]
    if block == "VEL ":
        array *= np.sqrt(head.time)
    if block == "POS " and array.dtype == np.float64:
        array = array.astype(np.float32)
    return array


def subfiles(snapshot):
    """[(path, SnapFile)] of a binary snapshot: the file itself when `snapshot` names one, else snapshot.0 ...
    snapshot.(filenum-1) (readsnap.py:186-193, 317-320)."""
    if os.path.exists(snapshot):
        return [(snapshot, _snapfile(snapshot))]
    first = _snapfile(snapshot + ".0") if os.path.exists(snapshot + ".0") else None
    if first is None:
        raise Exception("File not found!")
    out = [(snapshot + ".0", first)]
    for i in range(1, int(first.filenum)):
        out.append(("%s.%d" % (snapshot, i), _snapfile("%s.%d" % (snapshot, i))))
    return out


def _read_block_binary(snapshot, block, pt, out=None):
    """All files of a binary snapshot, one species: what readsnap.read_block(snapshot, block, parttype=pt) returns."""
    files = subfiles(snapshot)
    first = files[0][1]
    total = sum(int(sf.npart[pt]) for _, sf in files)
    if block == "MASS" and first.massarr[pt] > 0:
        return _header_mass(total, first.massarr[pt])
    if out is None:
        out = np.empty(total, dtype=first.out_dtype(block))
    lo = 0
    for _, sf in files:
        n = int(sf.npart[pt])
        if n:
            sf.read_into(block, pt, out[lo:lo + n])
        lo += n
    if block == "VEL ":
        _physical_velocity(out, first.time, first.redshift)
    return out


def read_block(snapshot, block, ptype, verbose=False):
    """readgadget.py:108-154: a block of the WHOLE snapshot (all files) for the particle types in the list `ptype`,
    concatenated in list order.  `ptype=[-1]` means every species (the reference indexes `Nall[-1]` there and fails;
    its callers -- Pk_snapshot.py:72,76 -- mean "all")."""
    filename, fformat = fname_format(snapshot)
    if block not in ("POS ", "VEL ", "MASS", "ID  "):
        raise Exception("block not implemented in readgadget!")
    types = [0, 1, 2, 3, 4, 5] if list(ptype) == [-1] else list(ptype)
    if fformat == "binary":
        files = subfiles(snapshot)
        counts = [sum(int(sf.npart[pt]) for _, sf in files) for pt in types]
        dtype = files[0][1].out_dtype(block)
        array = np.zeros(sum(counts), dtype=dtype)
        offset = 0
        for pt, n in zip(types, counts):
            if verbose:
                print("reading block %s" % block)
            if n:
                if block == "MASS" and files[0][1].massarr[pt] > 0:
                    array[offset:offset + n] = np.float32(files[0][1].massarr[pt])
                else:
                    _read_block_binary(snapshot, block, pt, out=array[offset:offset + n])
            offset += n
        return array
    head = header(filename)
    nall = np.asarray(head.nall64)
    dtype = _VECTOR if block in ("POS ", "VEL ") else (np.float32 if block == "MASS" else
                                                        read_field(filename, block, types[0]).dtype)
    array = np.zeros(int(sum(nall[pt] for pt in types)), dtype=dtype)
    offset = 0
    for pt in types:
        if head.filenum == 1:
            array[offset:offset + nall[pt]] = read_field(snapshot, block, pt)
            offset += int(nall[pt])
        else:
            for i in range(int(head.filenum)):
                fn = "%s.%d.hdf5" % (snapshot, i)
                n = int(header(fn).npart[pt])
                array[offset:offset + n] = read_field(fn, block, pt)
                offset += n
    if offset != array.shape[0]:
        raise Exception("not all particles read!!!!")
    return array
