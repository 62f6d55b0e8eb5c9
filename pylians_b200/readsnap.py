"""Mirror of library/readsnap.py (`snapshot_header` :26-93, `find_block` :97-156, `read_block` :161-379,
`list_format2_blocks` :382-398, `read_gadget_header` :401-430) on top of readgadget.SnapFile's one-pass record index:
the same names, arguments, block table and return values, for format-1 and format-2 files of either byte order.

    head = readsnap.snapshot_header("snap_063.0")
    mass = readsnap.read_block("snap_063", "MASS", parttype=5)

Host-side I/O only.  Errors raise (IOError / ValueError) where the reference prints a message and calls sys.exit()."""
import math
import os

import numpy as np

from .readgadget import _snapfile

_VEC = np.dtype((np.float32, 3))


def _resolve(filename):
    if os.path.exists(filename):
        return filename, True
    if os.path.exists(filename + ".0"):
        return filename + ".0", False
    raise IOError("file not found: %s" % filename)


class snapshot_header(object):
    """readsnap.py:26-93.  `filename` may omit the ".0" of a multi-file snapshot."""

    def __init__(self, filename):
        cur, _ = _resolve(filename)
        sf = _snapfile(cur)
        self.filename = filename
        self.format, self.swap = sf.format, sf.swap
        for name in ("npart", "massarr", "time", "redshift", "sfr", "feedback", "nall", "cooling", "filenum", "boxsize",
                     "omega_m", "omega_l", "hubble"):
            setattr(self, name, getattr(sf, name))


def block_table(block, massarr, arepo=0, no_masses=False):
    """(species holding data in the block, item dtype, record number in a format-1 file), readsnap.py:208-303."""
    blockadd = {0: 0, 1: 1, 2: 4}[arepo]
    blocksub = 1 if no_masses else 0
    has = np.zeros(6, bool)
    dt = np.dtype(np.float32)
    if block in ("POS ", "VEL ", "ID  ", "ACCE"):
        has[:] = True
        num = {"POS ": 2, "VEL ": 3, "ID  ": 4, "ACCE": 5}[block]
        dt = np.dtype(np.uint32) if block == "ID  " else _VEC
    elif block == "MASS":
        has[np.asarray(massarr) == 0] = True
        num = 5
    elif block in ("U   ", "RHO ", "VOL ", "CMCE", "AREA", "NFAC"):
        has[0] = True
        num = {"U   ": 6, "RHO ": 7, "VOL ": 8, "CMCE": 9, "AREA": 10, "NFAC": 11}[block] - blocksub
        if block == "CMCE":
            dt = _VEC
        if block == "NFAC":
            dt = np.dtype(np.int64)
    elif block in ("NE  ", "NH  ", "HSML", "SFR ", "MHI ", "TEMP"):
        has[0] = True
        num = {"NE  ": 8, "NH  ": 9, "HSML": 10, "SFR ": 11, "MHI ": 12, "TEMP": 13}[block] + blockadd - blocksub
    elif block == "AGE ":
        has[4] = True
        num = 12 + blockadd - blocksub
    elif block == "Z   ":
        has[0] = has[4] = True
        num = 13 + blockadd - blocksub
    elif block in ("BHMA", "BHMD"):
        has[5] = True
        num = {"BHMA": 14, "BHMD": 15}[block] + blockadd - blocksub
    else:
        raise ValueError("Sorry! Block type %s not known!" % block)
    return has, dt, num


def find_block(filename, format, swap, block, block_num, only_list_blocks=False):
    """readsnap.py:97-156: (byte offset of the block's payload, its size in bytes); or print the record table."""
    if not os.path.exists(filename):
        raise IOError("file not found: %s" % filename)
    sf = _snapfile(filename)
    if only_list_blocks:
        for k, (label, off, n) in enumerate(sf.records, 1):
            print(("%d %s %d %d" % (k, label, off, n)) if sf.format == 2 else ("%d %d %d" % (k, off, n)))
        return None
    return sf._record(block, block_num)


def read_block(filename, block, parttype=-1, physical_velocities=True, arepo=0, no_masses=False, verbose=False,
               nall=(0, 0, 0, 0, 0, 0)):
    """readsnap.py:161-379: one block of a snapshot (all sub-files unless `filename` names one), for one species or, with
    parttype = -1, for every species that has data in the block, ordered by species."""
    if verbose:
        print("reading block %s" % block)
    if parttype not in (-1, 0, 1, 2, 3, 4, 5):
        raise ValueError("wrong parttype given")
    cur, single_file = _resolve(filename)
    first = _snapfile(cur)
    nall = np.asarray(nall)
    if np.all(nall == 0):
        nall = first.nall
    massarr, filenum = first.massarr, int(first.filenum)
    has, dt, num = block_table(block, massarr, arepo, no_masses)
    if block == "MASS" and parttype >= 0 and massarr[parttype] > 0:          # :237-243
        n = int(first.npart[parttype]) if single_file else int(nall[parttype])
        return np.full(n, np.float32(massarr[parttype]), dtype=np.float32)
    actual = has.copy()
    if parttype >= 0:
        if not has[parttype]:
            raise ValueError("Error: no data for specified particle type %d in the block %s" % (parttype, block))
        actual[:] = False
        actual[parttype] = True
    elif block == "MASS":
        actual[:] = True
    files = [first] if single_file else [first] + [_snapfile("%s.%d" % (filename, i)) for i in range(1, filenum)]
    # totals from the per-file counts when the whole snapshot is read (exact beyond 2^32), else this file's
    total = {j: sum(int(sf.npart[j]) for sf in files) for j in range(6)}
    species_offset, allpartnum = np.zeros(6, np.int64), 0
    for j in range(6):
        species_offset[j] = allpartnum
        if actual[j]:
            allpartnum += total[j]
    data = None
    for i, sf in enumerate(files):
        npart = sf.npart
        cur_offset, curpartnum = np.zeros(6, np.int64), 0
        for j in range(6):
            cur_offset[j] = curpartnum
            if has[j]:
                curpartnum += int(npart[j])
        off, blocksize = sf._record(block, num)
        if i == 0:
            if block == "ID  " and curpartnum and blocksize == 8 * curpartnum:   # long IDs, :349-352
                dt = np.dtype(np.uint64)
            data = np.empty(allpartnum, dt)
        if dt.itemsize * curpartnum != blocksize:
            raise IOError("something wrong with blocksize! expected = %d actual = %d" % (dt.itemsize * curpartnum, blocksize))
        count, skip = (int(npart[parttype]), int(cur_offset[parttype])) if parttype >= 0 else (curpartnum, 0)
        width = 3 if dt == _VEC else 1                            # floats per item
        with open(sf.path, "rb") as f:
            f.seek(off + skip * dt.itemsize)
            curdat = np.fromfile(f, dtype=dt.base.newbyteorder(sf.order), count=count * width)
        curdat = curdat.astype(dt.base, copy=False)               # native byte order
        if width == 3:
            curdat = curdat.reshape(-1, 3)
        for j in range(6):
            if not actual[j]:
                continue
            n = int(npart[j])
            lo = int(species_offset[j])
            if block == "MASS" and massarr[j] > 0:
                data[lo:lo + n] = massarr[j]
            elif parttype >= 0:
                data[lo:lo + n] = curdat
            else:
                data[lo:lo + n] = curdat[cur_offset[j]:cur_offset[j] + n]
            species_offset[j] += n
    if physical_velocities and block == "VEL " and first.redshift != 0:
        data *= math.sqrt(first.time)
    return data


def list_format2_blocks(filename):
    """readsnap.py:382-398."""
    cur, _ = _resolve(filename)
    sf = _snapfile(cur)
    print("GADGET FORMAT  %d" % sf.format)
    print("#   OFFSET   SIZE" if sf.format != 2 else "#   BLOCK   OFFSET   SIZE")
    print("-------------------------")
    find_block(cur, sf.format, sf.swap, "XXXX", 0, only_list_blocks=True)
    print("-------------------------")


def read_gadget_header(filename):
    """readsnap.py:401-430: print the header and the Omegas it implies."""
    head = snapshot_header(filename)
    print("npar= %s" % head.npart)
    print("nall= %s" % head.nall)
    print("a= %s" % head.time)
    print("z= %s" % head.redshift)
    print("masses= %s Msun/h" % (head.massarr * 1e10))
    print("boxsize= %s kpc/h" % head.boxsize)
    print("filenum= %s" % head.filenum)
    print("cooling= %s" % head.cooling)
    print("Omega_m,Omega_l= %s %s" % (head.omega_m, head.omega_l))
    print("h= %s \\n" % head.hubble)
    rhocrit = 2.77536627e11 / 1e9                               # h^2 Msun/kpc^3
    Omega_CDM = head.nall[1] * head.massarr[1] * 1e10 / (head.boxsize ** 3 * rhocrit)
    print("DM mass=%.5e  Omega_DM = %.5f" % (head.massarr[1] * 1e10, Omega_CDM))
    if head.nall[2] > 0 and head.massarr[2] > 0:
        Omega_NU = head.nall[2] * head.massarr[2] * 1e10 / (head.boxsize ** 3 * rhocrit)
        print("NU mass=%.5e  Omega_NU = %.5f" % (head.massarr[2] * 1e10, Omega_NU))
        print("Sum of neutrino masses=%.5f eV" % (Omega_NU * head.hubble ** 2 * 94.1745))
