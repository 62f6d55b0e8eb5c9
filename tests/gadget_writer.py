"""Write small synthetic Gadget snapshots (format 1 or 2, either byte order, any number of sub-files) for the
reader / driver tests.  TEST INFRASTRUCTURE: the layout follows the public Gadget-2 snapshot specification
(256-byte header record, then POS, VEL, ID and -- if some species has no header mass -- MASS records, each framed
by 4-byte Fortran record markers; format 2 adds a 16-byte label record in front of each)."""
import numpy as np


MARKER_MOD = 1 << 32      # record markers hold the length modulo this (tests of >= 4 GiB records shrink it)


def _record(f, payload, order, label=None, fmt=1):
    u4 = np.dtype(order + "u4")
    if fmt == 2:
        f.write(np.array([8], u4).tobytes())
        f.write(label.encode("ascii"))
        f.write(np.array([len(payload) + 8], u4).tobytes())
        f.write(np.array([8], u4).tobytes())
    f.write(np.array([len(payload) % MARKER_MOD], u4).tobytes())
    f.write(payload)
    f.write(np.array([len(payload) % MARKER_MOD], u4).tobytes())


def make_particles(seed, counts, box_kpc, masstable, clustered=False):
    """counts[6] particles per species -> dict species -> (pos [kpc/h] float32 (n,3), vel float32 (n,3), ids uint32,
    mass float32 (1e10 Msun/h))."""
    rng = np.random.default_rng(seed)
    out, next_id = {}, 1
    for t, n in enumerate(counts):
        if n == 0:
            continue
        pos = rng.random((n, 3)) * box_kpc
        if clustered:
            pos = np.mod(pos + 0.05 * box_kpc * np.sin(2 * np.pi * pos[:, ::-1] / box_kpc), box_kpc)
        pos = pos.astype(np.float32)
        pos[pos >= np.float32(box_kpc)] = 0.0
        vel = (rng.standard_normal((n, 3)) * 300.0).astype(np.float32)
        ids = np.arange(next_id, next_id + n, dtype=np.uint32)
        next_id += n
        mass = (rng.random(n) * 0.02 + 0.01).astype(np.float32) if masstable[t] == 0 else None
        out[t] = (pos, vel, ids, mass)
    return out


def write_snapshot(base, parts, masstable, box_kpc, redshift, nfiles=1, fmt=1, order="<", omega_m=0.3175,
                   omega_l=0.6825, hubble=0.6711, extra=()):
    """Split every species evenly over `nfiles` files named base (nfiles == 1) or base.0 ... base.(nfiles-1).
    extra: further blocks after MASS, [(label, {species: float32 array (n,) or (n,3)}), ...] in file order (U, RHO, ...)."""
    time = 1.0 / (1.0 + redshift)
    nall = np.zeros(6, np.uint32)
    for t, p in parts.items():
        nall[t] = len(p[0])
    names = []
    for i in range(nfiles):
        name = base if nfiles == 1 else "%s.%d" % (base, i)
        names.append(name)
        chunk = {t: tuple(None if a is None else np.array_split(a, nfiles)[i] for a in p) for t, p in parts.items()}
        npart = np.zeros(6, np.int32)
        for t, p in chunk.items():
            npart[t] = len(p[0])
        head = np.zeros(1, np.dtype([("npart", order + "i4", 6), ("massarr", order + "f8", 6), ("time", order + "f8"),
                                     ("redshift", order + "f8"), ("sfr", order + "i4"), ("feedback", order + "i4"),
                                     ("nall", order + "u4", 6), ("cooling", order + "i4"), ("filenum", order + "i4"),
                                     ("boxsize", order + "f8"), ("omega_m", order + "f8"), ("omega_l", order + "f8"),
                                     ("hubble", order + "f8"), ("pad", "u1", 96)]))
        assert head.dtype.itemsize == 256
        head["npart"], head["massarr"], head["time"], head["redshift"] = npart, masstable, time, redshift
        head["nall"], head["filenum"], head["boxsize"] = nall, nfiles, box_kpc
        head["omega_m"], head["omega_l"], head["hubble"] = omega_m, omega_l, hubble
        types = sorted(chunk)
        with open(name, "wb") as f:
            _record(f, head.tobytes(), order, "HEAD", fmt)
            cat = lambda k, dt: np.concatenate([chunk[t][k] for t in types]).astype(order + dt).tobytes()   # noqa: E731
            _record(f, cat(0, "f4"), order, "POS ", fmt)
            _record(f, cat(1, "f4"), order, "VEL ", fmt)
            _record(f, cat(2, "u4"), order, "ID  ", fmt)
            with_mass = [t for t in types if masstable[t] == 0 and npart[t] > 0]
            if with_mass:
                _record(f, np.concatenate([chunk[t][3] for t in with_mass]).astype(order + "f4").tobytes(), order,
                        "MASS", fmt)
            for label, per_type in extra:
                pieces = [np.array_split(per_type[t], nfiles)[i] for t in sorted(per_type)]
                _record(f, np.concatenate(pieces).astype(order + "f4").tobytes(), order, label, fmt)
    return names
