"""Gadget reader (pylians_b200/readgadget.py) against the compiled, unmodified reference reader (oracle/_ref:
readgadget.py + readsnap.py) on synthetic format-1 snapshots, and format 2 / big-endian / multi-file files against
the format-1 result.  CPU only: this is host-side I/O."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gadget_writer as GW                              # noqa: E402
from oracle import ref_loader                           # noqa: E402
from pylians_b200 import readgadget as RG               # noqa: E402

BOX, Z = 25000.0, 0.5
MASSTABLE = np.array([0.0, 0.65, 0.0, 0.0, 0.0, 0.0])   # gas and stars carry individual masses
COUNTS = [700, 1500, 0, 0, 300, 0]


@pytest.fixture(scope="module")
def parts():
    return GW.make_particles(3, COUNTS, BOX, MASSTABLE)


def _expect(parts, block, pt, time):
    pos, vel, ids, mass = parts[pt]
    if block == "POS ":
        return pos
    if block == "VEL ":
        return vel * np.float32(np.sqrt(time))
    if block == "ID  ":
        return ids
    return mass if mass is not None else np.full(len(pos), np.float32(MASSTABLE[pt]), np.float32)


@pytest.mark.parametrize("fmt,order,nfiles", [(1, "<", 1), (1, "<", 3), (2, "<", 1), (2, "<", 3), (1, ">", 2), (2, ">", 2)])
def test_reader_returns_what_was_written(tmp_path, parts, fmt, order, nfiles):
    base = str(tmp_path / "snap_003")
    names = GW.write_snapshot(base, parts, MASSTABLE, BOX, Z, nfiles, fmt, order)
    h = RG.header(base)
    assert h.format == fmt and h.filenum == nfiles
    assert h.boxsize == BOX and h.redshift == Z and abs(h.time - 1 / (1 + Z)) < 1e-15
    assert list(h.nall) == COUNTS and list(h.nall64) == COUNTS
    np.testing.assert_array_equal(h.massarr, MASSTABLE)
    assert abs(h.Hubble - 100.0 * np.sqrt(h.omega_m * (1 + Z) ** 3 + h.omega_l)) < 1e-12
    for block in ("POS ", "VEL ", "ID  ", "MASS"):
        for pt in (0, 1, 4):
            got = RG.read_block(base, block, [pt])
            want = _expect(parts, block, pt, h.time)
            assert got.dtype == want.dtype and got.shape == want.shape
            np.testing.assert_array_equal(got, want)
    # several species, in list order; [-1] = all of them
    np.testing.assert_array_equal(RG.read_block(base, "POS ", [4, 0]), np.concatenate([parts[4][0], parts[0][0]]))
    np.testing.assert_array_equal(RG.read_block(base, "POS ", [-1]), np.concatenate([parts[t][0] for t in (0, 1, 4)]))
    # read_field: ONE file
    sub = RG.read_field(names[-1], "POS ", 1)
    np.testing.assert_array_equal(sub, np.array_split(parts[1][0], nfiles)[-1])
    assert RG.read_field(names[0], "MASS", 1).dtype == np.float32


def test_records_longer_than_the_marker_range(tmp_path, parts, monkeypatch):
    """A block of 4 GiB or more wraps its 32-bit Fortran markers (the POS block of one 2048^3 file is 103 GB).  Same
    logic with the modulus shrunk to 4096 bytes, so that every particle block of a small file wraps several times."""
    monkeypatch.setattr(GW, "MARKER_MOD", 4096)
    monkeypatch.setattr(RG.SnapFile, "_MARKER_MOD", 4096)
    base = str(tmp_path / "snap_wrap")
    GW.write_snapshot(base, parts, MASSTABLE, BOX, Z, 1, 1, "<")
    for pt in (0, 1, 4):
        for block in ("POS ", "VEL ", "ID  "):
            np.testing.assert_array_equal(RG.read_block(base, block, [pt]), _expect(parts, block, pt, 1 / (1 + Z)))


def test_errors(tmp_path, parts):
    with pytest.raises(Exception, match="File not found"):
        RG.header(str(tmp_path / "nothing_here"))
    base = str(tmp_path / "snap")
    GW.write_snapshot(base, parts, MASSTABLE, BOX, Z, 1, 1)
    with pytest.raises(Exception, match="not implemented"):
        RG.read_block(base, "U   ", [0])
    bad = str(tmp_path / "garbage")
    open(bad, "wb").write(b"\x01\x02\x03\x04" * 100)
    with pytest.raises(IOError):
        RG.header(bad)
    trunc = str(tmp_path / "trunc")
    data = open(base, "rb").read()
    open(trunc, "wb").write(data[:len(data) - 7])
    with pytest.raises(IOError):
        RG.read_block(trunc, "MASS", [0])


@pytest.mark.skipif(not ref_loader.extras_available(), reason="oracle/_ref readers not built (needs /root/reference)")
@pytest.mark.parametrize("nfiles", [1, 3])
def test_against_reference_reader_format1(tmp_path, parts, nfiles):
    """The reference's readgadget/readsnap, compiled unmodified.  Format 1 only: under Python 3 its format-2 label
    search compares str with bytes (readsnap.py:113-114) and never matches."""
    ref = ref_loader.load_extras()["readgadget"]
    base = str(tmp_path / "snap_010")
    GW.write_snapshot(base, parts, MASSTABLE, BOX, Z, nfiles, 1)
    hr, hg = ref.header(base), RG.header(base)
    for name in ("time", "redshift", "boxsize", "filenum", "omega_m", "omega_l", "hubble", "cooling", "format", "Hubble"):
        assert getattr(hr, name) == getattr(hg, name), name
    for name in ("massarr", "npart", "nall"):
        np.testing.assert_array_equal(getattr(hr, name), getattr(hg, name))
    for block in ("POS ", "VEL ", "ID  "):
        for types in ([1], [0], [4, 1], [0, 1, 4]):
            a, b = ref.read_block(base, block, types), RG.read_block(base, block, types)
            assert a.dtype == b.dtype and a.shape == b.shape
            np.testing.assert_array_equal(a, b)
    for pt in (0, 4):                                   # individual masses
        np.testing.assert_array_equal(ref.read_block(base, "MASS", [pt]), RG.read_block(base, "MASS", [pt]))
    # header masses: same values; the reference's dtype follows numpy's scalar promotion (float64 under numpy 2)
    np.testing.assert_array_equal(ref.read_block(base, "MASS", [1]).astype(np.float32), RG.read_block(base, "MASS", [1]))
    sub = base if nfiles == 1 else base + ".1"
    np.testing.assert_array_equal(ref.read_field(sub, "POS ", 1), RG.read_field(sub, "POS ", 1))
    np.testing.assert_array_equal(ref.read_field(sub, "VEL ", 4), RG.read_field(sub, "VEL ", 4))


# ---- property test: any mix of species / header masses / sub-files / format / byte order round-trips ----------------
from hypothesis import given, settings, strategies as st, HealthCheck     # noqa: E402


@settings(max_examples=30, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(counts=st.lists(st.integers(0, 40), min_size=6, max_size=6).filter(lambda c: sum(c) > 0),
       header_mass=st.lists(st.booleans(), min_size=6, max_size=6),
       nfiles=st.integers(1, 4), fmt=st.sampled_from([1, 2]), order=st.sampled_from(["<", ">"]),
       redshift=st.sampled_from([0.0, 2.0]))
def test_reader_roundtrip_property(tmp_path_factory, counts, header_mass, nfiles, fmt, order, redshift):
    masstable = np.array([0.37 if h else 0.0 for h in header_mass])
    parts = GW.make_particles(11, counts, BOX, masstable)
    base = str(tmp_path_factory.mktemp("snap") / "s")
    GW.write_snapshot(base, parts, masstable, BOX, redshift, nfiles, fmt, order)
    h = RG.header(base)
    assert list(h.nall) == counts and h.filenum == nfiles and h.format == fmt
    present = [t for t in range(6) if counts[t]]
    for block in ("POS ", "VEL ", "ID  ", "MASS"):
        got = RG.read_block(base, block, present)
        want = np.concatenate([_expect_generic(parts, block, t, h, masstable) for t in present])
        assert got.dtype == want.dtype and got.shape == want.shape, block
        np.testing.assert_array_equal(got, want, err_msg=block)
    # species without particles give empty arrays, not errors
    empty = [t for t in range(6) if not counts[t]]
    if empty:
        assert RG.read_block(base, "POS ", empty[:1]).shape == (0, 3)
    # per-file reads concatenate to the whole
    names = [base] if nfiles == 1 else ["%s.%d" % (base, i) for i in range(nfiles)]
    t = present[0]
    np.testing.assert_array_equal(np.concatenate([RG.read_field(n, "POS ", t) for n in names]), parts[t][0])


def _expect_generic(parts, block, pt, h, masstable):
    pos, vel, ids, mass = parts[pt]
    if block == "POS ":
        return pos
    if block == "VEL ":
        return vel * np.float32(np.sqrt(h.time)) if h.redshift != 0 else vel
    if block == "ID  ":
        return ids
    return mass if mass is not None else np.full(len(pos), np.float32(masstable[pt]), np.float32)
