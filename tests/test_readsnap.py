"""readsnap mirror (pylians_b200/readsnap.py) against the compiled, unmodified reference readsnap.py on format-1
snapshots (its format-2 label search cannot work under Python 3, readsnap.py:113-114), and format 2 / big-endian files
against the format-1 result.  CPU only: host-side I/O."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gadget_writer as GW                              # noqa: E402
from oracle import ref_loader                           # noqa: E402
from pylians_b200 import readsnap as RS                 # noqa: E402

BOX, Z = 30000.0, 1.5
MASSTABLE = np.array([0.0, 0.5, 0.0, 0.0, 0.0, 0.2])
COUNTS = [600, 900, 0, 0, 250, 40]
BLOCKS = ("POS ", "VEL ", "ID  ", "MASS", "U   ", "RHO ")


def _snapshot(tmp_path, nfiles, fmt=1, order="<"):
    parts = GW.make_particles(8, COUNTS, BOX, MASSTABLE)
    rng = np.random.default_rng(2)
    extra = [("U   ", {0: rng.random(COUNTS[0]).astype(np.float32)}), ("RHO ", {0: rng.random(COUNTS[0]).astype(np.float32)})]
    base = str(tmp_path / ("snap_f%d%s%d" % (fmt, "le" if order == "<" else "be", nfiles)))
    GW.write_snapshot(base, parts, MASSTABLE, BOX, Z, nfiles, fmt, order, extra=extra)
    return base, parts, extra


def _all_reads(mod, base, nfiles):
    out = {}
    for block in BLOCKS:
        types = [-1, 0, 1, 4, 5] if block in ("POS ", "VEL ", "ID  ", "MASS") else [-1, 0]
        for pt in types:
            out[(block, pt)] = np.asarray(mod.read_block(base, block, parttype=pt))
    out[("VEL raw", 1)] = np.asarray(mod.read_block(base, "VEL ", parttype=1, physical_velocities=False))
    sub = base if nfiles == 1 else base + ".1"
    out[("POS one file", -1)] = np.asarray(mod.read_block(sub, "POS "))
    out[("MASS one file", 4)] = np.asarray(mod.read_block(sub, "MASS", parttype=4))
    return out


@pytest.mark.skipif(not ref_loader.extras_available(), reason="oracle/_ref readers not built (needs /root/reference)")
@pytest.mark.parametrize("nfiles", [1, 3])
def test_against_reference_readsnap_format1(tmp_path, nfiles):
    ref = ref_loader.load_extras()["readsnap"]
    base, _, _ = _snapshot(tmp_path, nfiles)
    hr, hg = ref.snapshot_header(base), RS.snapshot_header(base)
    for name in ("format", "swap", "time", "redshift", "sfr", "feedback", "cooling", "filenum", "boxsize", "omega_m",
                 "omega_l", "hubble"):
        assert getattr(hr, name) == getattr(hg, name), name
    for name in ("npart", "massarr", "nall"):
        np.testing.assert_array_equal(getattr(hr, name), getattr(hg, name))
    a, b = _all_reads(ref, base, nfiles), _all_reads(RS, base, nfiles)
    for key in a:
        assert a[key].shape == b[key].shape, key
        if key[0].startswith("MASS") and key[1] >= 0 and MASSTABLE[key[1]] > 0:
            # header masses: same values; the reference's dtype follows numpy's scalar promotion (float64 under numpy 2)
            np.testing.assert_array_equal(a[key].astype(np.float32), b[key], err_msg=str(key))
        else:
            assert a[key].dtype == b[key].dtype, key
            np.testing.assert_array_equal(a[key], b[key], err_msg=str(key))
    assert ref.find_block(base if nfiles == 1 else base + ".0", 1, 0, "VEL ", 3) == RS.find_block(
        base if nfiles == 1 else base + ".0", 1, 0, "VEL ", 3)


@pytest.mark.parametrize("fmt,order,nfiles", [(2, "<", 2), (1, ">", 2), (2, ">", 1)])
def test_other_layouts_match_format1(tmp_path, fmt, order, nfiles):
    base1, parts, extra = _snapshot(tmp_path, nfiles)
    base2, _, _ = _snapshot(tmp_path, nfiles, fmt, order)
    a, b = _all_reads(RS, base1, nfiles), _all_reads(RS, base2, nfiles)
    for key in a:
        assert a[key].dtype == b[key].dtype and a[key].shape == b[key].shape, key
        np.testing.assert_array_equal(a[key], b[key], err_msg=str(key))
    # and they are what was written
    np.testing.assert_array_equal(b[("U   ", 0)], extra[0][1][0])
    np.testing.assert_array_equal(b[("POS ", -1)], np.concatenate([parts[t][0] for t in (0, 1, 4, 5)]))
    want_mass = np.concatenate([parts[0][3], np.full(COUNTS[1], np.float32(0.5)), parts[4][3], np.full(COUNTS[5], np.float32(0.2))])
    np.testing.assert_array_equal(b[("MASS", -1)], want_mass)


def test_errors_and_listing(tmp_path, capsys):
    base, _, _ = _snapshot(tmp_path, 1, 2)
    with pytest.raises(ValueError, match="not known"):
        RS.read_block(base, "XYZ ")
    with pytest.raises(ValueError, match="no data for specified particle type"):
        RS.read_block(base, "U   ", parttype=1)
    with pytest.raises(ValueError, match="wrong parttype"):
        RS.read_block(base, "POS ", parttype=7)
    with pytest.raises(IOError, match="file not found"):
        RS.snapshot_header(str(tmp_path / "missing"))
    RS.list_format2_blocks(base)
    RS.read_gadget_header(base)
    text = capsys.readouterr().out
    assert "GADGET FORMAT  2" in text and "POS " in text and "RHO " in text and "Omega_DM" in text
