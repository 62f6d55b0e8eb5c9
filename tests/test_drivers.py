"""Snapshot drivers: MASL.density_field_gadget, PKL.Pk_comp / PKL.Pk_Gadget on the GPU against the
outputs of the reference's own drivers (tests/golden/drivers.npz, written by tests/golden/make_golden.py from the
compiled, unmodified MAS_gadget.py / Pk_snapshot.py) on the same synthetic 3-file snapshot."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gadget_writer as GW                              # noqa: E402
import parity                                           # noqa: E402

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))

SNAP = dict(seed=9, counts=[3000, 5000, 0, 0, 1200, 0], box_kpc=40000.0, redshift=1.0,
            masstable=np.array([0.0, 0.7, 0.0, 0.0, 0.0, 0.0]), nfiles=3)        # = make_golden.DRIVER_SNAPSHOT
DIMS = 16


@pytest.fixture(scope="module")
def gd(golden_dir):
    return np.load(os.path.join(golden_dir, "drivers.npz"))


@pytest.fixture(scope="module")
def parts():
    return GW.make_particles(SNAP["seed"], SNAP["counts"], SNAP["box_kpc"], SNAP["masstable"], clustered=True)


def _write(tmp_path, parts, fmt=1, nfiles=None, order="<"):
    base = str(tmp_path / "snap_005")
    GW.write_snapshot(base, parts, SNAP["masstable"], SNAP["box_kpc"], SNAP["redshift"], nfiles or SNAP["nfiles"], fmt, order)
    return base


def test_snapshot_regenerates_bit_identically(gd, parts):
    """The golden outputs belong to exactly this particle set."""
    got = np.array([float(np.sum(parts[t][0], dtype=np.float64)) for t in (0, 1, 4)])
    np.testing.assert_array_equal(got, gd["pos_checksum"])


def _check_pk_file(got, want, what):
    assert got.shape == want.shape, what
    parity.assert_k_close(got[:, 0], want[:, 0], what + " k")
    parity.assert_exact(got[:, 4], want[:, 4], what + " Nmodes")
    p0 = np.abs(want[:, 1])
    floor = (p0 + np.median(p0))[:, None] * np.array([1.0, 5.0, 9.0])[None, :]
    parity.assert_spec_close(got[:, 1:4], want[:, 1:4], floor, what)


@pytest.mark.gpu
@pytest.mark.parametrize("fmt,order,nfiles", [(1, "<", 3), (2, ">", 2)])
def test_density_field_gadget_matches_reference(tmp_path, gd, parts, fmt, order, nfiles):
    import pylians_b200
    import MAS_library as MASL
    pylians_b200.set_verbose(False)
    base = _write(tmp_path, parts, fmt, nfiles, order)
    for key, args in (("dens_cdm_CIC", ([1], DIMS, "CIC", False, 0, False)),
                      ("dens_cdm_PCS_rsd1", ([1], DIMS, "PCS", True, 1, False)),
                      ("dens_gas_TSC", ([0], DIMS, "TSC", False, 0, False)),
                      ("dens_gas_stars_CIC_rsd2", ([0, 4], DIMS, "CIC", True, 2, False))):
        got = MASL.density_field_gadget(base, *args)
        assert isinstance(got, np.ndarray) and got.dtype == np.float32
        parity.assert_grid_close(got, gd[key], key)
    # verbose leg: total mass printed equals the grid sum
    d, num, num_dev = MASL.density_field_gadget_device(base, [0, 1, 4], DIMS, "CIC", False, 0, True)
    total = float(d.sum(dtype=__import__("torch").float64).item())
    assert abs(total - (num + float(num_dev.item()))) < 1e-5 * total


@pytest.mark.gpu
def test_pk_gadget_matches_reference(tmp_path, gd, parts):
    import pylians_b200
    import Pk_library as PKL
    pylians_b200.set_verbose(False)
    base = _write(tmp_path, parts)
    for tag, types, rsd, axis in (("cdm", [1], False, 0), ("cdm_rs2", [1], True, 2), ("gas_cdm", [0, 1], False, 0),
                                  ("gas_cdm_stars_rs0", [0, 1, 4], True, 0)):
        folder = str(tmp_path / tag)
        os.makedirs(folder)
        assert PKL.Pk_Gadget(base, DIMS, types, rsd, axis, 1, folder) is None
        want = {k.split("__", 1)[1]: gd[k] for k in gd.files if k.startswith("pk_%s__" % tag)}
        assert sorted(os.listdir(folder)) == sorted(want)
        for f, w in want.items():
            _check_pk_file(np.loadtxt(os.path.join(folder, f)), w, "%s/%s" % (tag, f))


@pytest.mark.gpu
def test_pk_comp_total_matter_against_oracle(tmp_path, parts):
    """ptype = -1 (all species, mass weighted).  The reference's own reader fails on this call (readgadget.py:119
    indexes Nall[-1]), so the expected spectrum is composed from the oracle's MA and Pk on the same particles."""
    import pylians_b200
    import Pk_library as PKL
    from oracle import pylians_oracle as O                  # checker only
    pylians_b200.set_verbose(False)
    base = _write(tmp_path, parts)
    PKL.Pk_comp(base, -1, DIMS, False, 0, 1, str(tmp_path))
    got = np.loadtxt(str(tmp_path / "Pk_matter_z=1.000.dat"))
    box = SNAP["box_kpc"] / 1e3
    pos = np.concatenate([parts[t][0] for t in (0, 1, 4)]) / np.float32(1e3)
    M = np.concatenate([parts[0][3] * np.float32(1e10), np.full(SNAP["counts"][1], np.float32(0.7e10), np.float32),
                        parts[4][3] * np.float32(1e10)])
    delta = np.zeros((DIMS,) * 3, np.float32)
    O.MA(pos, delta, box, "CIC", W=M)
    delta /= np.float32(np.sum(M, dtype=np.float64) / DIMS ** 3)
    delta -= np.float32(1.0)
    ref = O.Pk(delta, box, 0, "CIC", 1)
    _check_pk_file(got, np.transpose([ref.k3D, ref.Pk[:, 0], ref.Pk[:, 1], ref.Pk[:, 2], ref.Nmodes3D]), "Pk_matter")


@pytest.mark.gpu
def test_distributed_pk_comp_single_rank_matches_reference(tmp_path, gd, parts):
    """pylians_b200.dist.SlabPk.pk_comp with one rank (no process group): the slab pipeline fed by the streamed reader
    must write the same file as the reference's Pk_Gadget (world-2 is covered on CPU by tests/test_dist_cpu.py)."""
    import pylians_b200
    from pylians_b200.dist import SlabPk
    pylians_b200.set_verbose(False)
    base = _write(tmp_path, parts)
    for exchange, rsd, axis, fname, key in (
            ("grid", False, 0, "Pk_CDM_z=1.000.dat", "pk_cdm__Pk_CDM_z=1.000.dat"),
            ("particles", True, 2, "Pk_CDM_RS_axis=2_z=1.000.dat", "pk_cdm_rs2__Pk_CDM_RS_axis=2_z=1.000.dat")):
        out = SlabPk(DIMS, SNAP["box_kpc"] / 1e3, "CIC", axis, exchange=exchange).pk_comp(base, 1, rsd, str(tmp_path))
        assert out.Nmodes3D.shape == (gd[key].shape[0],)
        _check_pk_file(np.loadtxt(str(tmp_path / fname)), gd[key], fname)
