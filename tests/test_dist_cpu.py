"""CPU, world_size 2, gloo: the slab-decomposition host logic of pylians_b200.dist (partitioning,
reduce-scatter, transpose all-to-all, windowed binning, all-reduce) against the single-process oracle.
Local kernels are replaced by the numpy stand-in in tests/cpu_slab_ops.py; collectives are real."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, dims, axis, mas, xmode, exchange, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import parity
        from cpu_slab_ops import CpuOps
        from oracle import pylians_oracle as O
        from pylians_b200.dist import SlabPk
        box = 500.0
        rng = np.random.default_rng(7)
        pos = (rng.random((6 * dims ** 3, 3)) * box).astype(np.float32)
        pos2 = (rng.random((5 * dims ** 3, 3)) * box).astype(np.float32)
        W2 = (rng.random(len(pos2)) + 0.5).astype(np.float32)
        eng = SlabPk(dims, box, mas, axis, ops=CpuOps(), exchange=exchange)
        # single-process oracle on the full particle set
        def field(p, m, w=None):
            d = np.zeros((dims,) * 3, np.float32); O.MA(p, d, box, m, W=w)
            d /= np.mean(d, dtype=np.float64); d -= 1.0
            return d
        # Sharding the particles changes the fp32 summation order of the grid (~1e-7 per cell); the MAS
        # deconvolution amplifies that by up to (pi/2)^(2p) = 15x (CIC) .. 226x (PCS) in amplitude at the
        # Nyquist corner, so single-mode corner bins move by up to ~2e-4: this comparison uses 1e-3 (a layout bug gives O(1)).  Exact layout parity is test_slab_layout_single_process.
        if not xmode:
            got = eng.run(pos[rank::world])                       # each rank deposits its own shard
            parity.check_pk(got, O.Pk(field(pos, mas), box, axis, mas, 1), rtol=1e-3)
        else:
            got = eng.run_x([pos[rank::world], pos2[rank::world]], [None, W2[rank::world]], [mas, "PCS"])
            parity.check_xpk(got, O.XPk([field(pos, mas), field(pos2, "PCS", W2)], box, axis, [mas, "PCS"], 1), rtol=1e-3)
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["grid", "particles", "auto"])
@pytest.mark.parametrize("dims,axis,mas,xmode", [(16, 2, "CIC", False), (16, 0, "CIC", False), (16, 2, "CIC", True)])
def test_slab_pipeline_world2_gloo(dims, axis, mas, xmode, exchange):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, dims, axis, mas, xmode, exchange, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", "rank %d failed:\n%s" % (rank, msg)


def test_slab_layout_single_process():
    """G=1 degenerate case of the same code path (no process group): pack/fft_x/bin window = whole cube."""
    sys.path.insert(0, HERE)
    import parity
    from cpu_slab_ops import CpuOps
    from oracle import pylians_oracle as O
    from pylians_b200.dist import SlabPk
    dims, box = 12, 300.0
    rng = np.random.default_rng(3)
    pos = (rng.random((4 * dims ** 3, 3)) * box).astype(np.float32)
    d = np.zeros((dims,) * 3, np.float32); O.MA(pos, d, box, "PCS"); d /= np.mean(d, dtype=np.float64); d -= 1.0
    parity.check_pk(SlabPk(dims, box, "PCS", 1, ops=CpuOps(), exchange="grid").run(pos), O.Pk(d, box, 1, "PCS", 1))
    parity.check_pk(SlabPk(dims, box, "PCS", 1, ops=CpuOps(), exchange="particles").run(pos), O.Pk(d, box, 1, "PCS", 1))


def _worker_uneven(rank, world, port, q):
    """Particle exchange with 4 ranks, very uneven shards (one rank holds 5 particles, all in one slab) and a chunk
    count that divides nothing: the per-piece send/recv sizes must still agree on both ends."""
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import parity
        from cpu_slab_ops import CpuOps
        from oracle import pylians_oracle as O
        from pylians_b200.dist import SlabPk
        dims, box, mas = 16, 500.0, "PCS"
        rng = np.random.default_rng(17)
        pos = (rng.random((3 * dims ** 3, 3)) * box).astype(np.float32)
        pos[-5:, 0] = box * 0.9                                   # the last rank's 5 particles all go to the last slab
        bounds = [0, 7001, 7001 + 3, len(pos) - 5, len(pos)]      # 7001 / 3 / ~5279 / 5 particles
        W = (rng.random(len(pos)) + 0.5).astype(np.float32)
        for chunks in (3, 1, 7):
            eng = SlabPk(dims, box, mas, 2, ops=CpuOps(), exchange="particles", exchange_chunks=chunks)
            slab = eng.density_slab(pos[bounds[rank]:bounds[rank + 1]], W[bounds[rank]:bounds[rank + 1]], overdensity=False)
            ref = np.zeros((dims,) * 3, np.float32); O.MA(pos, ref, box, mas, W=W)
            nxl = dims // world
            parity.assert_grid_close(slab.numpy(), ref[rank * nxl:(rank + 1) * nxl], "slab of rank %d, %d pieces" % (rank, chunks),
                                     rtol=1e-4)
        q.put((rank, "ok"))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def _worker_batches(rank, world, port, q):
    """A rank's particles handed over as ParticleBatches (deposited batch after batch into one window / partial grid)
    give the slab of a single call, in both exchange modes."""
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import parity
        from cpu_slab_ops import CpuOps
        from pylians_b200.dist import SlabPk, ParticleBatches
        dims, box, mas = 16, 500.0, "TSC"
        rng = np.random.default_rng(23)
        pos = (rng.random((4 * dims ** 3, 3)) * box).astype(np.float32)[rank::world]
        W = (rng.random(len(pos)) + 0.5).astype(np.float32)
        cuts = [0, 1000, 1001, len(pos) // 2, len(pos)]           # uneven batches, one of a single particle
        for mode in ("grid", "particles", "auto"):
            eng = SlabPk(dims, box, mas, 2, ops=CpuOps(), exchange=mode)
            one = eng.density_slab(pos, W, overdensity=False)
            pb = ParticleBatches(lambda i: (pos[cuts[i]:cuts[i + 1]], W[cuts[i]:cuts[i + 1]]), len(cuts) - 1, len(pos))
            many = eng.density_slab(pb, None, overdensity=False)
            parity.assert_grid_close(many.numpy(), one.numpy(), "batched vs single deposit, %s" % mode, rtol=1e-5)
        q.put((rank, "ok"))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_particle_batches_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_batches, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", "rank %d failed:\n%s" % (rank, msg)


def test_particle_exchange_pieces_world4_uneven_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_uneven, args=(r, 4, port, q)) for r in range(4)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", "rank %d failed:\n%s" % (rank, msg)


def _worker_snapshot(rank, world, port, base, folder, q):
    """Distributed Pk_comp: each rank reads its own sub-files; rank 0 writes the reference's output file."""
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cpu_slab_ops import CpuOps
        from pylians_b200.dist import SlabPk
        for exchange, rsd, axis in (("grid", False, 0), ("particles", True, 2)):
            eng = SlabPk(16, 40.0, "CIC", axis, ops=CpuOps(), exchange=exchange)
            eng.pk_comp(base, 1, rsd, folder)
        q.put((rank, "ok"))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_distributed_pk_comp_matches_reference_driver(tmp_path, golden_dir):
    """3 sub-files over 2 ranks (rank 0 reads files 0 and 2, rank 1 reads file 1) against the output files of the
    reference's own Pk_Gadget on the same snapshot (tests/golden/drivers.npz); 1e-3 because sharding changes the
    fp32 summation order of the grid (see _worker)."""
    sys.path.insert(0, HERE)
    import gadget_writer as GW
    import parity
    from test_drivers import SNAP
    parts = GW.make_particles(SNAP["seed"], SNAP["counts"], SNAP["box_kpc"], SNAP["masstable"], clustered=True)
    base = str(tmp_path / "snap_005")
    GW.write_snapshot(base, parts, SNAP["masstable"], SNAP["box_kpc"], SNAP["redshift"], SNAP["nfiles"], 1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_snapshot, args=(r, 2, port, base, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", "rank %d failed:\n%s" % (rank, msg)
    gd = np.load(os.path.join(golden_dir, "drivers.npz"))
    for fname, key in (("Pk_CDM_z=1.000.dat", "pk_cdm__Pk_CDM_z=1.000.dat"),
                       ("Pk_CDM_RS_axis=2_z=1.000.dat", "pk_cdm_rs2__Pk_CDM_RS_axis=2_z=1.000.dat")):
        got, want = np.loadtxt(str(tmp_path / fname)), gd[key]
        parity.assert_exact(got[:, 4], want[:, 4], fname + " Nmodes")
        parity.assert_k_close(got[:, 0], want[:, 0], fname + " k")
        p0 = np.abs(want[:, 1])
        floor = (p0 + np.median(p0))[:, None] * np.array([1.0, 5.0, 9.0])[None, :]
        parity.assert_spec_close(got[:, 1:4], want[:, 1:4], floor, fname, rtol=1e-3)


def test_host_chunk_schedule_covers_every_particle_once():
    """The H2D chunk schedule (MAS_library._chunk_bounds): contiguous, complete, at most `chunk` long, and the last full
    chunk tapers into pieces no smaller than the floor."""
    from pylians_b200.MAS_library import _chunk_bounds
    for npart, chunk, floor in [(0, 10, 2), (5, 10, 2), (10, 10, 2), (100, 30, 4), (96, 32, 4), (96, 32, 100), (2 ** 30, 2 ** 28, 2 ** 24),
                                (2 ** 30 + 12345, 2 ** 28, 2 ** 24), (1000, 7, 1)]:
        b = _chunk_bounds(npart, chunk, floor)
        assert (b[0][0] == 0 and b[-1][1] == npart) if npart else b == []
        assert all(x[1] == y[0] for x, y in zip(b[:-1], b[1:]))
        assert all(0 < hi - lo <= chunk for lo, hi in b)
        full = [hi - lo for lo, hi in b if hi - lo == chunk]
        if len(b) >= 2 and chunk // 4 >= floor and npart % chunk == 0:
            assert [hi - lo for lo, hi in b[-3:]] == [chunk // 2, chunk // 4, chunk - chunk // 2 - chunk // 4] and len(full) == npart // chunk - 1
    assert _chunk_bounds(2 ** 30, 2 ** 28, 2 ** 24)[-3:] == [(3 * 2 ** 28, 3 * 2 ** 28 + 2 ** 27), (3 * 2 ** 28 + 2 ** 27, 3 * 2 ** 28 + 3 * 2 ** 26),
                                                           (3 * 2 ** 28 + 3 * 2 ** 26, 2 ** 30)]
