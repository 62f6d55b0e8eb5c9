#!/usr/bin/env python
"""Benchmark of the density-field -> power-spectrum hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one snapshot: deposit Np particles (MASL.MA) -> overdensity -> PKL.Pk.
  N = 1   workload = BASELINE configs[1]: 512^3 uniform particles, CIC, 512^3 grid, real-space Pk.
  N > 1   weak scaling: 512^3 particles per GPU, sharded; grid side chosen so cells ~ particles
          (640 / 800 / 1024 for N = 2 / 4 / 8); slab-decomposed FFT, see pylians_b200/dist.py.
Prints ONE JSON line (rank 0).  `value` = particles/s with inputs resident in HBM; `e2e` = same metric
through the public API with the particle array in pinned HOST memory (H2D inside the timed region,
D2H of the spectra).  The reference arm (--impl reference) times the reference's own CPU code
(oracle/_ref, the unmodified Cython/C compiled from its sources) on a bounded sample.
"""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (particles per GPU side, MAS, axis, description)
    "cfg2_512_cic": dict(nside=512, mas="CIC", axis=2),
    "cfg1_128_cic": dict(nside=128, mas="CIC", axis=2),
    "cfg3_1024_tsc": dict(nside=1024, mas="TSC", axis=2),
    "256_pcs": dict(nside=256, mas="PCS", axis=2),
    "cfg4_1024pg_pcs": dict(nside=1024, mas="PCS", axis=2),   # 8 GPUs -> 2048^3 particles onto a 2048^3 grid
}
GRID_FOR_GPUS = {1: 1.0, 2: 1.25, 4: 1.5625, 8: 2.0}     # grid side multiplier: 512 -> 640 / 800 / 1024
BOX = 1000.0


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """SM clock and throttle reasons sampled DURING the timed region: one NVML query pair per step, issued
    inline from the timing loop while the GPU is busy (each query costs ~1 us).  A sampler *thread* or an
    external `nvidia-smi -lms` process perturbs short runs by milliseconds per step (GIL hand-offs / process
    start-up: profiles/loop_overhead.py), so neither is used."""

    def __init__(self, index):
        self.index = index
        self.sm, self.reasons, self.max_mhz, self._ok = [], set(), None, False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.names = {"hw_slowdown": pynvml.nvmlClocksThrottleReasonHwSlowdown,
                          "hw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                          "sw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonSwThermalSlowdown,
                          "sw_power_cap": pynvml.nvmlClocksThrottleReasonSwPowerCap}
            self._ok = True
        except Exception:
            self._ok = False

    def sample(self):
        if not self._ok:
            return
        try:
            self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
            r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for n, bit in self.names.items():
                if r & bit:
                    self.reasons.add(n)
        except Exception:
            pass

    def stop(self):
        if not self._ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "samples": len(self.sm), "reasons": sorted(self.reasons)}


# --------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own CPU implementation (oracle/_ref) or, if that was
# never built, the oracle port.  Checker code is only ever TIMED here, never used by the product path.
# --------------------------------------------------------------------------------------------------
def cpu_snapshot_fn():
    import contextlib, io
    import numpy as np
    from oracle import ref_loader
    if ref_loader.available():
        MASL, PKL = ref_loader.load()
        kind = "reference"

        def run(pos, dims, mas, axis, threads, deposit="serial"):
            d = np.zeros((dims,) * 3, np.float32)
            # the reference offers two deposits: the serial Cython loop MA() dispatches to, and its OpenMP C kernel
            # (MAS_c.c through <MAS>c3D, all host threads; int-indexed, valid below 1291^3 cells).  cpu_measure()
            # times both once on the sample and keeps the faster one for this host.
            if deposit == "openmp":
                getattr(MASL, mas + "c3D")(pos, d, BOX, threads)
            else:
                MASL.MA(pos, d, BOX, mas)
            d /= np.mean(d, dtype=np.float64); d -= 1.0
            with contextlib.redirect_stdout(io.StringIO()):
                return PKL.Pk(d, BOX, axis, mas, threads)
    else:
        from oracle import pylians_oracle as O
        kind = "port"

        def run(pos, dims, mas, axis, threads, deposit="serial"):
            d = np.zeros((dims,) * 3, np.float32)
            O.MA(pos, d, BOX, mas)
            d /= np.mean(d, dtype=np.float64); d -= 1.0
            return O.Pk(d, BOX, axis, mas, threads)
    return run, kind


def cpu_measure(nside, mas, axis, repeats):
    import numpy as np
    run, kind = cpu_snapshot_fn()
    threads = os.cpu_count() or 1
    rng = np.random.default_rng(1)
    pos = (rng.random((nside ** 3, 3), dtype=np.float32) * np.float32(BOX)).astype(np.float32)
    deposit = "serial"
    if kind == "reference" and nside < 1291:
        v = cpu_deposit_variants(nside, mas, pos) or {}
        omp = [x for k, x in v.items() if k.startswith("openmp_c_") and isinstance(x, float)]
        if omp and omp[0] > v.get("serial_cython_1_core", float("inf")):
            deposit = "openmp"
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        run(pos, nside, mas, axis, threads, deposit)
        times.append(time.perf_counter() - t0)
    return times, kind, threads, deposit


def cpu_deposit_variants(nside, mas, pos=None):
    """The two deposit paths the reference offers, timed alone on the same sample: the serial Cython kernel
    MA() dispatches to, and the OpenMP C kernel (MAS_c.c via <MAS>c3D, all host threads; valid for N < 1291 only)."""
    import numpy as np
    from oracle import ref_loader
    if not ref_loader.available():
        return None
    MASL, _ = ref_loader.load()
    threads = os.cpu_count() or 1
    if pos is None:
        rng = np.random.default_rng(1)
        pos = (rng.random((nside ** 3, 3), dtype=np.float32) * np.float32(BOX)).astype(np.float32)
    out = {"sample": "%d^3 particles %s onto %d^3 grid" % (nside, mas, nside), "unit": "particles/s"}
    d = np.zeros((nside,) * 3, np.float32)
    t0 = time.perf_counter(); MASL.MA(pos, d, BOX, mas); out["serial_cython_1_core"] = nside ** 3 / (time.perf_counter() - t0)
    try:
        fn = getattr(MASL, mas + "c3D")
        d[...] = 0
        t0 = time.perf_counter(); fn(pos, d, BOX, threads); out["openmp_c_%d_threads" % threads] = nside ** 3 / (time.perf_counter() - t0)
    except Exception as e:  # noqa: BLE001
        out["openmp_c_error"] = repr(e)
    return out


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total = args.steps + args.warmup
    nside = 256 if total > 6 else 384                      # bounded sample of the 512^3 workload
    nside = min(nside, wl["nside"])
    times, kind, threads, deposit = cpu_measure(nside, wl["mas"], wl["axis"], total)
    timed = times[args.warmup:] if args.steps > 0 else times
    sec = sum(timed) / max(len(timed), 1)
    val = nside ** 3 / sec
    sample = "%d^3 particles %s onto %d^3 grid + overdensity + Pk (axis=%d): same path, reduced size; particles/s" % (
        nside, wl["mas"], nside, wl["axis"])
    line = {"impl": "reference", "metric": "MA+Pk snapshot throughput", "value": val, "unit": "particles/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic uniform random particles, seed 1",
            "config": workload_config(args, wl),
            "cpu_baseline": {"value": val, "unit": "particles/s", "cores": threads, "kind": kind, "deposit": deposit, "sample": sample,
                             "note": "deposit = the faster on this host of the reference's serial Cython loop (what MA() dispatches to) "
                                     "and its OpenMP C kernel (MAS_c, all threads), see `deposit`; Pk: threads feed the FFT, its mode loop is serial"},
            "e2e": {"value": val, "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "s_per_snapshot_sample": sec}
    print(json.dumps(line))


def zeldovich_particles(nside, box, gen, dev, sigma_cells=2.0, n_s=-1.0):
    """Lattice q = (i+1/2) L/n displaced by psi = IFFT(i k/k^2 delta_k) of a Gaussian field with a power-law
    spectrum k^n_s (the construction of density_field_library.pyx:106-124), scaled to an rms displacement
    of sigma_cells cells per axis, wrapped into [0, L).  Particles stay in lattice (x-major) order."""
    import torch
    n = nside
    k1 = torch.fft.fftfreq(n, d=1.0 / n, device=dev)
    kz = torch.fft.rfftfreq(n, d=1.0 / n, device=dev)
    k2 = k1[:, None, None] ** 2 + k1[None, :, None] ** 2 + kz[None, None, :] ** 2
    k2[0, 0, 0] = 1.0
    amp = k2 ** (n_s / 4.0)
    dk = torch.randn((n, n, n // 2 + 1), dtype=torch.complex64, device=dev, generator=gen) * amp
    dk[0, 0, 0] = 0
    q = (torch.arange(n, device=dev, dtype=torch.float32) + 0.5) * (box / n)
    pos = torch.empty((n, n, n, 3), dtype=torch.float32, device=dev)
    for a, ka in enumerate((k1[:, None, None], k1[None, :, None], kz[None, None, :])):
        psi = torch.fft.irfftn(1j * ka / k2 * dk, s=(n, n, n))
        psi *= sigma_cells * (box / n) / psi.std()
        shape = [1, 1, 1]; shape[a] = n
        pos[..., a] = torch.remainder(q.view(shape) + psi, box)
        del psi
    pos = pos.view(-1, 3)
    pos.clamp_(0.0, float(torch.nextafter(torch.tensor(box, dtype=torch.float32), torch.tensor(0.0))))
    return pos


def workload_config(args, wl):
    n = args.gpus
    gside = int(round(wl["nside"] * GRID_FOR_GPUS.get(n, 1.0)))
    return {"workload": "%s: %d^3 particles per GPU x %d GPU(s), %s onto %d^3 grid, BoxSize=%g, Pk axis=%d (l=0,2,4)" % (
                args.workload, wl["nside"], n, wl["mas"], gside, BOX, wl["axis"]),
            "particles_total": wl["nside"] ** 3 * n, "grid": gside, "mas": wl["mas"], "axis": wl["axis"],
            "l2": "inputs exceed L2 (pos %.2f GB, grid %.2f GB per GPU vs 126 MB L2)" % (
                wl["nside"] ** 3 * 12 / 1e9, gside ** 3 * 4 / 1e9 / n)}


# --------------------------------------------------------------------------------------------------
def run_ours(args, wl):
    import numpy as np
    import torch
    import pylians_b200
    from pylians_b200 import _lib, MAS_library as MASL, Pk_library as PKL
    pylians_b200.set_verbose(False)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    nside, mas, axis = wl["nside"], wl["mas"], wl["axis"]
    npart = nside ** 3
    gside = int(round(nside * GRID_FOR_GPUS.get(world, 1.0)))
    peaks, peak_src = read_peaks()

    # synthetic particles, generated on the device (seed 1 + rank)
    gen = torch.Generator(device=dev); gen.manual_seed(1 + rank)
    if args.data == "uniform":
        pos = torch.rand((npart, 3), device=dev, dtype=torch.float32, generator=gen) * BOX
    else:
        pos = zeldovich_particles(nside, BOX, gen, dev)
    if world > 1:
        from pylians_b200 import dist as pdist
        engine = pdist.SlabPk(gside, BOX, mas, axis)

        def snapshot(p, hook=None):
            slab = engine.density_slab(p)
            if hook is not None:
                hook()                                   # GPU busy with the deposit / exchange just queued
            return engine.pk_from_slab(slab)
    else:
        grid = torch.empty((gside,) * 3, device=dev, dtype=torch.float32)

        def snapshot(p, hook=None):
            grid.zero_()
            MASL.MA(p, grid, BOX, mas)
            if hook is not None:
                hook()                                   # GPU busy with the deposit kernels just queued
            MASL.overdensity(grid)
            return PKL.Pk(grid, BOX, axis, mas, 1)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(fn, steps, sampler=None):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(steps):
            out = fn(sampler.sample) if sampler is not None else fn()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if dist is not None:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    # ---- device-resident arm -----------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        snapshot(pos)
    _lib.timing_enable(True)
    for w in range(4):
        _lib.timing_collect(w)
    sampler = ClockSampler(local_rank); sampler.start()
    l0 = _lib.launch_count()
    ms, pk = timed_loop(lambda hook=None: snapshot(pos, hook), args.steps, sampler)
    launches = _lib.launch_count() - l0
    clocks = sampler.stop()
    ring_ms, ring_n = _lib.timing_collect(_lib.T_RING)
    tile_ms, tile_n = _lib.timing_collect(_lib.T_TILE)
    dir_ms, dir_n = _lib.timing_collect(_lib.T_DIRECT)
    _lib.timing_enable(False)
    ms_step = ms / args.steps
    value = npart * world / (ms_step * 1e-3)

    # ---- per-stage device times (one extra pass, CUDA events on the current stream) ---------------
    stages = {}
    if world == 1:
        def ev():
            e = torch.cuda.Event(enable_timing=True); e.record(); return e
        reps = 3
        acc = {"deposit_ms": 0.0, "overdensity_ms": 0.0, "pk_ms": 0.0}
        for _ in range(reps):
            grid.zero_(); e0 = ev(); MASL.MA(pos, grid, BOX, mas); e1 = ev(); MASL.overdensity(grid); e2 = ev()
            PKL.Pk(grid, BOX, axis, mas, 1); e3 = ev(); torch.cuda.synchronize()
            acc["deposit_ms"] += e0.elapsed_time(e1) / reps; acc["overdensity_ms"] += e1.elapsed_time(e2) / reps
            acc["pk_ms"] += e2.elapsed_time(e3) / reps
        stages = {k: round(v, 4) for k, v in acc.items()}
        stages["deposit_particles_per_s"] = npart / (acc["deposit_ms"] * 1e-3)
        stages["pk_modes_per_s"] = gside * gside * (gside // 2 + 1) / (acc["pk_ms"] * 1e-3)

    # ---- end-to-end arm: particles in pinned host memory, spectra read back ------------------------
    pos_host = torch.empty((npart, 3), dtype=torch.float32, pin_memory=True)
    pos_host.copy_(pos); torch.cuda.synchronize()
    d2h_bytes = 0

    def e2e_step():
        return snapshot(pos_host.to(dev, non_blocking=True)) if world > 1 else snapshot_host()

    def snapshot_host():
        grid.zero_()
        MASL.MA(pos_host, grid, BOX, mas)          # H2D of the particle array happens inside MA
        MASL.overdensity(grid)
        return PKL.Pk(grid, BOX, axis, mas, 1)     # D2H of the bins happens inside Pk

    for _ in range(2):
        e2e_step()
    e2e_steps = max(2, min(args.steps, 5))
    ms_e2e, pk_e = timed_loop(e2e_step, e2e_steps)
    L = PKL.get_layout(gside, 1)
    d2h_bytes = int(L.n_doubles * 8 + L.n_counts * 8)
    e2e_val = npart * world / (ms_e2e / e2e_steps * 1e-3)

    # ---- roofline of the dominant bandwidth-bound kernel (binning ring kernel) --------------------
    nmodes_ring = None
    roof = None
    if ring_n > 0:
        middle = gside // 2
        kz_hi = middle - 1 if gside % 2 == 0 else middle
        rows_local = gside * gside // world
        alg_bytes = 8.0 * rows_local * kz_hi                        # 8 B per complex mode, read once
        achieved = alg_bytes / (ring_ms / ring_n * 1e-3) / 1e9
        kname = "ring2_kernel<phase>" if gside % 2 == 0 else "ring_kernel<1,phase>"
        traffic = None                       # DRAM bytes per launch from the committed ncu --set full capture, if one matches
        try:
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_ncu_traffic.json")) as f:
                t = json.load(f).get(kname, {}).get(str(gside)) if world == 1 else None
            if t:
                traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
        except (OSError, ValueError, KeyError):
            pass
        roof = {"kernel": kname, "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": traffic,
                "peak_source": peak_src, "avg_launch_ms": ring_ms / ring_n, "launches": ring_n,
                "algorithmic_bytes_per_launch": alg_bytes}

    # ---- the deposit (the largest share of the step) against the same HBM roofline ----------------------------
    # Algorithmic bytes: 12 B per particle read + 4 B per cell read + 4 B per cell written (`number` is accumulated
    # into).  The deposit is NOT HBM-bound: its tile kernel is bound by the shared-memory atomic pipe and the sort
    # passes move the payload twice, so this fraction states the distance to an ideal
    # one-pass deposit, it is not a bandwidth the kernels could reach.
    roof_dep = None
    if world == 1 and stages.get("deposit_ms"):
        dep_bytes = 12.0 * npart + 8.0 * gside ** 3
        dep_gbs = dep_bytes / (stages["deposit_ms"] * 1e-3) / 1e9
        roof_dep = {"stage": "MASL.MA (hist + 2 sort passes + deposit_tile_kernel)", "bound": "shared-memory atomics (tile kernel), hbm (sort passes)",
                    "algorithmic_bytes": dep_bytes, "achieved": dep_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": dep_gbs / peaks["hbm_gbs"], "stage_ms": stages["deposit_ms"],
                    "tile_kernel_ms": tile_ms / max(tile_n, 1),
                    "tile_updates_per_s": (npart * {"NGP": 1, "CIC": 8, "TSC": 27, "PCS": 64}[mas] / (tile_ms / tile_n * 1e-3)) if tile_n else None}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cn = min(256, nside)
            times, kind, threads, dep_kind = cpu_measure(cn, mas, axis, 2)
            sec = min(times)
            cpu = {"value": cn ** 3 / sec, "unit": "particles/s", "cores": threads, "kind": kind, "deposit": dep_kind,
                   "sample": "%d^3 particles %s onto %d^3 grid + overdensity + Pk, best of 2 (%.1f s each); deposit = the faster "
                             "on this host of the reference's serial loop and its OpenMP C kernel (see `deposit`), Pk's mode loop is "
                             "serial by construction" % (cn, mas, cn, sec),
                   "deposit_only": cpu_deposit_variants(cn, mas)}
        line = {"metric": "MA+Pk snapshot throughput", "value": value, "unit": "particles/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
                "s_per_snapshot": ms_step * 1e-3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": ("synthetic uniform random particles" if args.data == "uniform" else
                                         "synthetic Zel'dovich-displaced lattice (rms 2 cells, lattice order)") + " generated on device, seed 1+rank",
                "config": workload_config(args, wl), "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e_val, "unit": "particles/s", "h2d_bytes_per_step": int(npart * 12),
                        "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e / e2e_steps,
                        "note": "particle array in pinned host memory -> MASL.MA -> overdensity -> PKL.Pk -> bins on host"},
                "roofline": roof, "roofline_deposit": roof_dep, "cpu_baseline": cpu, "stages": stages,
                "kernels": {"ring_ms": ring_ms / max(ring_n, 1), "tile_ms": tile_ms / max(tile_n, 1),
                            "direct_ms": dir_ms / max(dir_n, 1), "ring_launches": ring_n, "tile_launches": tile_n,
                            "direct_launches": dir_n},
                "check": {"P0_first_bins": [float(x) for x in pk.Pk[:3, 0]], "shot_noise_expected": BOX ** 3 / (npart * world),
                          "Nmodes_sum_ok": bool(pk.Nmodes3D.sum() + 1 == (gside ** 3 - 8) // 2 + 8) if gside % 2 == 0 else None}}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2_512_cic", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--data", default="uniform", choices=["uniform", "zeldovich"])
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
