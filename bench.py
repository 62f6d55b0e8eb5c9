#!/usr/bin/env python
"""Benchmark of the density-field -> power-spectrum hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--data uniform|zeldovich]

One "step" = one snapshot: deposit the particles (MASL.MA) -> overdensity -> PKL.Pk (or PKL.XPk for two fields).
Default workload `cfg4_1024pg_pcs` (BASELINE configs[3], weak scaling): 1024^3 particles PER GPU, PCS, grid side
1024 / 1280 / 1600 / 2048 for N = 1 / 2 / 4 / 8 -- at N = 8 this IS the north-star target (2048^3 particles onto a
2048^3 grid, multipoles l = 0, 2, 4), at N = 1 it is the largest single-GPU configuration.  Other workloads:
cfg1_128_cic, cfg2_512_cic (configs[1]), cfg3_1024_tsc (configs[2]), cfg5_1024_xpk (configs[4]: CDM + weighted gas,
XPk with per-field MAS).  At N = 1 the default run also reports cfg3, cfg5, cfg2 and the Zel'dovich input as `extra`.

Prints ONE JSON line (rank 0).  `value` = particles/s with inputs resident in HBM; `e2e` = same metric through the
public API with the particle arrays in pinned HOST memory (H2D inside the timed region, D2H of the spectra).
`check.parity` compares the CUDA path with the reference's CPU path ON THE SAME PARTICLES (N = 1: the 256^3 sample the
cpu_baseline leg times; N > 1: the NCCL slab pipeline against the single-GPU pipeline on one shared particle set).
The reference arm (--impl reference) times the reference's own CPU code (oracle/_ref, the unmodified Cython/C compiled
from its sources) on a bounded sample and says which size it really ran.
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
    # fields: list of (MAS, weighted) particle sets of nside^3 particles each (per GPU); kind pk | xpk
    "cfg4_1024pg_pcs": dict(nside=1024, fields=[("PCS", False)], axis=2),     # 8 GPUs -> 2048^3 particles onto 2048^3
    "cfg2_512_cic": dict(nside=512, fields=[("CIC", False)], axis=2),
    "cfg1_128_cic": dict(nside=128, fields=[("CIC", False)], axis=2),
    "cfg3_1024_tsc": dict(nside=1024, fields=[("TSC", False)], axis=2),
    "cfg5_1024_xpk": dict(nside=1024, fields=[("CIC", False), ("PCS", True)], axis=2),   # CDM + weighted gas
    # BASELINE configs[3] as written: 2048^3 particles in TOTAL onto a 2048^3 grid at 1 / 2 / 4 / 8 GPUs (strong scaling);
    # a rank whose share does not fit next to its grid generates and deposits it in batches of 2^30 particles
    "cfg4_2048_strong": dict(nside=2048, fields=[("PCS", False)], axis=2, strong=True),
    "512_strong": dict(nside=512, fields=[("PCS", False)], axis=2, strong=True, batch=1 << 24),     # the same code path, small
    "512_pcs": dict(nside=512, fields=[("PCS", False)], axis=2),
    "256_pcs": dict(nside=256, fields=[("PCS", False)], axis=2),
    "256_xpk": dict(nside=256, fields=[("CIC", False), ("PCS", True)], axis=2),
}
DEFAULT_WORKLOAD = "cfg4_1024pg_pcs"
GRID_FOR_GPUS = {1: 1.0, 2: 1.25, 4: 1.5625, 8: 2.0}     # grid side multiplier: 1024 -> 1280 / 1600 / 2048
BOX = 1000.0
STENCIL = {"NGP": 1, "CIC": 8, "TSC": 27, "PCS": 64}


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def mas_names(wl):
    return [m for m, _ in wl["fields"]]


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
# never built, the oracle port.  Checker code is only ever TIMED / COMPARED here, never used by the product path.
# --------------------------------------------------------------------------------------------------
def host_particles(nside, nfields, seed=1):
    """The bounded sample both arms see: uniform random float32 positions (and weights for weighted fields)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(nfields):
        pos = (rng.random((nside ** 3, 3), dtype=np.float32) * np.float32(BOX)).astype(np.float32)
        W = (rng.random(nside ** 3, dtype=np.float32) + np.float32(0.5)).astype(np.float32)
        out.append((pos, W))
    return out


def cpu_snapshot_fn():
    """run(parts, dims, wl, threads, deposit) -> (grids, spectrum object) through the reference's CPU code."""
    import contextlib, io
    import numpy as np
    from oracle import ref_loader
    if ref_loader.available():
        MASL, PKL = ref_loader.load()
        kind = "reference"
    else:
        from oracle import pylians_oracle as O
        MASL = PKL = O
        kind = "port"

    def run(parts, dims, wl, threads, deposit="serial"):
        grids = []
        for (pos, W), (mas, weighted) in zip(parts, wl["fields"]):
            d = np.zeros((dims,) * 3, np.float32)
            # the reference offers two deposits: the serial Cython loop MA() dispatches to, and its OpenMP C kernel
            # (MAS_c.c through <MAS>[W]c3D, all host threads; int-indexed, valid below 1291^3 cells)
            if deposit == "openmp" and kind == "reference":
                if weighted:
                    getattr(MASL, mas + "Wc3D")(pos, d, W, BOX, threads)
                else:
                    getattr(MASL, mas + "c3D")(pos, d, BOX, threads)
            else:
                MASL.MA(pos, d, BOX, mas, W=W if weighted else None)
            raw = d.copy()
            d /= np.mean(d, dtype=np.float64); d -= 1.0
            grids.append((raw, d))
        with contextlib.redirect_stdout(io.StringIO()):
            if len(grids) == 1:
                spec = PKL.Pk(grids[0][1], BOX, wl["axis"], wl["fields"][0][0], threads)
            else:
                spec = PKL.XPk([g[1] for g in grids], BOX, wl["axis"], mas_names(wl), threads)
        return [g[0] for g in grids], spec
    return run, kind


def cpu_deposit_variants(nside, mas, pos):
    """The two deposit paths the reference offers, timed alone on the same sample: the serial Cython kernel
    MA() dispatches to, and the OpenMP C kernel (MAS_c.c via <MAS>c3D, all host threads; valid for N < 1291 only)."""
    import numpy as np
    from oracle import ref_loader
    if not ref_loader.available():
        return None
    MASL, _ = ref_loader.load()
    threads = os.cpu_count() or 1
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


def cpu_measure(nside, wl, repeats, parts=None):
    """Times `repeats` CPU snapshots of the workload at nside^3 particles per field; returns the last result too."""
    run, kind = cpu_snapshot_fn()
    threads = os.cpu_count() or 1
    parts = parts or host_particles(nside, len(wl["fields"]))
    deposit, variants = "serial", None
    if kind == "reference" and nside < 1291:
        variants = cpu_deposit_variants(nside, wl["fields"][0][0], parts[0][0]) or {}
        omp = [x for k, x in variants.items() if k.startswith("openmp_c_") and isinstance(x, float)]
        if omp and omp[0] > variants.get("serial_cython_1_core", float("inf")):
            deposit = "openmp"
    times, res = [], None
    for _ in range(repeats):
        t0 = time.perf_counter()
        res = run(parts, nside, wl, threads, deposit)
        times.append(time.perf_counter() - t0)
    return dict(times=times, kind=kind, threads=threads, deposit=deposit, variants=variants, result=res, parts=parts)


def grid_side(wl, world):
    return wl["nside"] if wl.get("strong") else int(round(wl["nside"] * GRID_FOR_GPUS.get(world, 1.0)))


def rank_particles(wl, world, rank):
    """particles of one field held by `rank`: nside^3 each (weak scaling), or an even share of nside^3 (strong scaling)"""
    n = wl["nside"] ** 3
    if not wl.get("strong"):
        return n
    return n // world + (1 if rank < n % world else 0)


def workload_config(args, wl, world=None, sample_nside=None):
    n = world or args.gpus
    gside = grid_side(wl, n)
    nf = len(wl["fields"])
    strong = bool(wl.get("strong"))
    desc = "%s: %s%d^3 particles %s %d GPU(s), %s onto %d^3 grid, BoxSize=%g, %s axis=%d (l=0,2,4)" % (
        args.workload, "%d x " % nf if nf > 1 else "", wl["nside"], "in total over" if strong else "per GPU x", n,
        "+".join(m + ("W" if w else "") for m, w in wl["fields"]), gside, BOX, "XPk" if nf > 1 else "Pk", wl["axis"])
    cfg = {"workload": desc, "particles_total": wl["nside"] ** 3 * (1 if strong else n) * nf, "grid": gside, "mas": mas_names(wl),
           "axis": wl["axis"], "fields": nf,
           "l2": "inputs exceed L2 (pos %.2f GB per field, grid %.2f GB per GPU vs 126 MB L2)" % (
               wl["nside"] ** 3 * 12 / 1e9, gside ** 3 * 4 / 1e9 / n)}
    if sample_nside is not None:
        cfg["workload"] = ("REFERENCE ARM ran a bounded %d^3-particle / %d^3-grid sample of [%s] on the host CPU; "
                           "value is that sample's particles/s" % (sample_nside, sample_nside, desc))
        cfg["sample_nside"] = sample_nside
        cfg["same_config"] = bool(sample_nside == wl["nside"] and n == 1)
    return cfg


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total = args.steps + args.warmup
    nside = 256 if (total > 6 or len(wl["fields"]) > 1) else 384          # bounded sample of the workload
    nside = min(nside, wl["nside"])
    m = cpu_measure(nside, wl, max(total, 1))
    times = m["times"]
    timed = times[args.warmup:] if args.steps > 0 and len(times) > args.warmup else times
    sec = sum(timed) / max(len(timed), 1)
    nf = len(wl["fields"])
    val = nf * nside ** 3 / sec
    sample = "%d^3 particles per field (%s) onto %d^3 grid + overdensity + %s (axis=%d): same path, reduced size; particles/s" % (
        nside, "+".join(mas_names(wl)), nside, "XPk" if nf > 1 else "Pk", wl["axis"])
    line = {"impl": "reference", "metric": "MA+Pk snapshot throughput", "value": val, "unit": "particles/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic uniform random particles, numpy default_rng seed 1 (host)",
            "config": workload_config(args, wl, sample_nside=nside),
            "cpu_baseline": {"value": val, "unit": "particles/s", "cores": m["threads"], "kind": m["kind"], "deposit": m["deposit"],
                             "sample": sample,
                             "note": "deposit = the faster on this host of the reference's serial Cython loop (what MA() dispatches to) "
                                     "and its OpenMP C kernel (MAS_c, all threads), see `deposit`; Pk: threads feed the FFT, its mode loop is serial"},
            "e2e": {"value": val, "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "s_per_snapshot_sample": sec}
    emit(line)


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
    del amp
    q = (torch.arange(n, device=dev, dtype=torch.float32) + 0.5) * (box / n)
    pos = torch.empty((n, n, n, 3), dtype=torch.float32, device=dev)
    for a, ka in enumerate((k1[:, None, None], k1[None, :, None], kz[None, None, :])):
        psi = torch.fft.irfftn(1j * ka / k2 * dk, s=(n, n, n))
        psi *= sigma_cells * (box / n) / psi.std()
        shape = [1, 1, 1]; shape[a] = n
        pos[..., a] = torch.remainder(q.view(shape) + psi, box)
        del psi
    del dk, k2
    pos = pos.view(-1, 3)
    pos.clamp_(0.0, float(torch.nextafter(torch.tensor(box, dtype=torch.float32), torch.tensor(0.0))))
    return pos


# --------------------------------------------------------------------------------------------------
# parity helpers (numbers for check.parity; the assertions live in tests/)
# --------------------------------------------------------------------------------------------------
def grid_rel(a, b):
    """max |a-b| / (|b| + mean|b|): the grid contract of tests/parity.py is this <= 1e-5."""
    import numpy as np
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / (np.abs(b) + np.mean(np.abs(b)))))


def spec_rel(mine, ref, is_x):
    """max over bins of |P - P_ref| / (|P_ref| + floor) for the 3-D multipoles (floor: monopole level, as tests/parity.py),
    split at the Nyquist frequency (beyond it the MAS deconvolution amplifies fp32 summation-order noise of the grid)."""
    import numpy as np
    P, Q = np.asarray(mine.Pk, np.float64), np.asarray(ref.Pk, np.float64)
    if not is_x:
        P, Q = P[:, :, None], Q[:, :, None]
    p0 = np.abs(Q[:, 0, :])
    floor = (p0 + np.median(p0, axis=0)[None, :])[:, None, :] * np.array([1.0, 5.0, 9.0])[None, :, None]
    rel = np.abs(P - Q) / (np.abs(Q) + floor)
    k = np.asarray(ref.k3D)
    kN = np.pi * len(np.asarray(ref.Nmodes1D)) * 2 / BOX          # Nmodes1D has dims/2 entries
    inside = k <= kN
    out = {"pk_max_rel_below_nyquist": float(rel[inside].max()), "pk_max_rel_all": float(rel.max())}
    if is_x:
        X, Y = np.asarray(mine.XPk, np.float64), np.asarray(ref.XPk, np.float64)
        xf = np.sqrt(floor[:, :, 0] * floor[:, :, 1])[:, :, None]
        out["xpk_max_rel_all"] = float((np.abs(X - Y) / (np.abs(Y) + xf)).max())
    names = ("Nmodes3D", "Nmodes1D", "Nmodes2D", "kpar", "kper")
    out["nmodes_exact"] = bool(all(np.array_equal(np.asarray(getattr(mine, n)), np.asarray(getattr(ref, n))) for n in names))
    out["k3D_max_rel"] = float(np.max(np.abs(np.asarray(mine.k3D) / np.asarray(ref.k3D) - 1.0)))
    return out


# --------------------------------------------------------------------------------------------------
class Pipeline(object):
    """The snapshot step of one workload on this rank's GPU (single GPU: MASL/PKL; several: dist.SlabPk)."""

    def __init__(self, wl, world, dev, dist):
        import torch
        from pylians_b200 import MAS_library as MASL, Pk_library as PKL
        self.wl, self.world, self.dev, self.dist = wl, world, dev, dist
        self.MASL, self.PKL, self.torch = MASL, PKL, torch
        self.nf = len(wl["fields"])
        self.gside = grid_side(wl, world)
        self.axis = wl["axis"]
        self.rank = dist.get_rank() if dist is not None else 0
        self.npart = rank_particles(wl, world, self.rank)
        # a share that does not fit next to the grid is generated and deposited batch by batch
        self.batch = int(wl.get("batch", 1 << 30))
        self.streamed = bool(wl.get("strong")) and (self.npart * 12 > 24e9 or "batch" in wl)
        if world > 1:
            from pylians_b200 import dist as pdist
            self.engine = pdist.SlabPk(self.gside, BOX, wl["fields"][0][0], self.axis)
        else:
            self.grids = [torch.empty((self.gside,) * 3, device=dev, dtype=torch.float32) for _ in range(self.nf)]

    def make_particles(self, data, seed):
        torch = self.torch
        gen = torch.Generator(device=self.dev); gen.manual_seed(seed)
        n = self.wl["nside"]
        parts = []
        for _, weighted in self.wl["fields"]:
            if self.streamed:
                parts.append(self.batches(gen, weighted))
                continue
            if data == "uniform" or self.npart != n ** 3:
                pos = torch.rand((self.npart, 3), device=self.dev, dtype=torch.float32, generator=gen) * BOX
            else:
                pos = zeldovich_particles(n, BOX, gen, self.dev)
            W = (torch.rand(self.npart, device=self.dev, dtype=torch.float32, generator=gen) + 0.5) if weighted else None
            parts.append((pos, W))
        return parts

    def batches(self, gen, weighted, host=None):
        """This rank's share as ParticleBatches: batch i is drawn on the device when asked for (uniform random), or is
        the pinned host buffer `host` again (the end-to-end arm: every batch crosses PCIe, the bytes are what counts)."""
        from pylians_b200.dist import ParticleBatches
        torch, B, n = self.torch, self.batch, self.npart
        nb = (n + B - 1) // B

        def make(i):
            m = min(B, n - i * B)
            if host is not None:
                return host[0][:m], (host[1][:m] if weighted else None)
            pos = torch.rand((m, 3), device=self.dev, dtype=torch.float32, generator=gen) * BOX
            W = (torch.rand(m, device=self.dev, dtype=torch.float32, generator=gen) + 0.5) if weighted else None
            return pos, W
        return ParticleBatches(make, nb, n)

    def snapshot(self, parts, hook=None):
        """parts: list of (pos, W) per field, device or pinned-host tensors."""
        MASL, PKL = self.MASL, self.PKL
        if self.world > 1:
            eng = self.engine
            if self.nf == 1:
                (mas, _) = self.wl["fields"][0]
                pos, W = (parts[0], None) if self.streamed else parts[0]
                slab = eng.density_slab(pos, W, mas)
                if hook is not None:
                    hook()                               # GPU busy with the deposit / exchange just queued
                return eng.pk_from_slab(slab, mas)
            return eng.run_x([p for p, _ in parts], [w for _, w in parts], mas_names(self.wl))
        for g, part, (mas, _) in zip(self.grids, parts, self.wl["fields"]):
            g.zero_()
            for pos, W in (part if self.streamed else [part]):
                MASL.MA(pos, g, BOX, mas, W=W)           # host tensors: the H2D copy happens inside MA
                if hook is not None:
                    hook(); hook = None                  # GPU busy with the deposit kernels just queued
            MASL.overdensity(g)
        if self.nf == 1:
            return PKL.Pk(self.grids[0], BOX, self.axis, self.wl["fields"][0][0], 1)     # D2H of the bins inside Pk
        return PKL.XPk(self.grids, BOX, self.axis, mas_names(self.wl), 1)


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
        # stdout carries ONE JSON line: whatever NCCL logs (its version banner under NCCL_DEBUG=VERSION) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    peaks, peak_src = read_peaks()

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

    def measure(wl_, data, steps, warm, full):
        """Device-resident timing of one workload; `full` adds the e2e arm, the stage breakdown and the kernel brackets."""
        pipe = Pipeline(wl_, world, dev, dist)
        nf, gside = pipe.nf, pipe.gside
        npart = wl_["nside"] ** 3 / (world if wl_.get("strong") else 1)      # particles per field per rank (average)
        parts = pipe.make_particles(data, 1 + rank)
        for _ in range(max(warm, 3)):
            pipe.snapshot(parts)
        _lib.timing_enable(True)
        for w in range(7):
            _lib.timing_collect(w)
        sampler = ClockSampler(local_rank); sampler.start()
        l0 = _lib.launch_count()
        ms, spec = timed_loop(lambda hook=None: pipe.snapshot(parts, hook), steps, sampler)
        launches = _lib.launch_count() - l0
        clocks = sampler.stop()
        T = {name: _lib.timing_collect(idx) for name, idx in (("ring", _lib.T_RING), ("tile", _lib.T_TILE), ("direct", _lib.T_DIRECT),
                                                              ("bin", _lib.T_BIN), ("fft", _lib.T_FFT), ("sort", _lib.T_SORT))}
        _lib.timing_enable(False)
        ms_step = ms / steps
        res = {"ms_per_step": ms_step, "value": nf * npart * world / (ms_step * 1e-3), "launches": int(launches), "clocks": clocks,
               "grid": gside, "spec": spec,
               "kernels_ms_per_step": {k: (v[0] / steps) for k, v in T.items()},
               "kernel_launches_per_step": {k: v[1] / steps for k, v in T.items()}}
        if not full:
            del parts, pipe
            torch.cuda.empty_cache()
            return res
        # ---- per-stage device times (one extra pass, CUDA events on the current stream) ---------------
        if world == 1:
            def ev():
                e = torch.cuda.Event(enable_timing=True); e.record(); return e
            reps = 3
            acc = {"deposit_ms": 0.0, "overdensity_ms": 0.0, "pk_ms": 0.0}
            for _ in range(reps):
                for g in pipe.grids:
                    g.zero_()
                e0 = ev()
                for g, part, (mas, _) in zip(pipe.grids, parts, wl_["fields"]):
                    for pos, W in (part if pipe.streamed else [part]):
                        MASL.MA(pos, g, BOX, mas, W=W)
                e1 = ev()
                for g in pipe.grids:
                    MASL.overdensity(g)
                e2 = ev()
                if nf == 1:
                    PKL.Pk(pipe.grids[0], BOX, pipe.axis, wl_["fields"][0][0], 1)
                else:
                    PKL.XPk(pipe.grids, BOX, pipe.axis, mas_names(wl_), 1)
                e3 = ev(); torch.cuda.synchronize()
                acc["deposit_ms"] += e0.elapsed_time(e1) / reps; acc["overdensity_ms"] += e1.elapsed_time(e2) / reps
                acc["pk_ms"] += e2.elapsed_time(e3) / reps
            st = {k: round(v, 4) for k, v in acc.items()}
            st["deposit_particles_per_s"] = nf * npart / (acc["deposit_ms"] * 1e-3)
            st["pk_grid_cells_per_s"] = nf * gside ** 3 / (acc["pk_ms"] * 1e-3)
            res["stages"] = st
        # ---- end-to-end arm: particles in pinned host memory, spectra read back ------------------------
        host = []
        if pipe.streamed:
            # one pinned batch buffer per field, sent again for every batch: each batch crosses PCIe inside the timed region
            gen = torch.Generator(device=dev); gen.manual_seed(99 + rank)
            for _, weighted in wl_["fields"]:
                m = min(pipe.batch, pipe.npart)
                hp = torch.empty((m, 3), dtype=torch.float32, pin_memory=True)
                hp.copy_(torch.rand((m, 3), device=dev, dtype=torch.float32, generator=gen) * BOX)
                hw = torch.empty(m, dtype=torch.float32, pin_memory=True); hw.fill_(1.0)
                host.append(pipe.batches(None, weighted, host=(hp, hw)))
            h2d = sum((16 if weighted else 12) * pipe.npart for _, weighted in wl_["fields"])
        else:
            for pos, W in parts:
                hp = torch.empty(pos.shape, dtype=torch.float32, pin_memory=True); hp.copy_(pos)
                hw = None
                if W is not None:
                    hw = torch.empty(W.shape, dtype=torch.float32, pin_memory=True); hw.copy_(W)
                host.append((hp, hw))
            h2d = sum(p.numel() * 4 + (w.numel() * 4 if w is not None else 0) for p, w in host)
        torch.cuda.synchronize()

        def e2e_step():
            return pipe.snapshot(host)           # host tensors: MASL.MA / SlabPk stream them in chunks, H2D overlapped with the deposit

        for _ in range(2):
            e2e_step()
        e2e_steps = max(2, min(steps, 5))
        ms_e2e, _ = timed_loop(e2e_step, e2e_steps)
        L = PKL.get_layout(gside, nf)
        res["e2e"] = {"value": nf * npart * world / (ms_e2e / e2e_steps * 1e-3), "unit": "particles/s",
                      "h2d_bytes_per_step": int(h2d * world) if not wl_.get("strong") else int(sum((16 if w_ else 12) for _, w_ in wl_["fields"]) * wl_["nside"] ** 3), "d2h_bytes_per_step": int((L.n_doubles + L.n_counts) * 8 * world),
                      "ms_per_step": ms_e2e / e2e_steps,
                      "note": "particle arrays in pinned host memory on every rank -> MASL.MA / SlabPk (H2D in chunks on a side stream, "
                              "overlapped with the deposit of the previous chunk) -> overdensity -> PKL.Pk/XPk -> bins on host; byte "
                              "counts are totals over all ranks"}
        del parts, host, pipe
        torch.cuda.empty_cache()
        return res

    main = measure(wl, args.data, args.steps, args.warmup, True)
    nf, gside, spec = len(wl["fields"]), main["grid"], main["spec"]
    npart = wl["nside"] ** 3 / (world if wl.get("strong") else 1)
    steps = args.steps
    K = main["kernels_ms_per_step"]

    # ---- roofline of the fused binning (the bandwidth-bound kernel the north_star names) ------------------------
    # Algorithmic bytes: every stored complex mode of every field read exactly once = 8 F N^2 (N/2+1) / world per rank.
    # The bracket (PYLB_T_BIN) covers the whole pylb_pk_bin call: bin zeroing, MAS table, the self-conjugate columns
    # (special kernel), the ring kernel and its finish pass.  `ring_kernel_only` is the ring kernel alone.
    roof = None
    nbin = main["kernel_launches_per_step"]["bin"]
    if nbin > 0:
        alg_bytes = 8.0 * nf * gside * gside * (gside // 2 + 1) / world
        bin_ms = K["bin"] / nbin
        achieved = alg_bytes / (bin_ms * 1e-3) / 1e9
        ring_n = main["kernel_launches_per_step"]["ring"]
        ring_only = None
        if ring_n > 0:
            kz_hi = gside // 2 - 1 if gside % 2 == 0 else gside // 2
            rb = 8.0 * nf * gside * gside * kz_hi / world
            ring_only = {"avg_launch_ms": K["ring"] / ring_n, "algorithmic_bytes_per_launch": rb,
                         "achieved": rb / (K["ring"] / ring_n * 1e-3) / 1e9, "frac": rb / (K["ring"] / ring_n * 1e-3) / 1e9 / peaks["hbm_gbs"]}
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")) as f:
                t = json.load(f).get("binning", {}).get("%s_%d" % (args.workload, world))
            if t:
                traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
        except (OSError, ValueError, KeyError):
            pass
        roof = {"kernel": "fused binning (pylb_pk_bin: special + ring kernel + finish), F=%d" % nf, "bound": "hbm", "achieved": achieved,
                "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": traffic,
                "peak_source": peak_src, "avg_launch_ms": bin_ms, "launches": nbin * steps,
                "algorithmic_bytes_per_launch": alg_bytes, "ring_kernel_only": ring_only}
    # cuFFT (library) reported separately: R2C reads 4 B and writes ~4 B per cell at least once
    fft = None
    if main["kernel_launches_per_step"]["fft"] > 0:
        fb = 8.0 * nf * gside ** 3 / world
        fft = {"ms_per_step": K["fft"], "calls_per_step": main["kernel_launches_per_step"]["fft"], "min_bytes_per_step": fb,
               "achieved_GBs_vs_one_pass": fb / (K["fft"] * 1e-3) / 1e9, "note": "cuFFT; a 3-D transform makes ~3 passes over the data"}

    # ---- the deposit against the same HBM roofline ------------------------------------------------------------
    # Algorithmic bytes: 12 B per particle (+4 with weights) read + 4 B per cell read + 4 B per cell written (`number` is
    # accumulated into).  The sort passes are HBM-bound (they move the payload twice); the tile kernel is bound by
    # shared-memory atomics / issue slots, not HBM -- so `frac` states the distance to an ideal one-pass deposit.
    roof_dep = None
    st = main.get("stages")
    if st is None and world > 1 and K["tile"] > 0:
        # N > 1: no separate stage pass; rank 0's sort + tile brackets of the windowed deposits (the partition pass and the
        # exchange are outside these brackets)
        st = {"deposit_ms": K["sort"] + K["tile"], "note": "rank 0: sort + tile-kernel brackets of the windowed deposits only"}
    if st and K["tile"] > 0:
        dep_bytes = sum((16.0 if w else 12.0) for _, w in wl["fields"]) * npart + 8.0 * nf * gside ** 3 / world
        dep_gbs = dep_bytes / (st["deposit_ms"] * 1e-3) / 1e9
        updates = sum(STENCIL[m] for m in mas_names(wl)) * npart
        sort_bytes = sum((12.0 + (16.0 if w else 12.0) + 16.0 + 16.0 + 16.0 + (4.0 if w else 0.0)) for _, w in wl["fields"]) * npart
        roof_dep = {"stage": "MASL.MA (histogram + 2 counting-sort passes + tile kernel)", "bound": "hbm (sort passes); shared-memory atomics / issue (tile kernel)",
                    "algorithmic_bytes": dep_bytes, "achieved": dep_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": dep_gbs / peaks["hbm_gbs"], "stage_ms": st["deposit_ms"],
                    "sort_ms_per_step": K["sort"], "sort_traffic_bytes": sort_bytes,
                    "sort_achieved_GBs": sort_bytes / (K["sort"] * 1e-3) / 1e9 if K["sort"] > 0 else None,
                    "tile_kernel_ms_per_step": K["tile"], "tile_kernel_launches_per_step": main["kernel_launches_per_step"]["tile"],
                    "tile_updates_per_s": updates / (K["tile"] * 1e-3)}
        if "note" in st:
            roof_dep["note"] = st["note"]
        # atomic / L2 / LSU-pipe figures of the deposit kernels from the committed `ncu --set full` captures (512^3 launches)
        try:
            with open(os.path.join(ROOT, "profiles", "r2_ncu_deposit.json")) as f:
                keep = ("duration", "duration_unit", "l1tex__data_pipe_lsu_wavefronts_pct", "shared_atomic_instructions",
                        "shared_atomic_wavefronts", "shared_bank_conflicts", "l2_throughput_pct", "dram_throughput_pct",
                        "issue_active_pct", "warp_instructions", "achieved_occupancy_pct")
                roof_dep["ncu"] = {name: [{k: l[k] for k in keep if k in l} for l in launches]
                                   for name, launches in json.load(f).items() if not name.startswith(("dropped", "earlier"))}
        except (OSError, ValueError):
            pass

    # ---- parity on shared particles ---------------------------------------------------------------------------
    cpu = None
    parity = None
    if world == 1:
        if not args.no_cpu_baseline:
            parity, cpu = parity_single(args, wl, dev, rank)
    else:
        parity = parity_multi(args, wl, dev, rank, world, dist)

    # ---- other BASELINE configs and the Zel'dovich input, as extra keys of the same line -----------------------
    extra = {}
    if args.extras and args.workload == DEFAULT_WORKLOAD:
        todo = [("zeldovich", wl, "zeldovich")]
        if world == 1:
            todo += [(n, WORKLOADS[n], "uniform") for n in ("cfg3_1024_tsc", "cfg5_1024_xpk", "cfg2_512_cic")]
        for name, w2, data in todo:
            try:
                r = measure(w2, data, 3, 3, False)
                extra[name] = {"ms_per_step": r["ms_per_step"], "particles_per_s": r["value"], "grid": r["grid"], "steps": 3,
                               "kernels_ms_per_step": r["kernels_ms_per_step"], "gpu_launches": r["launches"],
                               "data": data, "mas": mas_names(w2)}
            except Exception as e:  # noqa: BLE001
                extra[name] = {"error": repr(e)}

    if rank == 0:
        ms_step = main["ms_per_step"]
        line = {"metric": "MA+Pk snapshot throughput", "value": main["value"], "unit": "particles/s", "n_gpus": world,
                "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
                "s_per_snapshot": ms_step * 1e-3, "higher_is_better": True, "scaling": "strong" if wl.get("strong") else "weak", "vs_baseline": None,
                "dtype": "f32", "data": ("synthetic uniform random particles" if args.data == "uniform" else
                                         "synthetic Zel'dovich-displaced lattice (rms 2 cells, lattice order)") + " generated on device, seed 1+rank",
                "config": workload_config(args, wl, world), "clocks": main["clocks"], "gpu_launches": main["launches"],
                "e2e": main["e2e"], "roofline": roof, "roofline_deposit": roof_dep, "fft": fft, "cpu_baseline": cpu,
                "stages": main.get("stages"), "kernels_ms_per_step": K,
                "grid_cells_per_s": nf * gside ** 3 / (ms_step * 1e-3),
                "check": {"P0_first_bins": [float(x) for x in (spec.Pk[:3, 0, 0] if nf > 1 else spec.Pk[:3, 0])],
                          "shot_noise_expected": BOX ** 3 / (npart * world),
                          "Nmodes_sum_ok": bool(spec.Nmodes3D.sum() + 1 == (gside ** 3 - 8) // 2 + 8) if gside % 2 == 0 else None,
                          "parity": parity},
                "extra": extra}
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


def parity_single(args, wl, dev, rank):
    """N = 1: the 256^3 sample of this workload through the reference's CPU code (timed: cpu_baseline) and through the CUDA
    path from the SAME host arrays; returns (check.parity, cpu_baseline)."""
    import numpy as np
    from pylians_b200 import MAS_library as MASL, Pk_library as PKL
    cn = min(256, wl["nside"])
    m = cpu_measure(cn, wl, 2)
    sec = min(m["times"])
    nf = len(wl["fields"])
    ref_grids, ref_spec = m["result"]
    grids, rel = [], []
    for (pos, W), (mas, weighted), rg in zip(m["parts"], wl["fields"], ref_grids):
        g = np.zeros((cn,) * 3, np.float32)
        MASL.MA(pos, g, BOX, mas, W=W if weighted else None)
        rel.append(grid_rel(g, rg))
        MASL.overdensity(g)
        grids.append(g)
    mine = PKL.Pk(grids[0], BOX, wl["axis"], wl["fields"][0][0], 1) if nf == 1 else PKL.XPk(grids, BOX, wl["axis"], mas_names(wl), 1)
    parity = {"against": m["kind"], "sample": "%d^3 particles per field, same host arrays fed to both arms" % cn,
              "grid_max_rel": max(rel), "tolerance": 1e-5}
    parity.update(spec_rel(mine, ref_spec, nf > 1))
    parity["pk_max_rel"] = parity["pk_max_rel_all"]
    # the contract of tests/parity.py: P(k) 1e-5; for a chain comparison (two different fp32 grids) of a
    # TSC/PCS field the bins BEYOND the Nyquist frequency get 1e-4 -- the deconvolution amplifies the summation-order noise
    # of the deposits there, and the reference's own serial and OpenMP deposits differ by 4e-6 in those bins at 128^3
    wide = any(mas in ("TSC", "PCS") for mas, _ in wl["fields"])
    parity["tolerance_beyond_nyquist"] = 1e-4 if wide else 1e-5
    parity["ok"] = bool(parity["nmodes_exact"] and parity["grid_max_rel"] <= 1e-5 and parity["pk_max_rel_below_nyquist"] <= 1e-5
                        and parity["pk_max_rel_all"] <= parity["tolerance_beyond_nyquist"])
    cpu = {"value": nf * cn ** 3 / sec, "unit": "particles/s", "cores": m["threads"], "kind": m["kind"], "deposit": m["deposit"],
           "sample": "%d^3 particles per field (%s) onto %d^3 grid + overdensity + %s, best of 2 (%.1f s each); deposit = the faster "
                     "on this host of the reference's serial loop and its OpenMP C kernel (see `deposit`), the Pk mode loop is "
                     "serial by construction" % (cn, "+".join(mas_names(wl)), cn, "XPk" if nf > 1 else "Pk", sec),
           "deposit_only": m["variants"]}
    return parity, cpu


def parity_multi(args, wl, dev, rank, world, dist):
    """N > 1: the NCCL slab pipeline (both exchange modes) against the single-GPU pipeline on ONE shared particle set
    (every rank draws the same 256^3 host particles, deposits its strided shard; each rank also runs the whole set alone)."""
    import numpy as np
    import torch
    from pylians_b200 import MAS_library as MASL, Pk_library as PKL, dist as pdist
    n, dims = 256, 256
    if dims % world:
        return {"skipped": "256 not divisible by %d ranks" % world}
    parts = host_particles(n, 1, seed=11)
    pos, W = parts[0]
    shard = torch.from_numpy(np.ascontiguousarray(pos[rank::world])).to(dev)
    wshard = torch.from_numpy(np.ascontiguousarray(W[rank::world])).to(dev)
    full = torch.from_numpy(pos).to(dev)
    wfull = torch.from_numpy(W).to(dev)
    cases = []
    worst = {"grid_max_rel": 0.0, "pk_max_rel_below_nyquist": 0.0, "pk_max_rel_all": 0.0, "nmodes_exact": True}
    for mas, weighted in (("CIC", False), ("PCS", False), ("TSC", True)):
        g = torch.zeros((dims,) * 3, device=dev)
        MASL.MA(full, g, BOX, mas, W=wfull if weighted else None)
        raw = g.clone()
        MASL.overdensity(g)
        single = PKL.Pk(g, BOX, wl["axis"], mas, 1)
        nxl = dims // world
        for mode in ("grid", "particles"):
            eng = pdist.SlabPk(dims, BOX, mas, wl["axis"], exchange=mode)
            slab = eng.density_slab(shard, wshard if weighted else None, mas, overdensity=False)
            grel = grid_rel(slab.cpu().numpy(), raw[rank * nxl:(rank + 1) * nxl].cpu().numpy())
            multi = eng.run(shard, wshard if weighted else None)
            r = spec_rel(multi, single, False)
            r["grid_max_rel"] = grel
            t = torch.tensor([r["grid_max_rel"], r["pk_max_rel_below_nyquist"], r["pk_max_rel_all"], 0.0 if r["nmodes_exact"] else 1.0],
                             device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            r["grid_max_rel"], r["pk_max_rel_below_nyquist"], r["pk_max_rel_all"] = float(t[0]), float(t[1]), float(t[2])
            r["nmodes_exact"] = bool(t[3].item() == 0.0)
            cases.append(dict(mas=mas + ("W" if weighted else ""), exchange=mode, **{k: r[k] for k in worst}))
            for k in ("grid_max_rel", "pk_max_rel_below_nyquist", "pk_max_rel_all"):
                worst[k] = max(worst[k], r[k])
            worst["nmodes_exact"] = worst["nmodes_exact"] and r["nmodes_exact"]
            del eng, slab
    out = {"against": "single-GPU pipeline of this library on every rank (itself pinned to the reference by check.parity at N=1 and tests/)",
           "sample": "%d^3 shared particles, rank r deposits rows r::%d; NCCL world %d" % (n, world, world),
           "cases": cases, "tolerance": "Nmodes exact; grid 1e-5; P(k) 1e-5 up to the Nyquist frequency (sharding changes the fp32 "
                                        "summation order of the grid; beyond k_N the TSC/PCS deconvolution amplifies that noise)"}
    out.update(worst)
    out["pk_max_rel"] = worst["pk_max_rel_below_nyquist"]
    out["ok"] = bool(worst["nmodes_exact"] and worst["grid_max_rel"] <= 1e-5 and worst["pk_max_rel_below_nyquist"] <= 1e-5)
    return out


_REAL_STDOUT = None


def _claim_stdout():
    """stdout must carry exactly ONE JSON line.  Libraries write to file descriptor 1 behind Python's back (NCCL prints its
    version banner there under NCCL_DEBUG=VERSION, whatever NCCL_DEBUG_FILE says; the compiled reference prints its timings),
    so fd 1 is pointed at stderr for the whole run and the JSON line goes to a private duplicate of the original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
        return
    while data:
        data = data[os.write(_REAL_STDOUT, data):]


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", dest="extras", action="store_false")
    ap.add_argument("--data", default="uniform", choices=["uniform", "zeldovich"])
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
