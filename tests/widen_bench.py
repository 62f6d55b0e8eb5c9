"""Timings of the widened rows on one B200 against the compiled, unmodified reference on the
box's host cores: snapshot driver (Pk_Gadget on a synthetic multi-file Gadget snapshot), field smoothing, the
bispectrum and the correlation function.  Writes one JSON line per case.

    python tests/widen_bench.py [--nside 256] [--out gpurun_out/widen_bench.jsonl]

It lives under tests/ because it runs the compiled reference (oracle/_ref) as the timed CPU baseline, and only tests/,
smoke() and bench.py's baseline legs may execute anything under oracle/.

GPU times: best of 3 wall-clock runs around the public call with a device synchronise on both sides (these calls
return host arrays or files, so wall clock is the user-visible time).  Reference: one run (it is slow), all host
threads offered (`threads=nproc`; its loops are serial).  The snapshot lives under $TMPDIR (page cache)."""
import argparse
import contextlib
import io
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))     # gadget_writer


def best_of(fn, n=3):
    import torch
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def once(fn):
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        fn()
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nside", type=int, default=256)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "widen_bench.jsonl"))
    ap.add_argument("--no-reference", action="store_true")
    a = ap.parse_args()
    import torch
    import pylians_b200
    import MAS_library as MASL, Pk_library as PKL, smoothing_library as SL
    import gadget_writer as GW
    from oracle import ref_loader
    pylians_b200.set_verbose(False)
    ref = None if a.no_reference or not ref_loader.extras_available() else ref_loader.load_extras()
    RPKL = ref_loader.load()[1] if ref else None
    threads = os.cpu_count() or 1
    n, box = a.nside, 1000.0
    lines = []

    def emit(case, gpu_s, ref_s, **kw):
        d = dict(case=case, gpu_s=gpu_s, reference_cpu_s=ref_s, speedup=(ref_s / gpu_s if ref_s else None),
                 host_threads=threads, **kw)
        lines.append(d); print(json.dumps(d), flush=True)

    # ---- snapshot driver: n^3 CDM particles with a header mass + n^3/8 gas particles with a MASS block, 4 files
    counts = [n ** 3 // 8, n ** 3, 0, 0, 0, 0]
    masstable = np.array([0.0, 0.65, 0, 0, 0, 0])
    parts = GW.make_particles(5, counts, box * 1e3, masstable)
    with tempfile.TemporaryDirectory() as tmp:
        base = os.path.join(tmp, "snap_000")
        GW.write_snapshot(base, parts, masstable, box * 1e3, 0.5, 4, 1)
        nbytes = sum(os.path.getsize("%s.%d" % (base, i)) for i in range(4))
        del parts
        for tag, types in (("Pk_Gadget [1] (real space)", [1]), ("Pk_Gadget [0,1] + RSD", [0, 1])):
            rsd = len(types) > 1
            out1 = os.path.join(tmp, "g"); os.makedirs(out1, exist_ok=True)
            g = best_of(lambda: PKL.Pk_Gadget(base, n, types, rsd, 2, threads, out1))
            r = None
            if ref:
                out2 = os.path.join(tmp, "r"); os.makedirs(out2, exist_ok=True)
                r = once(lambda: ref["Pk_snapshot"].Pk_Gadget(base, n, types, rsd, 2, threads, out2))
            emit(tag, g, r, particles=sum(counts[t] for t in types), grid=n, snapshot_bytes=nbytes, files=4)
        g = best_of(lambda: MASL.density_field_gadget(base, [1], n, "PCS", True, 0, False))
        r = once(lambda: ref["MAS_gadget"].density_field_gadget(base, [1], n, "PCS", True, 0, False)) if ref else None
        emit("density_field_gadget [1] PCS + RSD", g, r, particles=counts[1], grid=n)

    # ---- FFT consumers on an n^3 overdensity field
    gen = torch.Generator(device="cuda"); gen.manual_seed(2)
    pos = torch.rand((n ** 3, 3), device="cuda", dtype=torch.float32, generator=gen) * box
    delta = torch.zeros((n,) * 3, device="cuda", dtype=torch.float32)
    MASL.MA(pos, delta, box, "CIC"); MASL.overdensity(delta)
    d_host = delta.cpu().numpy()
    del pos
    W_k = SL.FT_filter(box, 20.0, n, "Gaussian", threads)
    g = best_of(lambda: SL.FT_filter(box, 20.0, n, "Gaussian", threads))
    r = once(lambda: ref["smoothing_library"].FT_filter(box, 20.0, n, "Gaussian", threads)) if ref else None
    emit("FT_filter Gaussian (numpy out)", g, r, grid=n)
    g = best_of(lambda: SL.field_smoothing(d_host, W_k, threads))
    r = once(lambda: ref["smoothing_library"].field_smoothing(d_host, W_k, threads)) if ref else None
    emit("field_smoothing (numpy in/out)", g, r, grid=n)
    W_dev = torch.from_numpy(W_k).cuda()
    g = best_of(lambda: SL.field_smoothing(delta, W_dev, threads))
    emit("field_smoothing (CUDA tensors in/out)", g, None, grid=n)
    g = best_of(lambda: PKL.Xi(d_host, box, "CIC", 2, threads))
    r = once(lambda: RPKL.Xi(d_host, box, "CIC", 2, threads)) if ref else None
    emit("Xi (numpy in)", g, r, grid=n)
    kF = 2 * np.pi / box
    theta = np.linspace(0.2, 3.0, 5)
    nb = min(n, 128)                       # the reference's Bk builds Python lists of cell IDs: keep it bounded
    db = np.ascontiguousarray(d_host[:nb, :nb, :nb])
    g = best_of(lambda: PKL.Bk(db, box, 6 * kF, 9 * kF, theta, "CIC", threads))
    r = once(lambda: ref["bispectrum_library"].Bk(db, box, 6 * kF, 9 * kF, theta, "CIC", threads)) if ref else None
    emit("Bk, 5 theta bins", g, r, grid=nb)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        for d in lines:
            f.write(json.dumps(d) + "\n")


if __name__ == "__main__":
    main()
