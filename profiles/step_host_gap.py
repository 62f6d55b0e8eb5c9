"""Where the time between two snapshots goes (host side): python profiles/step_host_gap.py [N=1024] [MAS=PCS]
Per step: host seconds spent inside grid.zero_ + MASL.MA (asynchronous launches), MASL.overdensity, PKL.Pk (blocks on
the read-back of the bins), and the CUDA-event time of the whole step."""
import sys, time, torch
sys.path.insert(0, '.')
import pylians_b200
from pylians_b200 import MAS_library as MASL, Pk_library as PKL
pylians_b200.set_verbose(False)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
mas = sys.argv[2] if len(sys.argv) > 2 else 'PCS'
dev = torch.device('cuda', 0)
gen = torch.Generator(device=dev); gen.manual_seed(1)
pos = torch.rand((N ** 3, 3), device=dev, generator=gen) * 1000.0
grid = torch.empty((N,) * 3, device=dev)
def step(rec):
    t0 = time.perf_counter()
    grid.zero_(); MASL.MA(pos, grid, 1000.0, mas)
    t1 = time.perf_counter()
    MASL.overdensity(grid)
    t2 = time.perf_counter()
    r = PKL.Pk(grid, 1000.0, 2, mas, 1)
    t3 = time.perf_counter()
    rec.append((t1 - t0, t2 - t1, t3 - t2))
    return r
for _ in range(3):
    step([])
torch.cuda.synchronize()
rec = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 6
w0 = time.perf_counter(); e0.record()
for _ in range(K):
    step(rec)
e1.record(); torch.cuda.synchronize(); w1 = time.perf_counter()
print("N=%d %s: %.2f ms per step by CUDA events, %.2f ms by wall clock" % (N, mas, e0.elapsed_time(e1) / K, (w1 - w0) / K * 1e3))
for a, b, c in rec:
    print("  host ms: zero+MA %.3f  overdensity %.3f  Pk (blocking) %.3f" % (a * 1e3, b * 1e3, c * 1e3))
# the same launches with ONE synchronisation at the end (no read-back between steps): the GPU-only time of K steps
e0.record()
for _ in range(K):
    grid.zero_(); MASL.MA(pos, grid, 1000.0, mas); MASL.overdensity(grid)
e1.record(); torch.cuda.synchronize()
print("  zero + deposit + overdensity queued back to back, no read-back: %.2f ms per step" % (e0.elapsed_time(e1) / K))

# the bench's clock sampling: two NVML queries per step, issued right after the deposit has been queued
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    for label, fn in (("nvmlDeviceGetClockInfo + ThrottleReasons", lambda: (pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))),
                      ("nvmlDeviceGetClockInfo only", lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))):
        torch.cuda.synchronize()
        q = []
        e0.record()
        for _ in range(K):
            grid.zero_(); MASL.MA(pos, grid, 1000.0, mas)
            t = time.perf_counter(); fn(); q.append((time.perf_counter() - t) * 1e3)
            MASL.overdensity(grid)
            PKL.Pk(grid, 1000.0, 2, mas, 1)
        e1.record(); torch.cuda.synchronize()
        print("  with %s after the deposit launch: %.2f ms per step; the queries took %s ms" % (label, e0.elapsed_time(e1) / K, ["%.2f" % x for x in q]))
except Exception as e:  # noqa: BLE001
    print("  nvml:", repr(e))
