"""Why is bench.py's timed loop slower than the sum of its stages?  Toggle the NVML sampler thread and the
per-kernel event brackets."""
import sys, time, torch
sys.path.insert(0, '.')
import bench, pylians_b200
from pylians_b200 import _lib, MAS_library as MASL, Pk_library as PKL
pylians_b200.set_verbose(False)
dev = torch.device('cuda', 0); N = 512; box = 1000.0
gen = torch.Generator(device=dev); gen.manual_seed(1)
pos = torch.rand((N ** 3, 3), device=dev, generator=gen) * box
grid = torch.empty((N,) * 3, device=dev)
def snap():
    grid.zero_(); MASL.MA(pos, grid, box, 'CIC'); MASL.overdensity(grid); return PKL.Pk(grid, box, 2, 'CIC', 1)
def loop(k=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t = time.perf_counter(); e0.record()
    for _ in range(k): snap()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k, (time.perf_counter() - t) / k * 1e3
for _ in range(3): snap()
print("plain                      ev %.3f ms  wall %.3f ms" % loop())
_lib.timing_enable(True); print("event brackets             ev %.3f ms  wall %.3f ms" % loop()); _lib.timing_enable(False)
for period in (0.02, 0.1):
    s = bench.ClockSampler(0, period); s.start(); r = loop(); c = s.stop()
    print("nvml sampler %.2fs          ev %.3f ms  wall %.3f ms  samples %s" % ((period,) + r + (c.get("samples"),)))
print("plain again                ev %.3f ms  wall %.3f ms" % loop())

# which NVML query is expensive?
import pynvml, threading
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
def timed_call(f, n=20):
    t = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t) / n * 1e3
print("idle  clock query %.3f ms, reasons query %.3f ms" % (timed_call(lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
      timed_call(lambda: pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))))
for name, fn in (("clock only", lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                 ("reasons only", lambda: pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))):
    stop = threading.Event(); cnt = [0]
    def bg():
        while not stop.is_set():
            fn(); cnt[0] += 1; stop.wait(0.02)
    th = threading.Thread(target=bg, daemon=True); th.start(); r = loop(); stop.set(); th.join()
    print("%-14s every 20 ms: ev %.3f ms wall %.3f ms (%d calls)" % ((name,) + r + (cnt[0],)))
