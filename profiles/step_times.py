"""Per-step wall time (with a device sync per step) of the bench's device-resident loop: looks for sporadic stalls."""
import sys, time, torch
sys.path.insert(0, '.')
import pylians_b200
from pylians_b200 import MAS_library as MASL, Pk_library as PKL
pylians_b200.set_verbose(False)
N = 512; dev = torch.device('cuda', 0)
gen = torch.Generator(device=dev); gen.manual_seed(1)
pos = torch.rand((N ** 3, 3), device=dev, dtype=torch.float32, generator=gen) * 1000.0
grid = torch.empty((N,) * 3, device=dev, dtype=torch.float32)
ts = []
for it in range(30):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    grid.zero_(); MASL.MA(pos, grid, 1000.0, "CIC"); t1 = time.perf_counter()
    MASL.overdensity(grid); t2 = time.perf_counter()
    PKL.Pk(grid, 1000.0, 2, "CIC", 1); torch.cuda.synchronize(); t3 = time.perf_counter()
    ts.append(((t3 - t0) * 1e3, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
for i, t in enumerate(ts):
    print("step %2d: total %8.3f ms   MA host %7.3f  overdensity host %7.3f  Pk (host+sync) %8.3f" % ((i,) + t))
