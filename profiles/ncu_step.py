"""A few snapshots of one workload for ncu: python profiles/ncu_step.py N MAS [steps]  (MA -> overdensity -> Pk, device-resident)"""
import sys, torch
sys.path.insert(0, '.')
import pylians_b200
from pylians_b200 import MAS_library as MASL, Pk_library as PKL
pylians_b200.set_verbose(False)
N, mas = int(sys.argv[1]), sys.argv[2]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device('cuda', 0)
gen = torch.Generator(device=dev); gen.manual_seed(1)
pos = torch.rand((N ** 3, 3), device=dev, generator=gen) * 1000.0
grid = torch.empty((N,) * 3, device=dev)
for _ in range(steps):
    grid.zero_(); MASL.MA(pos, grid, 1000.0, mas); MASL.overdensity(grid); PKL.Pk(grid, 1000.0, 2, mas, 1)
torch.cuda.synchronize()
