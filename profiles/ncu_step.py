"""Two device-resident MA -> overdensity -> Pk steps of the bench workload, for
   ncu --set full --import-source on --clock-control none -k regex:'ring2|special2|bin_hist|bin_pass|deposit_tile' --launch-skip 7 -c 7"""
import sys, torch
sys.path.insert(0, '.')
import pylians_b200
from pylians_b200 import MAS_library as MASL, Pk_library as PKL
pylians_b200.set_verbose(False)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
mas = sys.argv[2] if len(sys.argv) > 2 else "CIC"
dev = torch.device('cuda', 0)
gen = torch.Generator(device=dev); gen.manual_seed(1)
pos = torch.rand((N ** 3, 3), device=dev, dtype=torch.float32, generator=gen) * 1000.0
grid = torch.empty((N,) * 3, device=dev, dtype=torch.float32)
for _ in range(2):
    grid.zero_(); MASL.MA(pos, grid, 1000.0, mas); MASL.overdensity(grid); PKL.Pk(grid, 1000.0, 2, mas, 1)
torch.cuda.synchronize()
