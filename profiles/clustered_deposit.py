"""Deposit time on a clustered particle set (30 % of 512^3 particles in 4096 halo cells, sigma = 0.35 cells) against
the uniform set, with the tile kernel's warp-aggregated branch on or off (PYLB_MA_AGG, read once per process)."""
import os, sys, torch
sys.path.insert(0, '.')
import pylians_b200
from pylians_b200 import MAS_library as MASL, _lib
pylians_b200.set_verbose(False)
N, box = 512, 1000.0
dev = torch.device('cuda', 0)
gen = torch.Generator(device=dev); gen.manual_seed(6)
n = N ** 3
uni = torch.rand((n, 3), device=dev, generator=gen) * box
nh = int(0.3 * n)
centres = torch.rand((4096, 3), device=dev, generator=gen) * box
idx = torch.randint(0, 4096, (nh,), device=dev, generator=gen)
clu = uni.clone()
clu[:nh] = torch.remainder(centres[idx] + torch.randn((nh, 3), device=dev, generator=gen) * (0.35 * box / N), box)
clu = clu[torch.randperm(n, device=dev, generator=gen)]
grid = torch.zeros((N,) * 3, device=dev)
for name, pos in (("uniform", uni), ("clustered", clu)):
    for mas in ("CIC", "PCS"):
        for _ in range(2):
            grid.zero_(); MASL.MA(pos, grid, box, mas)
        _lib.timing_enable(True); _lib.timing_collect(_lib.T_TILE)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(3):
            grid.zero_(); MASL.MA(pos, grid, box, mas)
        e1.record(); torch.cuda.synchronize()
        t, k = _lib.timing_collect(_lib.T_TILE); _lib.timing_enable(False)
        print("PYLB_MA_AGG=%s %-9s %s: deposit %.2f ms  tile kernel %.2f ms  max cell %.0f" % (
            os.environ.get("PYLB_MA_AGG", "1"), name, mas, e0.elapsed_time(e1) / 3, t / max(k, 1), grid.max().item()))
