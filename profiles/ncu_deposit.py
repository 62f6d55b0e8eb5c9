"""One tiled deposit for ncu: python profiles/ncu_deposit.py N MAS kernel   (kernel: 1 float, 2 fixed point)"""
import sys, torch
sys.path.insert(0, '.')
import pylians_b200
from pylians_b200 import _lib, MAS_library as MASL
N, mas, k = int(sys.argv[1]), sys.argv[2], int(sys.argv[3])
dev = torch.device('cuda', 0)
gen = torch.Generator(device=dev); gen.manual_seed(1)
pos = torch.rand((N ** 3, 3), device=dev, generator=gen) * 1000.0
grid = torch.zeros((N,) * 3, device=dev)
_lib.load().pylb_ma_debug_path(100 * k if k else -1)
for _ in range(2):
    MASL.MA(pos, grid, 1000.0, mas)
torch.cuda.synchronize()
