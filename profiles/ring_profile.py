"""One binning call per variant on a random k-space field, for `ncu -k regex:ring` captures."""
import sys, torch
sys.path.insert(0, '.')
import pylians_b200
from pylians_b200 import _lib, Pk_library as PKL
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
algos = [int(a) for a in sys.argv[2].split(',')] if len(sys.argv) > 2 else [2]
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev); g.manual_seed(0)
dk = torch.view_as_complex(torch.randn((N, N, N // 2 + 1, 2), device=dev, generator=g))
for a in algos:
    for _ in range(2):
        PKL.bin_modes([dk], N, 2, [2], True, False, algo=a)
torch.cuda.synchronize()
