"""A few Pk calls on a random field for ncu: python profiles/ncu_pk.py N [calls]  (launch list of pylb_pk_bin's kernels)"""
import sys, torch
sys.path.insert(0, '.')
import pylians_b200
from pylians_b200 import Pk_library as PKL
pylians_b200.set_verbose(False)
N = int(sys.argv[1]); calls = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device('cuda', 0)
gen = torch.Generator(device=dev); gen.manual_seed(1)
grid = torch.randn((N,) * 3, device=dev, generator=gen)
for _ in range(calls):
    PKL.Pk(grid, 1000.0, 2, 'PCS', 1)
torch.cuda.synchronize()
