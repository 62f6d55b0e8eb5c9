import time, torch, numpy as np, sys
sys.path.insert(0,'.')
import pylians_b200
from pylians_b200 import _lib, MAS_library as MASL, Pk_library as PKL
pylians_b200.set_verbose(False)
dev=torch.device('cuda',0); N=512; box=1000.0
gen=torch.Generator(device=dev); gen.manual_seed(1)
pos=torch.rand((N**3,3),device=dev,generator=gen)*box
grid=torch.zeros((N,)*3,device=dev)
def T(f,n=5):
    f(); torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter()-t)/n*1e3
print("MA        %.3f ms"%T(lambda: MASL.MA(pos,grid,box,'CIC')))
print("overdens  %.3f ms"%T(lambda: MASL.overdensity(grid)))
grid.zero_(); MASL.MA(pos,grid,box,'CIC'); MASL.overdensity(grid)
print("Pk        %.3f ms"%T(lambda: PKL.Pk(grid,box,2,'CIC',1)))
lib=_lib.load(); st=torch.cuda.current_stream(dev)
print("fft       %.3f ms"%T(lambda: PKL._fft_field(lib,grid,N,dev,st)))
dk=PKL._fft_field(lib,grid,N,dev,st)
print("bin_modes %.3f ms"%T(lambda: PKL.bin_modes([dk],N,2,[2],True,False)))
L,s,c=PKL.bin_modes([dk],N,2,[2],True,False)
print("_Bins     %.3f ms"%T(lambda: PKL._Bins(L,s,c)))
b=PKL._Bins(L,s,c)
class O: pass
print("_finish   %.3f ms"%T(lambda: PKL._finish(O(),b,N,box,False)))
_lib.timing_enable(True)
import cProfile, pstats
pr=cProfile.Profile(); pr.enable()
for _ in range(5): PKL.Pk(grid,box,2,'CIC',1)
pr.disable(); pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
