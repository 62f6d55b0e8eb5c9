import time, torch, sys
sys.path.insert(0,'.')
import pylians_b200
from pylians_b200 import _lib, MAS_library as MASL
dev=torch.device('cuda',0); N=512; box=1000.0
gen=torch.Generator(device=dev); gen.manual_seed(1)
pos=torch.rand((N**3,3),device=dev,generator=gen)*box
grid=torch.zeros((N,)*3,device=dev)
lib=_lib.load()
def T(f,n=5):
    f(); torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter()-t)/n*1e3
for mas in ('CIC','TSC','PCS'):
    for path,name in ((0,'2-pass sort 16x16x32'),(10,'1-pass scatter 16x16x32'),(1,'2-pass sort 32^3'),(2,'radix+gather')):
        lib.pylb_ma_debug_path(path); _lib.timing_enable(True); _lib.timing_collect(1)
        t=T(lambda: MASL.MA(pos,grid,box,mas)); ms,n=_lib.timing_collect(1)
        print("%s %-18s MA %.3f ms  tile kernel %.3f ms"%(mas,name,t,ms/max(n,1)))
lib.pylb_ma_debug_path(-1)
