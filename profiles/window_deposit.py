"""Single-GPU timing of the two local kernels of the particle-exchange mode at the 8-GPU weak-scaling shape
(N = 1024, G = 8, 512^3 particles per rank): pylb_partition_xslab on this rank's shard, and pylb_ma_window on the
packed (x,y,z,w) records a rank receives (uniform inside its slab), plus a parity check of the window against
MASL.MA on the same particles."""
import sys, torch
sys.path.insert(0, '.')
import pylians_b200
from pylians_b200 import MAS_library as MASL
from pylians_b200.dist import CudaOps
pylians_b200.set_verbose(False)
N, G, box = 1024, 8, 1000.0
dev = torch.device('cuda', 0)
gen = torch.Generator(device=dev); gen.manual_seed(4)
ops = CudaOps()
nxl = N // G


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for mas, halo in (("CIC", 1), ("PCS", 3)):
    pos = torch.rand((512 ** 3, 3), device=dev, generator=gen) * box
    t_part = timed(lambda: ops.partition(pos, None, box, mas, G, N))
    # what rank 0 receives: particles whose lowest touched plane lies in its slab -> uniform in x over one slab width
    rec = torch.empty((512 ** 3, 4), device=dev)
    rec[:, :3] = pos
    lo = {"CIC": 0.0, "PCS": 1.0}[mas] * box / N          # PCS: base cell = floor(x) - 1
    rec[:, 0] = pos[:, 0] / G + lo
    rec[:, 3] = 1.0
    del pos
    grid = torch.zeros((nxl + halo, N, N), device=dev)
    t_win = timed(lambda: ops.deposit_window(rec, grid, 0, box, mas, False, N))
    grid.zero_(); ops.deposit_window(rec, grid, 0, box, mas, False, N)
    full = torch.zeros((N, N, N), device=dev)
    MASL.MA(rec[:, :3], full, box, mas)
    err = (full[:nxl + halo] - grid).abs().max().item()
    rest = full[nxl + halo:].abs().max().item()
    print("%s: partition %.2f ms   window deposit %.2f ms   max |window - full deposit| %.2e (max cell %.1f), outside window %.1e"
          % (mas, t_part, t_win, err, full.max().item(), rest))
    del rec, grid, full
