"""A/B of the tile kernels of the tiled deposit (round 2): lane per particle (k=1) against stencil lanes with the
optimistic CAS pair (k=2, TSC and PCS).  CUDA events around MASL.MA and the library's own brackets around the sort stage
and the tile kernel.

    python profiles/deposit_ab.py [sizes...]      # default 512 1024
"""
import sys, time, torch
sys.path.insert(0, '.')
import pylians_b200
from pylians_b200 import _lib, MAS_library as MASL
dev = torch.device('cuda', 0); box = 1000.0
lib = _lib.load()
sizes = [int(a) for a in sys.argv[1:]] or [512, 1024]


def run(pos, grid, mas, W, n=3):
    MASL.MA(pos, grid, box, mas, W=W); torch.cuda.synchronize()
    _lib.timing_enable(True)
    for w in (_lib.T_TILE, _lib.T_SORT):
        _lib.timing_collect(w)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        MASL.MA(pos, grid, box, mas, W=W)
    e1.record(); torch.cuda.synchronize()
    tile, _ = _lib.timing_collect(_lib.T_TILE); sort, _ = _lib.timing_collect(_lib.T_SORT)
    _lib.timing_enable(False)
    return e0.elapsed_time(e1) / n, sort / n, tile / n


for N in sizes:
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    pos = torch.rand((N ** 3, 3), device=dev, generator=gen) * box
    W = torch.rand(N ** 3, device=dev, generator=gen) + 0.5
    grid = torch.zeros((N,) * 3, device=dev)
    for mas in ('NGP', 'CIC', 'TSC', 'PCS'):
        for k in ((1, 2) if mas in ('PCS', 'TSC') else (1,)):
            for w in ((None, W) if mas in ('CIC', 'PCS') else (None,)):
                lib.pylb_ma_debug_path(100 * k)
                ma, sort, tile = run(pos, grid, mas, w)
                print("N=%d %s%s kernel=%d  MA %.3f ms  sort %.3f ms  tile %.3f ms  (%.3g particles/s, %.3g updates/s in the tile kernel)" % (
                    N, mas, 'W' if w is not None else '', k, ma, sort, tile, N ** 3 / ma * 1e3,
                    N ** 3 * {'NGP': 1, 'CIC': 8, 'TSC': 27, 'PCS': 64}[mas] / tile * 1e3), flush=True)
    lib.pylb_ma_debug_path(-1)
    del pos, W, grid
    torch.cuda.empty_cache()
