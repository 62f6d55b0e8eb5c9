"""The fused binning call on ONE rank's k-space window of the slab pipeline, on one GPU:
    python profiles/ring_slab.py [N=2048] [G=8] [rank=0]
k-space window (N kx, N/G ky, P kz) in the transposed layout with the even row pitch P = N/2+2, random data; CUDA events
around pylb_pk_bin (whole call) and the library's own bracket around the ring kernel.  Algorithmic bytes = 8 N (N/G)(N/2+1)."""
import sys, torch
sys.path.insert(0, '.')
import pylians_b200
from pylians_b200 import _lib, Pk_library as PKL
pylians_b200.set_verbose(False)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
G = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device('cuda', 0)
nyl, P = N // G, N // 2 + 2
gen = torch.Generator(device=dev); gen.manual_seed(3)
win = torch.view_as_complex(torch.randn((N, nyl, P, 2), device=dev, generator=gen))
ks = _lib.KSpace(N, 0, N, rank * nyl, nyl, nyl * P, P)
mi = [PKL.MAS_function('PCS')]
peak = 6532.5
for want_phase in (1, 0):
    for _ in range(3):
        PKL.bin_modes([win], N, 2, mi, want_phase, False, ks=ks)
    torch.cuda.synchronize()
    _lib.timing_enable(True)
    for w in (_lib.T_BIN, _lib.T_RING):
        _lib.timing_collect(w)
    n = 10
    for _ in range(n):
        PKL.bin_modes([win], N, 2, mi, want_phase, False, ks=ks)
    torch.cuda.synchronize()
    tb, _ = _lib.timing_collect(_lib.T_BIN); tr, _ = _lib.timing_collect(_lib.T_RING)
    _lib.timing_enable(False)
    by = 8.0 * N * nyl * (N // 2 + 1)
    print("N=%d G=%d rank=%d phase=%d: whole call %.4f ms = %.0f GB/s = %.3f of peak; ring kernel %.4f ms = %.3f; rest %.1f us" % (
        N, G, rank, want_phase, tb / n, by / (tb / n) / 1e6, by / (tb / n) / 1e6 / peak, tr / n, by / (tr / n) / 1e6 / peak,
        (tb - tr) / n * 1e3), flush=True)
