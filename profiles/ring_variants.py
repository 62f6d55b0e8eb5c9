"""Time the binning ring kernel alone (CUDA events around the kernel inside the library) for a few flag
combinations on a random 512^3 k-space field."""
import sys, torch
sys.path.insert(0, '.')
import pylians_b200
from pylians_b200 import _lib, Pk_library as PKL
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev); g.manual_seed(0)
dk = torch.view_as_complex(torch.randn((N, N, N // 2 + 1, 2), device=dev, generator=g))
dk2 = torch.view_as_complex(torch.randn((N, N, N // 2 + 1, 2), device=dev, generator=g))
nz = N // 2 + 1; nzp = nz + (nz & 1)
dkp = torch.view_as_complex(torch.randn((N, N, nzp, 2), device=dev, generator=g))[:, :, :nz]   # even row pitch: what Pk() feeds the kernel
_lib.timing_enable(True)
def run(fields, phase, algo, label, wb=False):
    for _ in range(3): PKL.bin_modes(fields, N, 2, [2] * len(fields), phase, wb, algo=algo)
    torch.cuda.synchronize(); _lib.timing_collect(_lib.T_RING)
    for _ in range(10): PKL.bin_modes(fields, N, 2, [2] * len(fields), phase, wb, algo=algo)
    ms, n = _lib.timing_collect(_lib.T_RING)
    ms /= n
    gb = 8.0 * N * N * (N // 2 - 1) * len(fields) / 1e9
    print("%-44s %.4f ms  %.0f GB/s  %.1f%% of 6553.6" % (label, ms, gb / ms * 1e3, gb / ms * 1e3 / 65.536))
run([dkp], True, 2, "F=1 phase ring2, even row pitch (Pk default)")
run([dkp], False, 2, "F=1 nophase ring2, even row pitch")
run([dk], True, 2, "F=1 phase ring2, dense rows (parity tables)")
run([dk], False, 2, "F=1 nophase ring2, dense rows")
run([dk], True, 2 | 64 | 32, "F=1 phase ring1 bulk")
run([dk], True, 2 | 64, "F=1 phase ring1 cp.async")
run([dk], False, 2 | 64 | 32, "F=1 nophase ring1 bulk")
run([dk], False, 2 | 64, "F=1 nophase ring1 cp.async")
run([dk], True, 2 | 16 | 32, "F=1 phase fp64-option bulk")
run([dk], True, 2 | 16, "F=1 phase fp64-option cp.async")
run([dk, dk2], False, 2 | 32, "F=2 (XPk) bulk")
run([dk, dk2], False, 2, "F=2 (XPk) cp.async (default)")
