mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_drivers.py -q -m gpu --tb=short -k "xpk or XPk or Gadget or gadget or comp" 2>&1 | tail -12
python - <<'PY'
import sys, time, torch
sys.path.insert(0, '.')
import pylians_b200
from pylians_b200 import Pk_library as PKL
import pylians_b200.Pk_library as P
pylians_b200.set_verbose(False)
fs = [torch.randn((256,) * 3, device='cuda') for _ in range(4)]
for algo, label in ((0, "three fields at a time (ring2x)"), (1, "one thread per mode")):
    P.ALGO = algo
    for _ in range(2): PKL.XPk(fs, 1000.0, 2, ['CIC'] * 4, 1)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(3): PKL.XPk(fs, 1000.0, 2, ['CIC'] * 4, 1)
    torch.cuda.synchronize(); print("XPk(4 fields, 256^3), %s: %.2f ms per call" % (label, (time.perf_counter() - t) / 3 * 1e3))
PY
