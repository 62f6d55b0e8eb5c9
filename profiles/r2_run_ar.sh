mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short -k "keep_deltak or pk_golden or pk_vs_oracle" 2>&1 | tail -8
python - <<'PY'
import sys, time, torch
sys.path.insert(0, '.')
import pylians_b200
from pylians_b200 import Pk_library as PKL
pylians_b200.set_verbose(False)
g = torch.randn((512,) * 3, device='cuda')
for axis in (2, 0):
    for _ in range(2): PKL.Pk(g, 1000.0, axis, 'CIC', 1, keep_deltak=True)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(3): PKL.Pk(g, 1000.0, axis, 'CIC', 1, keep_deltak=True)
    torch.cuda.synchronize(); print("Pk(512^3, axis=%d, keep_deltak=True): %.2f ms per call" % (axis, (time.perf_counter() - t) / 3 * 1e3))
PY
