"""Why 1/rho_mean is NOT folded into the binning kernel (SURVEY 8f #1 / VERDICT "fold the normalisation away"): CPU experiment.
fp32 FFT of delta = rho/mean - 1 (what the path does) against fp32 FFT of rho, scaled by 1/mean in k-space with the DC mode
zeroed; both compared with a double-precision transform of delta.  python profiles/normalisation_fold.py [N=256]"""
import sys
import numpy as np
import scipy.fft as sf
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rng = np.random.default_rng(0)
mod2 = lambda dk: dk.real.astype(np.float64) ** 2 + dk.imag.astype(np.float64) ** 2
for label, rho in (("shot noise, 1 particle per cell", rng.poisson(1.0, (N, N, N)).astype(np.float32)),
                   ("64 particles per cell (contrast 0.125)", rng.poisson(64.0, (N, N, N)).astype(np.float32)),
                   ("smooth field, contrast 1e-2", (1000.0 * (1 + 1e-2 * rng.standard_normal((N, N, N)))).astype(np.float32))):
    mean = np.mean(rho, dtype=np.float64)
    d = (rho / np.float32(mean) - np.float32(1.0)).astype(np.float32)
    ref = mod2(sf.rfftn(d.astype(np.float64)))
    a = mod2(sf.rfftn(d))
    b = sf.rfftn(rho) / np.float32(mean); b[0, 0, 0] = 0; b = mod2(b)
    m = ref > 0
    ea, eb = np.abs(a - ref)[m] / ref.mean(), np.abs(b - ref)[m] / ref.mean()
    print("%-42s mean |dP|/<P>: delta-then-FFT %.2e, FFT-then-scale %.2e ; worst mode: %.2e vs %.2e" % (label, ea.mean(), eb.mean(), ea.max(), eb.max()))
