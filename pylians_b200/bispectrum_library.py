"""Bispectrum on the GPU: mirror of library/Pk_library/bispectrum_library.pyx (`class Bk` :32-196, `F2` :200-204,
`Bispectrum_theory` :208-241) -- a consumer of PKL.FFT3Dr_f / IFFT3Dr_f.

    BBk = PKL.Bk(delta, BoxSize, k1, k2, theta, MAS='CIC', threads=1);  BBk.B, BBk.Q, BBk.k, BBk.Pk

Built for the GPU: instead of the reference's Python lists of cell IDs per shell, one fused kernel per shell reads
delta_k once, applies the fp32 MAS factor and writes the shell-filtered field and the shell indicator; both go
through cuFFT's C2R, and the sums of delta_i^2, I_i^2, delta1*delta2*delta3 and I1*I2*I3 are streaming double
reductions.  Everything stays in HBM; four doubles per theta bin come back."""
import time

import numpy as np
import torch

from . import _lib
from .MAS_library import _device
from . import Pk_library as PKL
from .Pk_library import _check_field, _fft_field, _say, _stream


class Bk(object):
    """bispectrum_library.pyx:32-196.  Attributes B, Q (one per theta), k = [k1, k2, k3...], Pk at those k."""

    def __init__(self, delta, BoxSize, k1, k2, theta, MAS="CIC", threads=1):
        start = time.time()
        _say("\nComputing bispectrum of the field...")
        lib, dev = _lib.load(), _device()
        delta = _check_field(delta)
        dims = len(delta)
        kF, kN, kmax_par, kmax_per, kmax = PKL.frequencies(BoxSize, dims)
        MAS_index = PKL.MAS_function(MAS)
        theta = np.asarray(theta, dtype=np.float64)
        bins = theta.shape[0]
        k3 = np.sqrt((k2 * np.sin(theta)) ** 2 + (k2 * np.cos(theta) + k1) ** 2)          # :62
        k_all = np.zeros(bins + 2, dtype=np.float64)
        k_all[0], k_all[1], k_all[2:] = k1, k2, k3
        k_min, k_max = (k_all - kF) / kF, (k_all + kF) / kF                                # :69-74
        st = _stream(dev)
        delta_k = _fft_field(lib, delta, dims, dev, torch.cuda.current_stream(dev))
        n = dims ** 3
        nk = (dims, dims, dims // 2 + 1)
        sums = torch.zeros(4, dtype=torch.float64, device=dev)

        def shell(i):
            """(delta_i, I_i) in real space for shell i, and P(k_i) (:132-147)."""
            dk_i = torch.empty(nk, dtype=torch.complex64, device=dev)
            ik_i = torch.empty(nk, dtype=torch.complex64, device=dev)
            _lib.check(lib.pylb_bk_shell(delta_k.data_ptr(), dk_i.data_ptr(), ik_i.data_ptr(), dims, MAS_index,
                                         float(k_min[i]), float(k_max[i]), st), "pylb_bk_shell")
            d = torch.empty((dims, dims, dims), dtype=torch.float32, device=dev)
            ind = torch.empty((dims, dims, dims), dtype=torch.float32, device=dev)
            _lib.check(lib.pylb_fft_c2r(dk_i.data_ptr(), d.data_ptr(), dims, 1, st), "pylb_fft_c2r")
            _lib.check(lib.pylb_fft_c2r(ik_i.data_ptr(), ind.data_ptr(), dims, 1, st), "pylb_fft_c2r")
            _lib.check(lib.pylb_prod_sum(d.data_ptr(), d.data_ptr(), None, n, sums[0:1].data_ptr(), st), "pylb_prod_sum")
            _lib.check(lib.pylb_prod_sum(ind.data_ptr(), ind.data_ptr(), None, n, sums[1:2].data_ptr(), st), "pylb_prod_sum")
            s = sums[:2].cpu().numpy()
            return d, ind, (s[0] / s[1]) * (BoxSize / dims ** 2) ** 3

        Pk = np.zeros(bins + 2, dtype=np.float64)
        B = np.zeros(bins, dtype=np.float64)
        Q = np.zeros(bins, dtype=np.float64)
        delta1, I1, Pk[0] = shell(0)
        delta2, I2, Pk[1] = shell(1)
        for j in range(bins):
            delta3, I3, Pk[j + 2] = shell(j + 2)
            _lib.check(lib.pylb_prod_sum(delta1.data_ptr(), delta2.data_ptr(), delta3.data_ptr(), n,
                                         sums[2:3].data_ptr(), st), "pylb_prod_sum")
            _lib.check(lib.pylb_prod_sum(I1.data_ptr(), I2.data_ptr(), I3.data_ptr(), n, sums[3:4].data_ptr(), st),
                       "pylb_prod_sum")
            s = sums[2:].cpu().numpy()
            B[j] = (s[0] / s[1]) * (BoxSize ** 2 / dims ** 3) ** 3                          # :194
            Q[j] = B[j] / (Pk[0] * Pk[1] + Pk[0] * Pk[j + 2] + Pk[1] * Pk[j + 2])
        self.B, self.Q, self.k, self.Pk = B, Q, k_all, Pk
        _say("Time to compute bispectrum = %.2f" % (time.time() - start))


def F2(k1_vec, k2_vec):
    """Second-order perturbation-theory kernel, bispectrum_library.pyx:200-204."""
    k1_mod = np.sqrt(np.dot(k1_vec, k1_vec))
    k2_mod = np.sqrt(np.dot(k2_vec, k2_vec))
    ctheta = np.dot(k1_vec, k2_vec) / (k1_mod * k2_mod)
    return 5.0 / 7.0 + 1.0 / 2.0 * ctheta * (k1_mod / k2_mod + k2_mod / k1_mod) + 2.0 / 7.0 * ctheta ** 2


def Bispectrum_theory(k, Pk, k1, k2):
    """Tree-level bispectrum on 50 angles given the linear P(k), bispectrum_library.pyx:208-241 (host arithmetic on
    50 numbers; the reference's per-angle debug prints are dropped)."""
    bins = 50
    B = np.zeros(bins, dtype=np.float64)
    thetas = np.linspace(0, np.pi, bins)
    k1_vec = np.array([0, 0, k1])
    Pk1 = np.interp(np.log(k1), np.log(k), Pk)
    Pk2 = np.interp(np.log(k2), np.log(k), Pk)
    for i, theta in enumerate(thetas):
        k2_vec = np.array([0, k2 * np.sin(theta), k2 * np.cos(theta)])
        k3_vec = np.array([0, -k2 * np.sin(theta), -k2 * np.cos(theta) - k1])
        Pk3 = np.interp(np.log(np.sqrt(np.dot(k3_vec, k3_vec))), np.log(k), Pk)
        B[i] = (2.0 * Pk1 * Pk2 * F2(k1_vec, k2_vec) + 2.0 * Pk1 * Pk3 * F2(k1_vec, k3_vec)
                + 2.0 * Pk2 * Pk3 * F2(k2_vec, k3_vec))
    return thetas, B
