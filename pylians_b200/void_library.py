"""`void_library.gaussian_smoothing` on the GPU (library/void_library/void_library.pyx:45-80): the k-space top-hat
smoothing the void finder applies to the density field -- a consumer of PKL.FFT3Dr_f / IFFT3Dr_f.
The void finder itself (void_library.pyx:89-...) is outside the MA -> Pk path and is not provided."""
import numpy as np
import torch

from . import _lib
from .MAS_library import _device
from .Pk_library import _check_field, _fft_field, _like_input, _stream


def gaussian_smoothing(delta, BoxSize, R, threads=1):
    """Top-hat (sic) smoothing of radius R: IFFT(FFT(delta) * 3(sin kR - kR cos kR)/(kR)^3).  float32 in and out."""
    lib, dev = _lib.load(), _device()
    delta = _check_field(delta)
    dims = delta.shape[0]
    # `float BoxSize, float R`; prefact = R*2.0*PI/BoxSize evaluated in double, stored in a float (:61)
    prefact = np.float32(float(np.float32(R)) * 2.0 * 3.141592653589793 / float(np.float32(BoxSize)))
    dk = _fft_field(lib, delta, dims, dev, torch.cuda.current_stream(dev))
    _lib.check(lib.pylb_tophat_k(dk.data_ptr(), dims, float(prefact), _stream(dev)), "pylb_tophat_k")
    out = torch.empty((dims, dims, dims), dtype=torch.float32, device=dev)
    _lib.check(lib.pylb_fft_c2r(dk.data_ptr(), out.data_ptr(), dims, 1, _stream(dev)), "pylb_fft_c2r")
    return _like_input(out, delta)
