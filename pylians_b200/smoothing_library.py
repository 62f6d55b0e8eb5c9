"""Field smoothing on the GPU: mirror of library/smoothing_library/smoothing_library.pyx (`FT_filter` :19-83,
`field_smoothing` :89-114) -- a consumer of PKL.FFT3Dr_f / IFFT3Dr_f.

    W_k   = SL.FT_filter(BoxSize, R, dims, 'Top-Hat' | 'Gaussian', threads)
    field = SL.field_smoothing(field, W_k, threads)

numpy in -> numpy out, CUDA tensors in -> CUDA tensors out (a filter kept on the device can be reused for many
fields without touching the host).  The filter is built, summed and normalised by two streaming kernels, both
transforms are cuFFT, the mode loop is one complex multiply per stored mode.  `threads` is accepted and ignored."""
import numpy as np
import torch

from . import _lib
from .MAS_library import _device, _is_torch
from .Pk_library import _check_field, _fft_field, _like_input, _stream


def FT_filter(BoxSize, R, dims, Filter, threads=1, device_out=False):
    """smoothing_library.pyx:19-83.  Returns complex64 (dims,dims,dims/2+1): numpy, or a CUDA tensor with device_out."""
    if Filter not in ["Top-Hat", "Gaussian"]:
        raise Exception("Filter %s not implemented!" % Filter)
    lib, dev = _lib.load(), _device()
    dims = int(dims)
    # `float BoxSize, float R`; R_grid = (R*dims/BoxSize) and R2 = R_grid**2 are C floats (:31-32)
    R_grid = np.float32(np.float32(np.float32(R) * np.float32(dims)) / np.float32(BoxSize))
    R2 = np.float32(R_grid * R_grid)
    field = torch.empty((dims, dims, dims), dtype=torch.float32, device=dev)
    scratch = torch.empty(1, dtype=torch.float64, device=dev)
    _lib.check(lib.pylb_filter_real(field.data_ptr(), dims, float(R2), 0 if Filter == "Top-Hat" else 1,
                                    scratch.data_ptr(), _stream(dev)), "pylb_filter_real")
    field_k = _fft_field(lib, field, dims, dev, torch.cuda.current_stream(dev))
    return field_k if device_out else field_k.cpu().numpy()


def field_smoothing(field, filter_k, threads=1):
    """smoothing_library.pyx:89-114: IFFT(FFT(field)*filter_k).  `field` float32 (dims,dims,dims) is not modified."""
    lib, dev = _lib.load(), _device()
    field = _check_field(field)
    dims = field.shape[0]
    if dims != filter_k.shape[0]:
        raise Exception("field and filter have different grids!!!")
    fk = filter_k if _is_torch(filter_k) else torch.from_numpy(np.ascontiguousarray(filter_k))
    if fk.dtype != torch.complex64:
        raise ValueError("Buffer dtype mismatch, expected 'complex64_t' but got '%s'" % str(fk.dtype).replace("torch.", ""))
    fk = fk.to(dev, non_blocking=True).contiguous()
    field_k = _fft_field(lib, field, dims, dev, torch.cuda.current_stream(dev))       # a fresh buffer
    _lib.check(lib.pylb_cmul_c64(field_k.data_ptr(), fk.data_ptr(), field_k.numel(), _stream(dev)), "pylb_cmul_c64")
    out = torch.empty((dims, dims, dims), dtype=torch.float32, device=dev)
    _lib.check(lib.pylb_fft_c2r(field_k.data_ptr(), out.data_ptr(), dims, 1, _stream(dev)), "pylb_fft_c2r")
    return _like_input(out, field)
