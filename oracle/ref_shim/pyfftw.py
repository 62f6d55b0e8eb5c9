"""Minimal stand-in for the `pyfftw` module, used ONLY to run the compiled reference
(oracle/_ref) in an image that has neither pyfftw nor FFTW.  TEST INFRASTRUCTURE.

Covers exactly what Pk_library.pyx:120-245 touches: `empty_aligned(shape, dtype)` and
`FFTW(a_in, a_out, axes, flags, direction, threads)(a_in, a_out)`.  Transforms are done by
scipy.fft (pocketfft), which keeps float32 -> complex64 like the reference's FFT3Dr_f.
Shapes arrive as floats because `dims/2+1` is true division under python 3.

Normalisation follows pyfftw, not raw FFTW: `pyfftw.FFTW.__call__(input_array=None, output_array=None,
normalise_idft=True, ortho=False)` scales a BACKWARD transform by 1/N unless told otherwise, and the reference
calls its plans with two positional arguments only (Pk_library.pyx:165, 229), so IFFT3Dr_f / IFFT2Dr_f return the
normalised inverse.  The reference's own arithmetic confirms it: `Xi` divides the inverse transform of
|delta_k|^2 by dims^3 exactly once (Pk_library.pyx:2139-2143), which is the correlation function only if the
inverse transform already carried the other 1/dims^3.  scipy's ifftn/irfftn have the same convention.
"""
import numpy as np
import scipy.fft as _sf


def empty_aligned(shape, dtype="float64", n=None):
    return np.empty(tuple(int(s) for s in shape), dtype=dtype)


class FFTW(object):
    def __init__(self, a_in, a_out, axes=(0,), flags=(), direction="FFTW_FORWARD", threads=1):
        self.axes = tuple(axes)
        self.direction = direction
        self.threads = int(threads)

    def __call__(self, a_in, a_out):
        if self.direction == "FFTW_FORWARD":
            if np.iscomplexobj(a_in):
                a_out[...] = _sf.fftn(a_in, axes=self.axes, workers=self.threads)
            else:
                a_out[...] = _sf.rfftn(a_in, axes=self.axes, workers=self.threads)
        else:
            if np.iscomplexobj(a_out):
                a_out[...] = _sf.ifftn(a_in, axes=self.axes, workers=self.threads)
            else:
                s = [a_out.shape[ax] for ax in self.axes]
                a_out[...] = _sf.irfftn(a_in, s=s, axes=self.axes, workers=self.threads)
        return a_out
