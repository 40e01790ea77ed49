"""fluidfft plugin module for ``type_fft = "fft2d.with_b200"`` (see ``fft3d_with_b200``)."""

from .fft import FFT2DWithB200 as FFTclass

__all__ = ["FFTclass"]
