"""fluidfft plugin module for ``type_fft = "fft3d.with_b200"``.

fluidfft >= 0.4 resolves ``params.oper.type_fft`` through the entry-point group
``fluidfft.plugins`` to a module exposing ``FFTclass``
(``/root/reference/fluidsim/operators/operators3d.py:156,229``); register this module as

    [project.entry-points."fluidfft.plugins"]
    "fft3d.with_b200" = "fluidsim_b200.fft3d_with_b200"
"""

from .fft import FFT3DWithB200 as FFTclass

__all__ = ["FFTclass"]
