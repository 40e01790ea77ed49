"""fluidsim_b200 -- B200-native pseudo-spectral time step behind the fluidsim / fluidfft API.

Public surface (host-side mirror of the reference interface for this path):

* ``fluidsim_b200.fft.FFT3DWithB200 / FFT2DWithB200``  -- fluidfft-plugin-shaped FFT classes
* ``fluidsim_b200.operators.OperatorsPseudoSpectral3D / 2D``
* ``fluidsim_b200.solvers.SimulNS3D / SimulNS3DStrat / SimulNS2D`` with
  ``sim.tendencies_nonlin`` and ``sim.time_stepping`` (``TimeSteppingPseudoSpectralB200``)

Everything runs through ``libb200spectral.so`` (hand-written sm_100a CUDA, C ABI in
``include/b200spectral.h``); importing the submodules raises if the library is not built.
"""

__version__ = "0.1.0"
