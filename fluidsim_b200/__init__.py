"""fluidsim_b200 -- B200-native pseudo-spectral time step behind the fluidsim / fluidfft API."""

__version__ = "0.1.0"
