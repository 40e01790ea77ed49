"""GPU-resident state containers (mirror of ``/root/reference/fluidsim/base/state.py:209-332``,
``solvers/ns3d/state.py:43-52`` and ``solvers/ns2d/state.py:30-113``).

``state_spect`` is the primary state and lives on the GPU for the whole run.  ``state_phys`` is
LAZY: the reference refreshes it after every step (3-4 inverse FFTs,
``solvers/ns3d/time_stepping.py:17``); here it is recomputed only when something reads it (CFL,
outputs), which removes those transforms from the fused step.
"""

import torch

from .setofvariables import SetOfVariables


class StatePseudoSpectral:
    keys_state_phys = ()
    keys_state_spect = ()

    def __init__(self, sim, oper=None):
        self.sim = sim
        self.params = sim.params
        self.oper = sim.oper if oper is None else oper
        oper = self.oper
        self.state_spect = SetOfVariables(
            keys=self.keys_state_spect,
            shape_variable=oper.shapeK_loc,
            dtype=torch.complex128,
            info="state_spect",
            value=0.0,
            device=oper.device,
        )
        # Any writable view of state_spect handed out through the container's own interface
        # (get_var, indexing, set_var, initialize, +=) may be used to edit the state in place, the
        # reference idiom.  The pruned transforms need a dealiased state, so such an access drops the
        # "state is dealiased" knowledge; the next fused step re-establishes it with one device-side
        # check (b2_check_dealiased).  Only raw ``state_spect.tensor`` writes need an explicit
        # ``mark_spect_modified()``.
        self.state_spect._on_touch = self._spect_touched
        self._state_phys = None
        self._phys_dirty = True
        self.vars_computed = {}
        self.it_computed = {}

    # ---- lazy physical state ------------------------------------------------------------------------
    @property
    def state_phys(self):
        if self._state_phys is None:
            self._state_phys = SetOfVariables(
                keys=self.keys_state_phys,
                shape_variable=self.oper.shapeX_loc,
                dtype=torch.float64,
                info="state_phys",
                value=0.0,
                device=self.oper.device,
            )
            self._phys_dirty = True
        if self._phys_dirty:
            self._phys_dirty = False
            self._statephys_from_statespect()
        return self._state_phys

    def _spect_touched(self):
        self._phys_dirty = True
        self.sim._state_dealiased = False

    def mark_spect_modified(self):
        """Call after writing ``state_spect`` from outside the time stepper: invalidates the lazy
        ``state_phys`` and the "state is dealiased" knowledge the pruned transforms rely on."""
        self._phys_dirty = True
        self.sim._state_dealiased = False

    def statephys_from_statespect(self):
        """base/state.py:326-332 -- deferred until ``state_phys`` is read.  Called by user code after
        editing ``state_spect`` (reference idiom): the state may no longer be dealiased."""
        self._phys_dirty = True
        self.sim._state_dealiased = False

    def _statephys_from_statespect(self):
        ifft_as_arg = self.oper.ifft_as_arg
        for ik in range(self.state_spect.nvar):  # raw tensors: reading is not a "touch"
            ifft_as_arg(self.state_spect.tensor[ik], self._state_phys.tensor[ik])

    def statespect_from_statephys(self):
        """base/state.py:318-324.  Uses the physical arrays as they are (edits made through a held
        reference included): no lazy refresh from state_spect here."""
        if self._state_phys is None:
            phys = self.state_phys
        else:
            phys = self._state_phys
        fft_as_arg = self.oper.fft_as_arg
        for ik in range(self.state_spect.nvar):
            fft_as_arg(phys.tensor[ik], self.state_spect.tensor[ik])
        self._phys_dirty = False
        self.sim._state_dealiased = False

    def get_var(self, key):
        if key in self.keys_state_spect:
            return self.state_spect.get_var(key)
        if key in self.keys_state_phys:
            return self.state_phys.get_var(key)
        raise ValueError(f'Do not know how to compute "{key}".')

    def init_statespect_from(self, **kwargs):
        """base/state.py: set the given keys, zero the others."""
        self.state_spect.initialize(0.0)
        for key, value in kwargs.items():
            if key not in self.keys_state_spect:
                raise ValueError(f"{key} is not a key of state_spect")
            self.state_spect.set_var(key, value)
        self.mark_spect_modified()

    def check_energy_equal_phys_spect(self):
        """base/state.py:385-392."""
        energy_phys = self.compute_energy_phys()
        energy_spect = self.compute_energy_spect()
        return abs(energy_phys - energy_spect) <= 1e-8 + 1e-5 * abs(energy_spect)


class StateNS3D(StatePseudoSpectral):
    keys_state_phys = ("vx", "vy", "vz")
    keys_state_spect = ("vx_fft", "vy_fft", "vz_fft")

    def compute_energy_phys(self):
        p = self.state_phys
        return 0.5 * float((p[0] ** 2 + p[1] ** 2 + p[2] ** 2).mean())

    def compute_energy_spect(self):
        return 0.5 * self.oper.oper_fft.sum_wavenumbers_abs2(self.state_spect.tensor[:3])


class StateNS3DStrat(StateNS3D):
    keys_state_phys = ("vx", "vy", "vz", "b")
    keys_state_spect = ("vx_fft", "vy_fft", "vz_fft", "b_fft")


class StateNS2D(StatePseudoSpectral):
    keys_state_phys = ("ux", "uy", "rot")
    keys_state_spect = ("rot_fft",)

    def _statephys_from_statespect(self):
        """solvers/ns2d/state.py:95-106."""
        oper = self.oper
        rot_fft = self.state_spect.tensor[0]
        ux_fft, uy_fft = oper.vecfft_from_rotfft(rot_fft)
        oper.ifft_as_arg(rot_fft, self._state_phys.get_var("rot"))
        ifft_as_arg_destroy = oper.oper_fft.ifft_as_arg_destroy
        ifft_as_arg_destroy(ux_fft, self._state_phys.get_var("ux"))
        ifft_as_arg_destroy(uy_fft, self._state_phys.get_var("uy"))

    def statespect_from_statephys(self):
        """solvers/ns2d/state.py:108-113."""
        phys = self.state_phys if self._state_phys is None else self._state_phys
        self.oper.fft_as_arg(phys.get_var("rot"), self.state_spect.tensor[0])
        self._phys_dirty = False
        self.sim._state_dealiased = False

    def compute_energy_phys(self):
        p = self.state_phys
        return 0.5 * float((p.get_var("ux") ** 2 + p.get_var("uy") ** 2).mean())

    def compute_energy_spect(self):
        oper = self.oper
        rot_fft = self.state_spect.tensor[0]
        return oper.sum_wavenumbers(0.5 * rot_fft.abs() ** 2 / oper.K2_not0)


class StateNS2DStrat(StateNS2D):
    """solvers/ns2d/strat/state.py:19-163 and solvers/ns2d/bouss/state.py:19-122 (same state)."""

    keys_state_phys = ("ux", "uy", "rot", "b")
    keys_state_spect = ("rot_fft", "b_fft")

    def _statephys_from_statespect(self):
        """ns2d/strat/state.py:137-151."""
        super()._statephys_from_statespect()
        self.oper.ifft_as_arg(self.state_spect.tensor[1], self._state_phys.get_var("b"))

    def statespect_from_statephys(self):
        """ns2d/strat/state.py:153-163."""
        phys = self.state_phys if self._state_phys is None else self._state_phys
        self.oper.fft_as_arg(phys.get_var("rot"), self.state_spect.tensor[0])
        self.oper.fft_as_arg(phys.get_var("b"), self.state_spect.tensor[1])
        self._phys_dirty = False
        self.sim._state_dealiased = False

    def init_from_rotbfft(self, rot_fft, b_fft):
        """ns2d/strat/state.py:165-173."""
        self.oper.dealiasing(rot_fft)
        self.oper.dealiasing(b_fft)
        self.state_spect.set_var("rot_fft", rot_fft)
        self.state_spect.set_var("b_fft", b_fft)
        self.statephys_from_statespect()

    def init_from_rotfft(self, rot_fft):
        """ns2d/strat/state.py:233-236."""
        self.init_from_rotbfft(rot_fft, self.oper.create_arrayK(value=0.0))
