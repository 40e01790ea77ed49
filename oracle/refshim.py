"""Import the UNMODIFIED reference hot-path modules from /root/reference (TEST INFRASTRUCTURE).

Only usable in the build container (the GPU box has no /root/reference).  The
reference's pip dependencies (transonic, fluiddyn, fluidfft, h5py ...) are absent and
cannot be installed offline, so this module registers small stub modules in
``sys.modules``:

* ``transonic``      -- identity ``boost``; ``Transonic().is_transpiled == False`` so every
                        ``ts.use_block`` site takes its NumPy branch
                        (``base/time_stepping/pseudo_spect.py:128-140,920-984``).
* ``fluiddyn``       -- ``util.mpi`` (rank 0 / nb_proc 1), ``calcul.setofvariables``,
                        ``util.paramcontainer``, ``io``.
* ``fluidfft``       -- ``fft3d.operators`` / ``fft2d.operators`` -> ``oracle.fluidfft_np``.
* ``fluidsim``       -- namespace package over /root/reference/fluidsim so that the real
                        ``fluidsim/__init__.py`` (h5py, matplotlib ...) is not executed.
* ``fluidsim_core``, ``h5py``, ``h5netcdf`` -- empty stand-ins for import statements only.

After ``install()`` the reference's own ``pseudo_spect.py``, ``solvers/ns3d/solver.py``,
``operators3d.py`` ... are importable and are executed as-is by ``RefSim*`` below,
which wire them together the way ``SimulBase.__init__`` does
(``base/solvers/base.py:117-223``) but without the params/output machinery.
"""

import os
import sys
import types
from copy import deepcopy

import numpy as np

REF_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "fluidsim"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Anything:
    """Permissive placeholder for typing-only names (Array, Type, NDim ...)."""

    def __init__(self, *a, **k):
        pass

    def __getitem__(self, item):
        return self

    def __call__(self, *a, **k):
        return self

    def __sub__(self, other):
        return self

    def __add__(self, other):
        return self

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return self


class ParamContainer:
    """Tiny attribute tree standing in for fluiddyn's ParamContainer."""

    def __init__(self, tag="params", attribs=None, **kwargs):
        self._tag = tag
        if attribs:
            self.__dict__.update(attribs)

    def _set_child(self, tag, attribs=None, **kwargs):
        child = ParamContainer(tag=tag, attribs=attribs)
        setattr(self, tag, child)
        return child

    def _set_attribs(self, attribs):
        self.__dict__.update(attribs)

    def _set_attrib(self, key, value):
        setattr(self, key, value)

    def _set_doc(self, doc):
        self._doc = doc

    _doc = ""


_installed = False


def install():
    """Register the stub modules (idempotent)."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not available at " + REF_ROOT)
    from . import fluidfft_np

    # transonic ---------------------------------------------------------------
    def boost(obj=None, **kwargs):
        if obj is None:
            return lambda o: o
        return obj

    class Transonic:
        is_transpiled = False
        is_transpiling = False
        is_compiled = False

        def use_block(self, name):
            raise RuntimeError("transonic stub: blocks are never transpiled")

    any_ = _Anything()
    _mod(
        "transonic",
        boost=boost,
        Transonic=Transonic,
        Type=any_,
        NDim=any_,
        Array=any_,
        Union=any_,
        const=any_,
    )

    # fluiddyn ----------------------------------------------------------------
    mpi = _mod(
        "fluiddyn.util.mpi",
        rank=0,
        nb_proc=1,
        comm=None,
        MPI=None,
        printby0=print,
        print_sorted=print,
    )
    util = _mod("fluiddyn.util", mpi=mpi)
    pc = _mod("fluiddyn.util.paramcontainer", ParamContainer=ParamContainer)
    util.paramcontainer = pc
    sov = _mod("fluiddyn.calcul.setofvariables", SetOfVariables=fluidfft_np.SetOfVariables)
    calcul = _mod("fluiddyn.calcul", setofvariables=sov)
    io_ = _mod("fluiddyn.io", FLUIDSIM_PATH="/tmp", FLUIDDYN_PATH_SCRATCH=None)
    _mod(
        "fluiddyn",
        util=util,
        calcul=calcul,
        io=io_,
        time_as_str=lambda *a, **k: "0000-00-00_00-00-00",
    )

    # fluidfft ----------------------------------------------------------------
    ops3 = _mod(
        "fluidfft.fft3d.operators",
        OperatorsPseudoSpectral3D=fluidfft_np.OperatorsPseudoSpectral3D,
        vector_product=fluidfft_np.vector_product,
    )
    ops2 = _mod(
        "fluidfft.fft2d.operators",
        OperatorsPseudoSpectral2D=fluidfft_np.OperatorsPseudoSpectral2D,
    )
    f3 = _mod("fluidfft.fft3d", operators=ops3)
    f2 = _mod("fluidfft.fft2d", operators=ops2)
    _mod("fluidfft", fft3d=f3, fft2d=f2)

    # io libs (import statements only)
    _mod("h5py")
    _mod("h5netcdf")

    # fluidsim_core (params plumbing only) --------------------------------------
    class Parameters(ParamContainer):
        pass

    class SimulCore:
        pass

    core_params = _mod(
        "fluidsim_core.params", Parameters=Parameters, iter_complete_params=lambda *a, **k: None
    )
    core_solver = _mod("fluidsim_core.solver", SimulCore=SimulCore)
    core_info = _mod("fluidsim_core.info", InfoSolverCore=ParamContainer, create_info_simul=lambda *a, **k: None)
    _mod("fluidsim_core", params=core_params, solver=core_solver, info=core_info)

    # fluidsim namespace ---------------------------------------------------------
    pkg = _mod("fluidsim", _is_testing=False)
    pkg.__path__ = [os.path.join(REF_ROOT, "fluidsim")]
    # base/params.py and base/solvers/info_base.py pull in the whole info/param
    # machinery; the hot path only needs the names.
    _mod("fluidsim.base.params", Parameters=Parameters)

    _installed = True


# ----------------------------------------------------------------------------- wiring
class _NS:
    """Attribute bag."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def make_params(solver, nx, ny, nz=None, Lx=2 * np.pi, Ly=2 * np.pi, Lz=2 * np.pi, **kw):
    """Duck-typed ``params`` with the attribute names the reference consumes."""
    p = ParamContainer()
    p.ONLY_COARSE_OPER = False
    p.short_name_type_run = ""
    oper = dict(
        type_fft="default",
        coef_dealiasing=kw.pop("coef_dealiasing", 2.0 / 3),
        nx=nx,
        ny=ny,
        Lx=Lx,
        Ly=Ly,
        truncation_shape=kw.pop("truncation_shape", "cubic"),
        NO_SHEAR_MODES=False,
    )
    if nz is not None:
        oper.update(nz=nz, Lz=Lz, type_fft2d="sequential")
    else:
        oper.update(NO_KY0=False)
    p._set_child("oper", oper)
    p.nu_2 = kw.pop("nu_2", 0.0)
    p.nu_4 = kw.pop("nu_4", 0.0)
    p.nu_8 = kw.pop("nu_8", 0.0)
    p.nu_m4 = kw.pop("nu_m4", 0.0)
    p._set_child(
        "time_stepping",
        dict(
            USE_T_END=False,
            t_end=10.0,
            it_end=10,
            USE_CFL=False,
            type_time_scheme=kw.pop("type_time_scheme", "RK4"),
            deltat0=kw.pop("deltat0", 1e-2),
            deltat_max=0.2,
            cfl_coef=None,
            max_elapsed=None,
        ),
    )
    # TimeSteppingPseudoSpectral._complete_params_with_default (pseudo_spect.py:159-167)
    p.time_stepping._set_child(
        "phaseshift_random",
        dict(nb_pairs=kw.pop("nb_pairs", 1), nb_steps_compute_new_pair=kw.pop("nb_steps_compute_new_pair", None)),
    )
    p._set_child("forcing", dict(enable=False))
    if solver.startswith("ns3d"):
        p.f = kw.pop("f", None)
        p.no_vz_kz0 = bool(kw.pop("no_vz_kz0", False))
        p.projection = kw.pop("projection", None)
    if solver in ("ns3d.strat", "ns3d.bouss"):
        p.N = kw.pop("N", 1.0)
    if solver.startswith("ns2d"):
        p.beta = kw.pop("beta", 0.0)
    if solver == "ns2d.strat":  # ns2d/strat/solver.py:65-69
        p.N = kw.pop("N", 1.0)
    if kw:
        raise TypeError(f"unknown params {sorted(kw)}")
    return p


class RefSim:
    """Wire the reference's own classes for one solver (ns3d | ns3d.strat | ns3d.bouss | ns2d |
    ns2d.strat | ns2d.bouss).

    Mirrors ``SimulBase.__init__`` (``base/solvers/base.py:117-223``): Operators ->
    State -> TimeStepping, then fields are set by the caller with ``set_state_spect``.
    Everything executed below ``self.time_stepping.one_time_step_computation()`` is the
    reference's code; only the fluidfft layer is ``oracle.fluidfft_np``.
    """

    def __init__(self, solver, params):
        install()
        import importlib

        self.solver = solver
        self.params = params
        if solver.startswith("ns2d"):
            from fluidsim.operators.operators2d import OperatorsPseudoSpectral2D as Oper
        else:
            from fluidsim.operators.operators3d import OperatorsPseudoSpectral3D as Oper
        self.oper = Oper(params)

        modsolver = importlib.import_module(f"fluidsim.solvers.{solver}.solver")
        SimulRef = modsolver.Simul
        statemod, statecls, tsmod, tscls = {
            "ns3d": ("ns3d.state", "StateNS3D", "ns3d.time_stepping", "TimeSteppingPseudoSpectralNS3D"),
            "ns3d.strat": (
                "ns3d.strat.state",
                "StateNS3DStrat",
                "ns3d.time_stepping",
                "TimeSteppingPseudoSpectralNS3D",
            ),
            "ns3d.bouss": (
                "ns3d.strat.state",
                "StateNS3DStrat",
                "ns3d.time_stepping",
                "TimeSteppingPseudoSpectralNS3D",
            ),
            "ns2d": ("ns2d.state", "StateNS2D", None, None),
            # ns2d.strat has its own time-stepping class, but it only changes the CFL rule
            # (ns2d/strat/time_stepping.py:31-110), which the shim does not run
            "ns2d.strat": ("ns2d.strat.state", "StateNS2DStrat", None, None),
            "ns2d.bouss": ("ns2d.bouss.state", "StateNS2DBouss", None, None),
        }[solver]
        State = getattr(importlib.import_module("fluidsim.solvers." + statemod), statecls)
        if tsmod is None:
            from fluidsim.base.time_stepping.pseudo_spect import TimeSteppingPseudoSpectral as TS
        else:
            TS = getattr(importlib.import_module("fluidsim.solvers." + tsmod), tscls)

        # a Simul instance without running SimulBase.__init__ (params/output plumbing)
        sim = SimulRef.__new__(SimulRef)
        sim.params = params
        sim.oper = self.oper
        sim.is_forcing_enabled = False
        sim.info_solver = self._info_solver(State)
        sim.info = _NS(solver=sim.info_solver)
        sim.output = _NS()
        self.sim = sim
        sim.state = State(sim)
        if solver.startswith("ns3d"):
            sim._init_projection()
        sim.time_stepping = self._make_time_stepping(TS, sim)

    @staticmethod
    def _info_solver(State):
        info = ParamContainer()
        info._set_child("classes")
        info.classes._set_child("State")
        State._complete_info_solver(info)
        return info

    @staticmethod
    def _make_time_stepping(TS, sim):
        ts = TS.__new__(TS)
        ts.sim = sim
        ts.params = sim.params
        ts.it = 0
        ts.t = 0.0
        ts.deltat = float(sim.params.time_stepping.deltat0)
        ts._has_to_stop = False
        # TimeSteppingPseudoSpectral.init_from_params minus _init_compute_time_step
        ts._init_freq_lin()
        ts._init_exact_linear_coef()
        ts._init_time_scheme()
        return ts

    # state I/O ------------------------------------------------------------------
    def set_state_spect(self, arr):
        st = self.sim.state
        st.state_spect[...] = arr
        st.statephys_from_statespect()

    def step(self):
        ts = self.sim.time_stepping
        ts.one_time_step_computation()
        ts.t += ts.deltat
        ts.it += 1
        return np.array(self.sim.state.state_spect)
