"""Minimal ``params`` container with the attribute names the reference consumes.

The real fluidsim builds a ``fluiddyn`` ``ParamContainer`` tree through each class's
``_complete_params_with_default`` (``/root/reference/fluidsim/base/time_stepping/base.py:33-94``,
``base/time_stepping/pseudo_spect.py:158-167``, ``operators/operators3d.py:152-205``,
``operators/operators2d.py:100-113``, ``base/solvers/pseudo_spect.py:106-132``,
``solvers/ns3d/solver.py`` / ``strat/solver.py`` / ``ns2d/solver.py``).  The GPU classes are
duck-typed on attribute names, so a real fluidsim ``params`` object works as well; this module only
provides the same tree with the same defaults when fluidsim is not installed.
"""

from math import pi


class ParamContainer:
    def __init__(self, tag="params", **attribs):
        self._tag = tag
        self.__dict__.update(attribs)

    def _set_child(self, tag, attribs=None):
        child = ParamContainer(tag, **(attribs or {}))
        setattr(self, tag, child)
        return child

    def _set_attribs(self, attribs):
        self.__dict__.update(attribs)

    def _set_attrib(self, key, value):
        setattr(self, key, value)

    def __repr__(self):
        items = {k: v for k, v in self.__dict__.items() if not k.startswith("_")}
        return f"<params {self._tag}: {items}>"


def create_default_params(solver="ns3d"):
    """Same names and defaults as ``Simul.create_default_params()`` for ns3d / ns3d.strat / ns2d."""
    p = ParamContainer()
    p.ONLY_COARSE_OPER = False
    p.short_name_type_run = ""
    p.NEW_DIR_RESULTS = True
    p.nu_2 = 0.0
    p.nu_4 = 0.0
    p.nu_8 = 0.0
    p.nu_m4 = 0.0
    if solver in ("ns3d", "ns3d.strat", "ns3d.bouss"):
        p._set_child(
            "oper",
            dict(
                type_fft="fft3d.with_b200",
                type_fft2d="fft2d.with_b200",
                coef_dealiasing=2.0 / 3,
                nx=48,
                ny=48,
                nz=48,
                Lx=2 * pi,
                Ly=2 * pi,
                Lz=2 * pi,
                truncation_shape="cubic",
                NO_SHEAR_MODES=False,
            ),
        )
        p.f = None
        p.no_vz_kz0 = False
        p.projection = None
        if solver == "ns3d.strat":
            p.N = 1.0
    elif solver in ("ns2d", "ns2d.strat", "ns2d.bouss"):
        p._set_child(
            "oper",
            dict(
                type_fft="fft2d.with_b200",
                coef_dealiasing=2.0 / 3,
                nx=48,
                ny=48,
                Lx=8,
                Ly=8,
                truncation_shape="cubic",
                NO_SHEAR_MODES=False,
                NO_KY0=False,
            ),
        )
        p.beta = 0.0
        if solver == "ns2d.strat":  # ns2d/strat/solver.py:65-69
            p.N = 1.0
    else:
        raise ValueError(f"unknown solver {solver!r}")
    p._set_child(
        "time_stepping",
        dict(
            USE_T_END=True,
            t_end=10.0,
            it_end=10,
            USE_CFL=True,
            type_time_scheme="RK4",
            deltat0=0.2,
            deltat_max=0.2,
            cfl_coef=None,
            max_elapsed=None,
        ),
    )
    if solver == "ns2d.strat":  # ns2d/strat/time_stepping.py:24-29
        p.time_stepping.cfl_coef_group = None
    # pseudo_spect.py:159-167
    p.time_stepping._set_child("phaseshift_random", dict(nb_pairs=1, nb_steps_compute_new_pair=None))
    # base/forcing/base.py:63-76, 189-196; specific.py:436-451, 777-786
    p._set_child(
        "forcing",
        dict(
            enable=False,
            type="",
            available_types=["in_script", "proportional", "tcrandom"],
            forcing_rate=1.0,
            key_forced=None,
            nkmax_forcing=5,
            nkmin_forcing=4,
            random_seed=None,
        ),
    )
    p.forcing._set_child("normalized", dict(type="2nd_degree_eq", which_root="minabs", constant_rate_of=None))
    p.forcing._set_child("random", dict(only_positive=False))
    p.forcing._set_child("tcrandom", dict(time_correlation="based_on_forcing_rate"))
    p._set_child("init_fields", dict(type="constant"))
    p.init_fields._set_child("noise", dict(velo_max=1.0, length=None))
    p.init_fields._set_child("from_file", dict(path=""))  # base/init_fields.py:144-151
    p._set_child("output", dict(path_run=None))
    return p
