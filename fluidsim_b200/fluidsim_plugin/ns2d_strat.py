"""fluidsim solver module for the key ``ns2d.strat.b200`` (entry point in ``pyproject.toml``): the GPU
``Simul`` of ``fluidsim_b200.solvers.SimulNS2DStrat`` (mirror of ``fluidsim.solvers.ns2d.strat.solver``)."""

from ..solvers import SimulNS2DStrat
from . import make_info_solver


class Simul(SimulNS2DStrat):
    """``fluidsim.load / fluidsim-bench -s ns2d.strat.b200`` entry: same constructor contract as the
    reference solver (``Simul(params)``, ``Simul.create_default_params()``)."""

    InfoSolver = make_info_solver("fluidsim.solvers.ns2d.strat.solver", "InfoSolverNS2DStrat", __name__,
                                  "ns2d.strat.b200", "StateNS2DStrat", SimulNS2DStrat)


__all__ = ["Simul"]
