"""fluidsim solver module for the key ``ns3d.strat.b200`` (entry point in ``pyproject.toml``): the GPU
``Simul`` of ``fluidsim_b200.solvers.SimulNS3DStrat`` (mirror of ``fluidsim.solvers.ns3d.strat.solver``)."""

from ..solvers import SimulNS3DStrat
from . import make_info_solver


class Simul(SimulNS3DStrat):
    """``fluidsim.load / fluidsim-bench -s ns3d.strat.b200`` entry: same constructor contract as the
    reference solver (``Simul(params)``, ``Simul.create_default_params()``)."""

    InfoSolver = make_info_solver("fluidsim.solvers.ns3d.strat.solver", "InfoSolverNS3DStrat", __name__, "ns3d.strat.b200", "StateNS3DStrat", SimulNS3DStrat)


__all__ = ["Simul"]
