"""fluidsim solver module for the key ``ns2d.b200`` (entry point in ``pyproject.toml``): the GPU
``Simul`` of ``fluidsim_b200.solvers.SimulNS2D`` (mirror of ``fluidsim.solvers.ns2d.solver``)."""

from ..solvers import SimulNS2D
from . import make_info_solver


class Simul(SimulNS2D):
    """``fluidsim.load / fluidsim-bench -s ns2d.b200`` entry: same constructor contract as the
    reference solver (``Simul(params)``, ``Simul.create_default_params()``)."""

    InfoSolver = make_info_solver("fluidsim.solvers.ns2d.solver", "InfoSolverNS2D", __name__, "ns2d.b200", "StateNS2D", SimulNS2D)


__all__ = ["Simul"]
