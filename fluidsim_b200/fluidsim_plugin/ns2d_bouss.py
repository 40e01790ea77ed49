"""fluidsim solver module for the key ``ns2d.bouss.b200`` (entry point in ``pyproject.toml``): the GPU
``Simul`` of ``fluidsim_b200.solvers.SimulNS2DBouss`` (mirror of ``fluidsim.solvers.ns2d.bouss.solver``)."""

from ..solvers import SimulNS2DBouss
from . import make_info_solver


class Simul(SimulNS2DBouss):
    """``fluidsim.load / fluidsim-bench -s ns2d.bouss.b200`` entry: same constructor contract as the
    reference solver (``Simul(params)``, ``Simul.create_default_params()``)."""

    InfoSolver = make_info_solver("fluidsim.solvers.ns2d.bouss.solver", "InfoSolverNS2DBouss", __name__,
                                  "ns2d.bouss.b200", "StateNS2DStrat", SimulNS2DBouss)


__all__ = ["Simul"]
