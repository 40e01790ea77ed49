"""fluidsim solver module for the key ``ns3d.bouss.b200`` (entry point in ``pyproject.toml``): the GPU
``Simul`` of ``fluidsim_b200.solvers.SimulNS3DBouss`` (mirror of ``fluidsim.solvers.ns3d.bouss.solver``)."""

from ..solvers import SimulNS3DBouss
from . import make_info_solver


class Simul(SimulNS3DBouss):
    """``fluidsim.load / fluidsim-bench -s ns3d.bouss.b200`` entry: same constructor contract as the
    reference solver (``Simul(params)``, ``Simul.create_default_params()``)."""

    InfoSolver = make_info_solver("fluidsim.solvers.ns3d.bouss.solver", "InfoSolverNS3DBouss", __name__,
                                  "ns3d.bouss.b200", "StateNS3DStrat", SimulNS3DBouss)


__all__ = ["Simul"]
