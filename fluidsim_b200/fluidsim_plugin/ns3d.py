"""fluidsim solver module for the key ``ns3d.b200`` (entry point in ``pyproject.toml``): the GPU
``Simul`` of ``fluidsim_b200.solvers.SimulNS3D`` (mirror of ``fluidsim.solvers.ns3d.solver``)."""

from ..solvers import SimulNS3D
from . import make_info_solver


class Simul(SimulNS3D):
    """``fluidsim.load / fluidsim-bench -s ns3d.b200`` entry: same constructor contract as the
    reference solver (``Simul(params)``, ``Simul.create_default_params()``)."""

    InfoSolver = make_info_solver("fluidsim.solvers.ns3d.solver", "InfoSolverNS3D", __name__, "ns3d.b200", "StateNS3D", SimulNS3D)


__all__ = ["Simul"]
