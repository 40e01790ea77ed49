"""Solver modules registered in fluidsim's entry-point groups (``pyproject.toml``):

    fluidsim.solvers.ns3d : b200       -> fluidsim_b200.fluidsim_plugin.ns3d        (key "ns3d.b200")
    fluidsim.solvers.ns3d : strat.b200 -> fluidsim_b200.fluidsim_plugin.ns3d_strat  (key "ns3d.strat.b200")
    fluidsim.solvers.ns3d : bouss.b200 -> fluidsim_b200.fluidsim_plugin.ns3d_bouss  (key "ns3d.bouss.b200")
    fluidsim.solvers.ns2d : b200       -> fluidsim_b200.fluidsim_plugin.ns2d        (key "ns2d.b200")
    fluidsim.solvers.ns2d : strat.b200 -> fluidsim_b200.fluidsim_plugin.ns2d_strat  (key "ns2d.strat.b200")
    fluidsim.solvers.ns2d : bouss.b200 -> fluidsim_b200.fluidsim_plugin.ns2d_bouss  (key "ns2d.bouss.b200")

fluidsim resolves a solver key to such a module and takes its ``Simul`` class
(``/root/reference/lib/fluidsim_core/loader.py:17-74``, ``fluidsim/util/util.py:75-106``); the class
must offer ``create_default_params()`` and be constructible from the params
(``fluidsim-bench -s ns3d.b200``: ``/root/reference/fluidsim/util/console/bench.py:275``).

Each module exposes the GPU ``Simul`` of ``fluidsim_b200.solvers`` under that contract and, when the
real fluidsim is importable, an ``InfoSolver`` subclass describing the class tree the way fluidsim's
own derived solvers do (``/root/reference/fluidsim/solvers/ns3d/strat/solver.py:36-54``).
"""


def make_info_solver(parent_module, parent_class, module_name, short_name, state_class, simul_class):
    """InfoSolver subclass of the reference solver with the GPU classes swapped in, or None when
    fluidsim itself is not installed (the GPU Simul works without it)."""
    try:
        import importlib

        parent = getattr(importlib.import_module(parent_module), parent_class)
    except Exception:  # fluidsim (or one of its binary dependencies) is absent
        return None

    class InfoSolverB200(parent):
        def _init_root(self):
            super()._init_root()
            self.module_name = module_name
            self.class_name = "Simul"
            self.short_name = short_name
            classes = self.classes
            classes.Operators.module_name = "fluidsim_b200.operators"
            classes.Operators.class_name = simul_class.Operators.__name__
            classes.State.module_name = "fluidsim_b200.state"
            classes.State.class_name = state_class
            classes.TimeStepping.module_name = "fluidsim_b200.time_stepping"
            classes.TimeStepping.class_name = "TimeSteppingPseudoSpectralB200"
            classes.Forcing.module_name = "fluidsim_b200.forcing"
            classes.Forcing.class_name = "ForcingB200"

    InfoSolverB200.__name__ = "InfoSolver" + simul_class.__name__[5:] + "B200"
    return InfoSolverB200
