"""Checkpoint / restart in fluidsim's ``state_phys`` file layout.

Mirror of ``PhysFieldsBase.save`` + ``save_file`` (``/root/reference/fluidsim/base/output/phys_fields.py:130-202``,
``fluidsim/util/output.py:47-160``) and of ``InitFieldsFromFile.__call__``
(``base/init_fields.py:140-298``):

    /                      attrs: "date saving", name_solver, name_run, axes (|S9 array)
    /state_phys            attrs: what, name_type_variables, time, it ; one dataset per state_phys key
    /info_simul/solver     attrs: module_name, class_name, short_name ; /classes/<Class> attrs
    /info_simul/params     the params tree: sub-containers = groups, parameters = attributes
                           (None stored as the string "None", fluiddyn's convention) + SAVE, NEW_DIR_RESULTS

The file is plain HDF5 (``.h5``), the flavour the reference writes with an MPI-enabled h5py and reads with
``h5py.File`` for every extension but ``.nc``.  It is written with h5py when that is importable, else by
``fluidsim_b200.minihdf5`` -- the case in this image, where h5py / libhdf5 do not exist (format parity of
that writer is unpinned, see the module).
The state travels device -> host (``state_phys`` is float64 in X space, as in the reference) only here.
"""

import datetime
import glob
import os

import numpy as np
import torch

from . import minihdf5

try:  # the reference's own container library, when the deployment has it
    import h5py
except ImportError:
    h5py = None


def _use_h5py():
    return h5py is not None and os.environ.get("B2_MINIHDF5", "0") in ("0", "")


def write_hdf5(path, root):
    """Write the nested-dict tree (model of ``minihdf5``) with h5py when it is importable -- then the
    file is libhdf5's own -- else with the built-in minimal writer."""
    if not _use_h5py():
        return minihdf5.write_hdf5(path, root)

    def put(group, node):
        for key, value in (node.get("@attrs") or {}).items():
            group.attrs[key] = "None" if value is None else value
        for name, child in node.items():
            if name == "@attrs":
                continue
            if isinstance(child, dict):
                put(group.create_group(name), child)
            else:
                group.create_dataset(name, data=child)

    with h5py.File(str(path), "w") as h5file:
        put(h5file, root)


def read_hdf5(path):
    if not _use_h5py():
        return minihdf5.read_hdf5(path)

    def get(group):
        node = {}
        attrs = dict(group.attrs.items())
        if attrs:
            node["@attrs"] = attrs
        for name, item in group.items():
            node[name] = get(item) if isinstance(item, h5py.Group) else item[...]
        return node

    with h5py.File(str(path), "r") as h5file:
        return get(h5file)


KEYS_PHYS_NEEDED = {
    "ns3d": ("vx", "vy", "vz"),  # solvers/ns3d/state.py:28
    "ns3d.strat": ("vx", "vy", "vz", "b"),  # solvers/ns3d/strat/state.py:26
    "ns3d.bouss": ("vx", "vy", "vz", "b"),
    "ns2d": ("rot",),  # solvers/ns2d/state.py:35
    "ns2d.strat": ("rot", "b"),  # solvers/ns2d/strat/state.py:45
    "ns2d.bouss": ("rot", "b"),
}


def _is_container(obj):
    return hasattr(obj, "_set_child") or hasattr(obj, "_tag_children")


def params_to_tree(container):
    """A params container (ours or fluiddyn's ParamContainer) as the nested-dict model of minihdf5."""
    if hasattr(container, "_key_attribs"):  # fluiddyn.util.paramcontainer.ParamContainer
        keys = list(container._key_attribs)
        children = list(container._tag_children)
    else:
        items = {k: v for k, v in vars(container).items() if not k.startswith("_")}
        keys = [k for k, v in items.items() if not _is_container(v)]
        children = [k for k, v in items.items() if _is_container(v)]
    attrs = {}
    for key in keys:
        value = getattr(container, key)
        if isinstance(value, (list, tuple)):
            if len(value) == 0:
                value = "[]"
            elif all(isinstance(v, str) for v in value):
                value = np.array([v.encode("utf-8") for v in value])
            else:
                value = np.asarray(value)
                if value.dtype.kind == "O":
                    value = repr(list(getattr(container, key)))
        elif isinstance(value, os.PathLike):
            value = os.fspath(value)
        elif value is not None and not isinstance(value, (bool, int, float, str, bytes, np.generic, np.ndarray)):
            value = repr(value)
        attrs[key] = value
    node = {"@attrs": attrs}
    for key in children:
        node[key] = params_to_tree(getattr(container, key))
    return node


def tree_to_params(node, container):
    """Inverse of ``params_to_tree`` onto an existing container (unknown keys are added)."""
    for key, value in (node.get("@attrs") or {}).items():
        if isinstance(value, bytes):
            value = value.decode("utf-8")
        if isinstance(value, str) and value == "None":
            value = None
        elif isinstance(value, np.ndarray) and value.dtype.kind in "SOU":
            value = [v.decode("utf-8") if isinstance(v, bytes) else str(v) for v in value.tolist()]
        elif isinstance(value, np.generic):
            value = value.item()
        setattr(container, key, value)
    for key, child in node.items():
        if key == "@attrs" or not isinstance(child, dict):
            continue
        sub = getattr(container, key, None)
        if sub is None or not _is_container(sub):
            sub = container._set_child(key)
            if sub is None:  # fluiddyn's _set_child returns None
                sub = getattr(container, key)
        tree_to_params(child, sub)
    return container


def _solver_info(sim):
    classes = {}
    for label, cls in (("Operators", type(sim.oper)), ("State", type(sim.state)),
                       ("TimeStepping", type(sim.time_stepping))):
        classes[label] = {"@attrs": {"module_name": cls.__module__, "class_name": cls.__name__}}
    if getattr(sim, "forcing", None) is not None:
        cls = type(sim.forcing)
        classes["Forcing"] = {"@attrs": {"module_name": cls.__module__, "class_name": cls.__name__}}
    return {
        "@attrs": {"module_name": type(sim).__module__, "class_name": type(sim).__name__,
                   "short_name": sim.short_name},
        "classes": classes,
    }


def file_name(sim, path_run):
    """phys_fields.py:138-166: ``state_phys_t{time:0{width}.3f}.h5``."""
    params = sim.params
    time = sim.time_stepping.t
    if params.time_stepping.USE_T_END:
        existing = sorted(glob.glob(os.path.join(path_run, "state_phys*")))
        if existing:
            str_width = len(os.path.basename(existing[0])[12:-3])
        else:
            str_width = int(np.log10(params.time_stepping.t_end)) + 7
    else:
        str_width = 7
    return f"state_phys_t{time:0{str_width}.3f}.h5"


def save_state_phys(sim, path_run=".", name_run="b200", particular_attr=None):
    """Write the current physical state; returns the path.  Same rule as the reference for a file that
    already exists (phys_fields.py:175-189): identical ``it`` -> nothing written, else ``_it=`` suffix."""
    os.makedirs(path_run, exist_ok=True)
    ts = sim.time_stepping
    path = os.path.join(path_run, file_name(sim, path_run))
    if os.path.exists(path):
        if int(read_hdf5(path)["state_phys"]["@attrs"]["it"]) == ts.it:
            return path
        path = os.path.join(path_run, f"state_phys_t{ts.t:07.3f}_it={ts.it}.h5")
    state_phys = sim.state.state_phys
    group_state = {"@attrs": {"what": "obj state_phys for fluidsim", "name_type_variables": state_phys.info,
                              "time": float(ts.t), "it": int(ts.it)}}
    for key in state_phys.keys:
        group_state[key] = state_phys.get_var(key).detach().cpu().numpy()
    axes = ("z", "y", "x") if sim.ndim == 3 else ("y", "x")  # operators3d.py:209, operators2d.py:117
    root_attrs = {"date saving": str(datetime.datetime.now()).encode(), "name_solver": sim.short_name,
                  "name_run": name_run, "axes": np.array(axes, dtype="|S9")}
    if particular_attr is not None:
        root_attrs["particular_attr"] = particular_attr
    params_tree = params_to_tree(sim.params)
    params_tree["@attrs"].update({"SAVE": 1, "NEW_DIR_RESULTS": 1})
    root = {"@attrs": root_attrs, "state_phys": group_state,
            "info_simul": {"solver": _solver_info(sim), "params": params_tree}}
    write_hdf5(path, root)
    return path


def load_state_phys(sim, path):
    """``InitFieldsFromFile.__call__`` (base/init_fields.py:153-298): check the grid of the file against
    ``sim.params.oper``, fill ``state_phys`` (missing needed keys -> 0), rebuild ``state_spect``, set
    ``time_stepping.t / it``."""
    try:
        root = read_hdf5(path)
    except Exception as exc:
        raise ValueError("Is file " + str(path) + " really a netCDF4/HDF5 file?") from exc
    try:
        oper_attrs = root["info_simul"]["params"]["oper"]["@attrs"]
    except KeyError:
        raise ValueError("The file " + str(path) + " does not contain a params object")
    if "state_phys" not in root:
        raise ValueError("The file " + str(path) + " does not contain a state_phys object")
    group_state = root["state_phys"]
    axes = root.get("@attrs", {}).get("axes", np.array([b"y", b"x"]))
    po = sim.params.oper
    for r in axes:
        r = r.decode("utf-8") if hasattr(r, "decode") else r
        if getattr(po, "n" + r) != oper_attrs["n" + r]:
            raise ValueError("this is not a correct state for this simulation\n" f"self.n{r} != params_file.n{r}")
        if "L" + r in oper_attrs and getattr(po, "L" + r) != oper_attrs["L" + r]:
            raise ValueError(
                "this is not a correct state for this simulation\n" f"self.params.oper.L{r} != params_file.L{r}"
            )
    state = sim.state
    state_phys = state.state_phys
    for key in KEYS_PHYS_NEEDED[sim.short_name]:
        if key in group_state:
            field = torch.from_numpy(np.ascontiguousarray(group_state[key], dtype=np.float64))
            state_phys.set_var(key, field.to(state_phys.tensor.device))
        else:
            state_phys.get_var(key).fill_(0.0)
    state.statespect_from_statephys()
    state.statephys_from_statespect()
    attrs = group_state.get("@attrs", {})
    sim.time_stepping.t = float(attrs["time"])
    sim.time_stepping.it = int(attrs.get("it", 0))
    return sim
