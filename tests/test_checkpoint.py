"""Checkpoint interop: the minimal HDF5 writer / reader and the
``state_phys`` file layout of fluidsim (``fluidsim/util/output.py:47-160``,
``base/init_fields.py:140-298``).  CPU tests: format structure + round trips (no HDF5 library exists in
this image, so parity with libhdf5 is unpinned and these tests pin the specification rules instead)."""
import importlib.util
import os
import struct

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    # by path: importing the package would need the CUDA library to be loadable, these modules do not
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "fluidsim_b200", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


minihdf5 = _load("minihdf5")


def _tree():
    rng = np.random.default_rng(0)
    return {
        "@attrs": {"name_solver": "ns3d", "axes": np.array(["z", "y", "x"], dtype="|S9"), "n": 3, "x": 0.25},
        "state_phys": {
            "@attrs": {"what": "obj state_phys for fluidsim", "time": 0.125, "it": 7},
            "vx": rng.random((4, 6, 8)),
            "vy": rng.random((4, 6, 8)),
            "count": np.arange(12, dtype=np.int64).reshape(3, 4),
            "single": np.ones(5, dtype=np.float32),
        },
        "info_simul": {"params": {"@attrs": {"nu_2": 0.1, "none": None, "flag": True},
                                  "oper": {"@attrs": {"nx": 8, "Lx": 2 * np.pi, "type_fft": "fft3d.with_b200"}},
                                  "empty": {}}},
    }


def test_minihdf5_round_trip(tmp_path):
    path = tmp_path / "t.h5"
    tree = _tree()
    minihdf5.write_hdf5(path, tree)
    back = minihdf5.read_hdf5(path)
    for key in ("vx", "vy", "count", "single"):
        assert back["state_phys"][key].dtype == tree["state_phys"][key].dtype
        assert np.array_equal(back["state_phys"][key], tree["state_phys"][key])
    a = back["state_phys"]["@attrs"]
    assert a["time"] == 0.125 and a["it"] == 7 and a["what"] == b"obj state_phys for fluidsim"
    assert list(back["@attrs"]["axes"]) == [b"z", b"y", b"x"]
    pa = back["info_simul"]["params"]["@attrs"]
    assert pa["none"] == b"None" and pa["flag"] == 1 and pa["nu_2"] == 0.1
    assert back["info_simul"]["params"]["oper"]["@attrs"]["type_fft"] == b"fft3d.with_b200"
    assert back["info_simul"]["params"]["empty"] == {}


def test_minihdf5_structure_follows_the_specification(tmp_path):
    """Byte-level rules of the HDF5 file format specification for the structures written."""
    path = tmp_path / "t.h5"
    minihdf5.write_hdf5(path, _tree())
    d = path.read_bytes()
    # superblock version 0: signature, versions, offset / length sizes, K values, addresses
    assert d[:8] == b"\x89HDF\r\n\x1a\n"
    assert d[8:16] == bytes([0, 0, 0, 0, 0, 8, 8, 0])
    leaf_k, internal_k, flags = struct.unpack_from("<HHI", d, 16)
    assert (leaf_k, internal_k, flags) == (minihdf5.LEAF_K, minihdf5.INTERNAL_K, 0)
    base, free, eof, drv = struct.unpack_from("<QQQQ", d, 24)
    assert base == 0 and free == minihdf5.UNDEF and drv == minihdf5.UNDEF and eof == len(d)
    name_off, oh, cache, _ = struct.unpack_from("<QQII", d, 56)
    btree, heap = struct.unpack_from("<QQ", d, 80)
    assert name_off == 0 and cache == 1
    # root object header: version 1, messages 8-byte aligned, first message = symbol table (0x0011)
    version, _r, nmsg, refcount, size = struct.unpack_from("<BBHII", d, oh)
    assert version == 1 and refcount == 1 and oh % 8 == 0 and size % 8 == 0
    mtype, msize, _f = struct.unpack_from("<HHB", d, oh + 16)
    assert mtype == 0x0011 and msize == 16
    assert struct.unpack_from("<QQ", d, oh + 24) == (btree, heap)
    # B-tree node: one leaf, keys = heap offsets ("" first, largest name last), full-size node
    assert d[btree : btree + 4] == b"TREE" and d[btree + 4] == 0 and d[btree + 5] == 0
    (used,) = struct.unpack_from("<H", d, btree + 6)
    assert used == 1
    assert struct.unpack_from("<QQ", d, btree + 8) == (minihdf5.UNDEF, minihdf5.UNDEF)
    key0, snod, key1 = struct.unpack_from("<QQQ", d, btree + 24)
    assert key0 == 0 and d[snod : snod + 4] == b"SNOD" and d[snod + 4] == 1
    # local heap: free list closed by H5HL_FREE_NULL, names sorted in the symbol node
    assert d[heap : heap + 4] == b"HEAP"
    seg_size, free_off, seg = struct.unpack_from("<QQQ", d, heap + 8)
    assert struct.unpack_from("<QQ", d, seg + free_off) == (1, seg_size - free_off)
    (nsym,) = struct.unpack_from("<H", d, snod + 6)
    names = []
    for j in range(nsym):
        off, child_oh, ctype = struct.unpack_from("<QQI", d, snod + 8 + 40 * j)
        names.append(d[seg + off : d.index(b"\0", seg + off)])
        assert child_oh % 8 == 0 and ctype == 1  # both children of the root are groups
    assert names == sorted(names) == [b"info_simul", b"state_phys"]
    assert seg + key1 == seg + struct.unpack_from("<Q", d, snod + 8 + 40 * (nsym - 1))[0]


def test_minihdf5_float64_datatype_bytes():
    """The IEEE float64 / int64 little-endian datatype messages, byte for byte (format spec IV.A.2.d)."""
    f8 = minihdf5._datatype_message(np.float64)
    assert f8 == bytes.fromhex("11203f00" "08000000" "0000" "4000" "34" "0b" "00" "34" "ff030000")
    i8 = minihdf5._datatype_message(np.int64)
    assert i8 == bytes.fromhex("10080000" "08000000" "0000" "4000")
    s9 = minihdf5._datatype_message(np.dtype("|S9"))
    assert s9 == bytes.fromhex("13010000" "09000000")


def test_minihdf5_rejects_other_files(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not hdf5 at all")
    with pytest.raises(ValueError):
        minihdf5.read_hdf5(p)


# ---------------------------------------------------------------- state_phys files (fluidsim layout)
class _FakeState:
    """CPU stand-in for the GPU state container (same attribute surface as fluidsim_b200.state)."""

    def __init__(self, keys, shape):
        import torch

        from fluidsim_b200.setofvariables import SetOfVariables

        self.state_phys = SetOfVariables(keys=keys, shape_variable=shape, dtype=torch.float64, info="state_phys",
                                         value=0.0, device=torch.device("cpu"))
        self.calls = []

    def statespect_from_statephys(self):
        self.calls.append("spect_from_phys")

    def statephys_from_statespect(self):
        self.calls.append("phys_from_spect")


class _FakeSim:
    short_name = "ns3d.strat"
    ndim = 3
    forcing = None

    def __init__(self, shape=(4, 6, 8)):
        from fluidsim_b200.params import create_default_params

        self.params = create_default_params("ns3d.strat")
        self.params.oper.nz, self.params.oper.ny, self.params.oper.nx = shape
        self.oper = object()
        self.state = _FakeState(("vx", "vy", "vz", "b"), shape)
        self.time_stepping = type("TS", (), {"t": 0.0, "it": 0})()


def test_state_phys_file_layout_and_restart(tmp_path):
    """save -> file with the groups / attributes the reference's InitFieldsFromFile reads
    (base/init_fields.py:170-290) -> load into a second simulation."""
    import torch

    from fluidsim_b200 import checkpoint

    sim = _FakeSim()
    rng = np.random.default_rng(1)
    data = rng.standard_normal((4, 4, 6, 8))
    sim.state.state_phys.tensor.copy_(torch.from_numpy(data))
    sim.time_stepping.t, sim.time_stepping.it = 1.25, 17
    sim.params.nu_2 = 1e-3
    sim.params.forcing.key_forced = ["vt_fft", "vp_fft"]
    path = checkpoint.save_state_phys(sim, str(tmp_path), name_run="run0")
    assert os.path.basename(path) == "state_phys_t0001.250.h5"  # USE_T_END, t_end = 10 -> width 8
    root = minihdf5.read_hdf5(path)
    assert root["@attrs"]["name_solver"] == b"ns3d.strat" and root["@attrs"]["name_run"] == b"run0"
    assert [a.decode() for a in root["@attrs"]["axes"]] == ["z", "y", "x"]
    gs = root["state_phys"]
    assert gs["@attrs"]["what"] == b"obj state_phys for fluidsim"
    assert gs["@attrs"]["name_type_variables"] == b"state_phys"
    assert gs["@attrs"]["time"] == 1.25 and gs["@attrs"]["it"] == 17
    for i, key in enumerate(("vx", "vy", "vz", "b")):
        assert np.array_equal(gs[key], data[i])
    po = root["info_simul"]["params"]["oper"]["@attrs"]
    assert (po["nx"], po["ny"], po["nz"]) == (8, 6, 4) and po["Lx"] == 2 * np.pi
    pp = root["info_simul"]["params"]["@attrs"]
    assert pp["SAVE"] == 1 and pp["NEW_DIR_RESULTS"] == 1 and pp["nu_2"] == 1e-3 and pp["f"] == b"None"
    assert root["info_simul"]["solver"]["@attrs"]["short_name"] == b"ns3d.strat"
    # same `it` again: nothing new is written; another `it` at the same time: "_it=" suffix
    assert checkpoint.save_state_phys(sim, str(tmp_path)) == path
    sim.time_stepping.it = 18
    assert checkpoint.save_state_phys(sim, str(tmp_path)).endswith("state_phys_t001.250_it=18.h5")

    sim2 = _FakeSim()
    checkpoint.load_state_phys(sim2, path)
    assert np.array_equal(sim2.state.state_phys.tensor.numpy(), data)
    assert sim2.state.calls == ["spect_from_phys", "phys_from_spect"]
    assert sim2.time_stepping.t == 1.25 and sim2.time_stepping.it == 17
    # params tree round trip (None, lists of strings, nested containers)
    from fluidsim_b200.params import ParamContainer

    p2 = checkpoint.tree_to_params(root["info_simul"]["params"], ParamContainer())
    assert p2.f is None and p2.nu_2 == 1e-3 and p2.oper.nx == 8 and p2.forcing.key_forced == ["vt_fft", "vp_fft"]
    assert p2.time_stepping.phaseshift_random.nb_pairs == 1

    sim3 = _FakeSim(shape=(4, 6, 16))
    with pytest.raises(ValueError, match="not a correct state"):
        checkpoint.load_state_phys(sim3, path)
    sim4 = _FakeSim()
    sim4.params.oper.Ly = 3.0
    with pytest.raises(ValueError, match="params.oper.Ly"):
        checkpoint.load_state_phys(sim4, path)
    bad = tmp_path / "bad.h5"
    bad.write_bytes(b"garbage")
    with pytest.raises(ValueError, match="really a netCDF4/HDF5 file"):
        checkpoint.load_state_phys(sim2, str(bad))
