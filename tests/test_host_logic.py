"""CPU tests of host-side logic: slab band / split arithmetic, bench accounting, params defaults."""
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


class _FakeSlab:
    """SlabSimul._local_band without a GPU (the method only uses nyl / world / cyclic)."""

    def __init__(self, ny, world, cyclic):
        self.ny, self.world, self.cyclic, self.nyl = ny, world, cyclic, ny // world


@pytest.mark.parametrize("ny,world", [(16, 2), (16, 8), (1024, 4), (1024, 8), (64, 1)])
@pytest.mark.parametrize("cyclic", [False, True])
@pytest.mark.parametrize("band", [(6, 11), (342, 683), (0, 0), (3, 16)])
def test_local_band_matches_brute_force(ny, world, cyclic, band):
    from fluidsim_b200.slab import SlabSimul

    gy_lo, gy_hi = band
    if gy_hi > ny:
        pytest.skip("band outside the grid")
    if gy_lo == gy_hi:
        gy_lo = gy_hi = ny  # "no band" convention
    fake = _FakeSlab(ny, world, cyclic)
    total_kept = 0
    for r in range(world):
        lo, hi = SlabSimul._local_band(fake, r, gy_lo, gy_hi)
        rows = [yl * world + r if cyclic else r * fake.nyl + yl for yl in range(fake.nyl)]
        in_band = [gy_lo <= g < gy_hi for g in rows]
        # the local band must be exactly the set of local rows inside the global band
        expect = [lo <= yl < hi for yl in range(fake.nyl)]
        assert in_band == expect, (r, lo, hi)
        total_kept += fake.nyl - (hi - lo)
    assert total_kept == ny - (gy_hi - gy_lo)


def test_cyclic_distribution_balances_kept_rows():
    from fluidsim_b200.slab import SlabSimul

    ny, world, band = 1024, 8, (342, 683)
    kept = {}
    for cyclic in (False, True):
        fake = _FakeSlab(ny, world, cyclic)
        kept[cyclic] = [fake.nyl - (hi - lo) for lo, hi in (SlabSimul._local_band(fake, r, *band) for r in range(world))]
    assert max(kept[False]) == 128 and min(kept[False]) == 0  # blocks: middle ranks idle
    assert max(kept[True]) - min(kept[True]) <= 1            # cyclic: balanced


def test_exchange_layout_is_a_bijection_and_unpacks_to_natural_order():
    """exchange_index (K side, what SlabMapper computes) fills every slot of the send buffer once, and
    natural_from_exchanged (what RowMap::xoff does on the receiving side) restores (zc, ny, nk)."""
    from fluidsim_b200.slab import exchange_index, natural_from_exchanged

    world, nyl, nzl, nk, nchunks = 4, 3, 4, 5, 2
    ny, nz, zc = world * nyl, world * nzl, nzl // nchunks
    yl, z, kx = np.meshgrid(np.arange(nyl), np.arange(nz), np.arange(nk), indexing="ij")
    idx = exchange_index(z, kx, yl, nzl, nyl, nk, nchunks, ny)
    assert sorted(idx.reshape(-1).tolist()) == list(range(nyl * nz * nk))
    for cyclic in (False, True):
        # a received chunk holding, at [r][zl][yl][kx], the global row that slot stands for
        chunk = np.empty((world, zc, nyl, nk))
        for r in range(world):
            for y in range(nyl):
                chunk[r, :, y, :] = (y * world + r) if cyclic else (r * nyl + y)
        nat = natural_from_exchanged(chunk.reshape(-1), world, zc, nyl, nk, cyclic)
        assert nat.shape == (zc, ny, nk)
        assert np.array_equal(nat[0, :, 0], np.arange(ny))


def test_bench_traffic_accounting():
    b = _bench()
    full = b.class_passes("ns3d")
    assert full[:5] == [12.0, 12.0, 9.0, 6.0, 6.0]
    f = 2.0 / 3
    pr = b.class_passes("ns3d", f, f, f)
    assert abs(pr[2] - 9 * f) < 1e-12                      # x pass: R 6 fx + W 3 fx
    assert abs(pr[0] - (6 * f**3 + 6 * f**2)) < 1e-12      # z-inverse: R kept box, W all z of kept columns
    assert all(p_ <= q_ for p_, q_ in zip(pr, full))
    assert b.class_passes("ns2d")[1] == 0.0 and b.class_passes("ns2d")[4] == 0.0
    assert b.STEP_PASSES[("ns3d", "RK4")] == 195


def test_default_params_have_reference_names():
    """Attribute names consumed by the reference's hot path
    (base/time_stepping/base.py:33-94, operators3d.py:152-205, base/solvers/pseudo_spect.py:106-132)."""
    from fluidsim_b200.params import create_default_params

    p = create_default_params("ns3d.strat")
    for k in ("nx", "ny", "nz", "Lx", "Ly", "Lz", "coef_dealiasing", "type_fft", "truncation_shape", "NO_SHEAR_MODES"):
        assert hasattr(p.oper, k)
    for k in ("USE_T_END", "t_end", "it_end", "USE_CFL", "type_time_scheme", "deltat0", "deltat_max", "cfl_coef", "max_elapsed"):
        assert hasattr(p.time_stepping, k)
    for k in ("nu_2", "nu_4", "nu_8", "nu_m4", "f", "N", "no_vz_kz0", "projection"):
        assert hasattr(p, k)
    assert p.oper.coef_dealiasing == 2.0 / 3 and p.time_stepping.type_time_scheme == "RK4"
    p2 = create_default_params("ns2d")
    assert p2.oper.Lx == 8 and hasattr(p2, "beta")
    with pytest.raises(ValueError):
        create_default_params("sw1l")


def test_pyproject_entry_points_resolve():
    """Every entry point declared in pyproject.toml (fluidfft.plugins, fluidsim.solvers.*) names an
    importable module with the attribute the reference's loaders look up: ``FFTclass`` for fluidfft
    plugins (operators3d.py:229 through fluidfft.import_fft_class), ``Simul`` for solver modules
    (lib/fluidsim_core/loader.py:17-74)."""
    import importlib
    import os
    import tomllib

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "pyproject.toml"), "rb") as f:
        eps = tomllib.load(f)["project"]["entry-points"]
    assert set(eps) == {"fluidfft.plugins", "fluidsim.solvers.ns3d", "fluidsim.solvers.ns2d"}
    for name, target in eps["fluidfft.plugins"].items():
        mod = importlib.import_module(target)
        cls = mod.FFTclass
        layout = "get_dimX_K" if name.startswith("fft3d") else "get_is_transposed"  # 3-D / 2-D plugins
        for meth in ("fft_as_arg", "ifft_as_arg", "ifft_as_arg_destroy", "fft", "ifft", "get_shapeX_loc",
                     "get_shapeK_loc", "get_shapeK_seq", layout, "get_seq_indices_first_K",
                     "get_k_adim_loc", "sum_wavenumbers", "create_arrayX", "create_arrayK"):
            assert hasattr(cls, meth), (name, meth)
    for group in ("fluidsim.solvers.ns3d", "fluidsim.solvers.ns2d"):
        for name, target in eps[group].items():
            Simul = importlib.import_module(target).Simul
            p = Simul.create_default_params()
            assert p.time_stepping.type_time_scheme == "RK4"
            assert hasattr(Simul, "InfoSolver") and hasattr(Simul, "tendencies_nonlin")


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("cyclic", [False, True])
@pytest.mark.parametrize("nchunks", [1, 2])
def test_pruned_exchange_formulas_for_every_world_size(world, cyclic, nchunks):
    """Formula-level emulation of one pruned global transpose, K side -> z-slab side, for 2 / 4 / 8 ranks:
    the z-inverse pass stores through ``SlabMapper`` (csrc/strided.cu), the all-to-all moves the per-peer
    blocks of ``SlabSimul._exchange_plan``, the y-inverse pass loads through ``RowMap::xoff``
    (csrc/passes.cuh) with the per-rank bands of ``b2i_slab_local_band`` (csrc/api.cu, =
    ``SlabSimul._local_band``).  Restated here in numpy; every kept (ky, z, kx) must arrive where the
    y pass looks for it."""
    from fluidsim_b200.slab import SlabSimul

    ny, nz, nk, keepx = 32, 16, 9, 6
    gy_lo, gy_hi = 11, 22  # global dealiased ky band
    nyl, nzl = ny // world, nz // world
    if nzl % nchunks:
        pytest.skip("chunks do not divide the local z range")
    zc = nzl // nchunks
    pitch = keepx
    fake = _FakeSlab(ny, world, cyclic)
    bands = [SlabSimul._local_band(fake, r, gy_lo, gy_hi) for r in range(world)]
    nkr = [nyl - (hi - lo) for lo, hi in bands]
    cs_a = ny * zc * pitch        # chunk stride of a K-side send buffer (SlabMapper::cstride)
    cs_b = sum(nkr) * zc * pitch  # chunk stride of a z-slab-side receive buffer (cs_x in b2i_slab_ypass)
    blk = np.concatenate([[0], np.cumsum([n * zc * pitch for n in nkr])])  # RowMap::blk

    def ident(g, z, kx):  # unique value of mode (global ky row g, z, kx)
        return (g * nz + z) * nk + kx

    def grow(r, yl):  # global ky row of local row yl of rank r
        return yl * world + r if cyclic else r * nyl + yl

    # z-inverse store on every rank s (SlabMapper::operator())
    send = [np.full(nchunks * cs_a, -1.0) for _ in range(world)]
    for s in range(world):
        lo, hi = bands[s]
        gap = hi - lo
        for yl in range(nyl):
            if lo <= yl < hi:
                continue  # dealiased row: not visited
            ylc = yl if yl < lo else yl - gap
            for z in range(nz):
                q, zl = divmod(z, nzl)
                c, zlc = divmod(zl, zc)
                base = c * cs_a + ((q * zc + zlc) * nkr[s] + ylc) * pitch
                send[s][base:base + pitch] = ident(grow(s, yl), z, np.arange(pitch))
    # all-to-all per chunk: rank s -> peer q, equal blocks of nkr[s] * zc * pitch ("mine" / "theirs")
    recv = [np.full(nchunks * cs_b, -1.0) for _ in range(world)]
    for c in range(nchunks):
        for s in range(world):
            n = nkr[s] * zc * pitch
            for q in range(world):
                recv[q][c * cs_b + blk[s]: c * cs_b + blk[s] + n] = send[s][c * cs_a + q * n: c * cs_a + (q + 1) * n]
    # y-inverse load on every rank q (RowMap::xoff)
    d = world if cyclic else nyl
    for q in range(world):
        assert (recv[q] >= 0).all()  # every slot of the receive buffer was written
        for c in range(nchunks):
            for i in range(ny):
                if gy_lo <= i < gy_hi:
                    continue  # band rows are never loaded
                qq, m = divmod(i, d)
                r, yl = (m, qq) if cyclic else (qq, m)
                lo, hi = bands[r]
                ylc = yl if yl < lo else yl - (hi - lo)
                for zlc in range(zc):
                    off = c * cs_b + blk[r] + (zlc * nkr[r] + ylc) * pitch
                    z = q * nzl + c * zc + zlc
                    assert np.array_equal(recv[q][off:off + pitch], ident(i, z, np.arange(pitch))), (q, c, i, zlc)
