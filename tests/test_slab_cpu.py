"""CPU tests (gloo, world_size 2) of the slab layout algebra used by the multi-GPU path: a numpy
model of phases A/B/C with ``torch.distributed.all_to_all_single`` in between must reproduce the
sequential 3-D FFT, and scatter/gather must round-trip."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model_inverse_fft(k_loc, rank, world, nz, ny, nx, nchunks=1, cyclic=False):
    """Distributed unnormalised inverse FFT of one field following the GPU phases' layouts."""
    from fluidsim_b200.slab import exchange_index, natural_from_exchanged

    nyl, nzl, nk = ny // world, nz // world, nx // 2 + 1
    zc = nzl // nchunks
    # phase A: z-inverse on (ny_loc, nz, nk), stored in the exchange layout
    a = np.fft.ifft(k_loc, axis=1) * nz
    send = np.empty(nyl * nz * nk, dtype=np.complex128)
    yl, z, kx = np.meshgrid(np.arange(nyl), np.arange(nz), np.arange(nk), indexing="ij")
    send[exchange_index(z, kx, yl, nzl, nyl, nk, nchunks, ny)] = a
    recv = np.empty_like(send)
    cs = ny * zc * nk  # chunk stride (complex elements): one all-to-all per chunk
    out = np.empty((ny, nzl, nx))
    for c in range(nchunks):
        dist.all_to_all_single(torch.view_as_real(torch.from_numpy(recv[c * cs:(c + 1) * cs])).view(-1),
                               torch.view_as_real(torch.from_numpy(send[c * cs:(c + 1) * cs])).view(-1))
        # phase B: the received chunk is [rank][z in chunk][ky_loc][kx]; the y-inverse reads it through
        # the row map and works on the natural (zc, ny, nk) array; then c2r along x
        b = natural_from_exchanged(recv[c * cs:(c + 1) * cs], world, zc, nyl, nk, cyclic)
        b = np.fft.ifft(b, axis=1) * ny
        out[:, c * zc:(c + 1) * zc] = np.swapaxes(np.fft.irfft(b, n=nx, axis=2) * nx, 0, 1)
    return out  # (ny, nz_loc, nx)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fluidsim_b200.slab import global_from_local, local_from_global

        nz, ny, nx = 8, 12, 10
        rng = np.random.default_rng(0)
        x = rng.random((nz, ny, nx))
        kg = np.fft.rfftn(x) / x.size  # sequential layout (nz, ny, nk)
        nzl = nz // world
        ref = np.swapaxes(x[rank * nzl:(rank + 1) * nzl], 0, 1)  # (ny, nz_loc, nx)
        err = 0.0
        for cyclic in (False, True):
            k_loc = local_from_global(kg, rank, world, cyclic)
            assert k_loc.shape == (ny // world, nz, nx // 2 + 1)
            # scatter / gather round trip
            parts = [None] * world
            dist.all_gather_object(parts, k_loc)
            assert np.array_equal(global_from_local(parts, cyclic), kg)
            # distributed inverse transform == this rank's z-slab of the sequential inverse
            for nchunks in (1, 2):
                got = _model_inverse_fft(k_loc, rank, world, nz, ny, nx, nchunks, cyclic)
                err = max(err, np.abs(got - ref).max())
        q.put((rank, float(err)))
    finally:
        dist.destroy_process_group()


def test_slab_layout_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(2))
    assert set(res) == {0, 1}
    assert max(res.values()) < 1e-13


def test_divisibility_error():
    from fluidsim_b200.slab import check_divisible

    with pytest.raises(ValueError):
        check_divisible(10, 12, 4)


def test_slab_layout_surface_matches_the_layout_algebra():
    """slab_layout (the fluidfft MPI-class vocabulary: shapes, dimX_K, first indices, local k) agrees with
    local_from_global for block and cyclic ky distributions."""
    from fluidsim_b200.slab import global_from_local, local_from_global, slab_layout

    nx, ny, nz, world = 8, 12, 6, 3
    nk = nx // 2 + 1
    ky = np.fft.fftfreq(ny, 1.0 / ny)
    ky[ny // 2] = ny // 2
    KY = np.broadcast_to(ky[None, :, None], (nz, ny, nk))  # sequential K array (nz, ny, nk)
    for cyclic in (False, True):
        parts = []
        for rank in range(world):
            lay = slab_layout(nx, ny, nz, rank, world, cyclic)
            loc = local_from_global(KY, rank, world, cyclic)
            assert loc.shape == lay["shapeK_loc"] == (ny // world, nz, nk)
            assert lay["shapeX_loc"] == (nz // world, ny, nx) and lay["dimX_K"] == (1, 0, 2)
            assert lay["seq_indices_first_X"] == (rank * (nz // world), 0, 0)
            assert np.array_equal(loc[:, 0, 0], lay["k_adim_loc"][0])
            assert lay["seq_indices_first_K"][0] == lay["ky_indices_loc"][0]
            if not cyclic:  # fftwmpi3d blocks (operators3d.py:384-391)
                assert lay["seq_indices_first_K"] == (rank * (ny // world), 0, 0)
            parts.append(loc)
        assert np.array_equal(global_from_local(parts, cyclic), KY)
    with pytest.raises(ValueError):
        slab_layout(8, 10, 6, 0, 4)
