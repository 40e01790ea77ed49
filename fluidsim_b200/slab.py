"""Slab-decomposed (multi-GPU) time stepping: one process per GPU, ``torch.distributed`` for the
plumbing, the two global transposes of every 3-D FFT as NCCL all-to-alls between the CUDA phases.

Decomposition (the layout fluidsim already handles for
``fft3d.mpi_with_fftwmpi3d``, ``/root/reference/fluidsim/operators/operators3d.py:384-391``):

* physical / semi-spectral side: split along z -> local ``(ny, nz_loc, .)`` line sets
* spectral side: split along ky -> local K layout ``(ny_loc, nz, nx/2+1)``, ``dimX_K = (1, 0, 2)``

One RK stage = phase A (z-inverse written straight into the exchange layout: the "pack" is fused
into the FFT store), all-to-all, phase B (y-inverse straight from the received blocks into the natural ``(nz_loc, ny, nk)``
array: the "unpack" is fused into the FFT load; fused x pass; y-forward back into the exchange
layout), all-to-all, phase C (z-forward reading the exchange layout + RK epilogue).  The exchange
layout of a field is ``[z chunk][peer][z in chunk][ky_loc][kx]``: each all-to-all moves ``world``
contiguous blocks per field, and with z outside ky in a block the y passes walk rows ``nk`` apart.

The functions ``local_from_global`` / ``global_from_local`` / ``exchange_index`` define the layout
algebra and are exercised on CPU (gloo, world_size 2) in ``tests/test_slab_cpu.py``.
"""

import ctypes as C

import numpy as np

from ._lib import SCHEME_IDS, SOLVER_IDS


# ----------------------------------------------------------------------------------- layout algebra
def local_from_global(arr_global, rank, world, cyclic=False):
    """Global sequential K array ``(..., nz, ny, nk)`` -> this rank's ``(..., ny_loc, nz, nk)``.

    ``cyclic``: ky rows dealt round-robin (global row = yl * world + rank) instead of in blocks."""
    ny = arr_global.shape[-2]
    nyl = ny // world
    if cyclic:
        sl = arr_global[..., :, rank::world, :]
    else:
        sl = arr_global[..., :, rank * nyl:(rank + 1) * nyl, :]
    return np.ascontiguousarray(np.swapaxes(sl, -3, -2))


def global_from_local(parts, cyclic=False):
    """Inverse of ``local_from_global`` given the list of all ranks' local arrays."""
    sw = [np.swapaxes(p, -3, -2) for p in parts]  # (..., nz, ny_loc, nk)
    if not cyclic:
        return np.ascontiguousarray(np.concatenate(sw, axis=-2))
    world = len(sw)
    shape = list(sw[0].shape)
    shape[-2] *= world
    out = np.empty(shape, dtype=sw[0].dtype)
    for r, a in enumerate(sw):
        out[..., :, r::world, :] = a
    return out


def exchange_index(z, kx, yl, nzl, nyl, nk, nchunks=1, ny=None):
    """Offset of K-layout element (yl, z, kx) in the exchange layout
    [z chunk][peer][z in chunk][ky_loc][kx] (restated by ``SlabMapper`` in csrc/strided.cu; unpruned
    case: every local ky row and kx column is exchanged).  Chunk regions are ``ny * zc * nk``
    elements apart.  Inside a peer block z runs OUTSIDE ky, so that the receiving side's y passes
    walk rows ``nk`` apart (the single-GPU y-pass access pattern)."""
    r = z // nzl
    zl = z - r * nzl
    zc = nzl // nchunks
    c = zl // zc
    zlc = zl - c * zc
    cstride = 0 if nchunks == 1 else ny * zc * nk
    return c * cstride + ((r * zc + zlc) * nyl + yl) * nk + kx


def natural_from_exchanged(chunk, world, zc, nyl, nk, cyclic=False):
    """Received chunk (flat) ``[rank r][z in chunk][ky_loc of r][kx]`` -> natural ``(zc, ny, nk)`` (what
    the y-inverse pass does on the fly through ``RowMap::xoff`` in csrc/passes.cuh; unpruned case)."""
    b = np.asarray(chunk).reshape(world, zc, nyl, nk)
    if cyclic:  # global ky row = yl * world + r
        return np.ascontiguousarray(b.transpose(1, 2, 0, 3)).reshape(zc, nyl * world, nk)
    return np.ascontiguousarray(b.transpose(1, 0, 2, 3)).reshape(zc, nyl * world, nk)


def slab_layout(nx, ny, nz, rank, world, cyclic=False):
    """Layout description of one rank in the vocabulary of a fluidfft MPI FFT class (what fluidsim reads
    from ``oper_fft``: ``/root/reference/fluidsim/operators/operators3d.py:253-261,384-391``).

    X side: z-slabs ``(nz/P, ny, nx)``; K side: ky-slabs stored ``(ny/P, nz, nx/2+1)``, ``dimX_K = (1, 0, 2)``.
    With the block distribution the local ky rows are the contiguous block starting at
    ``seq_indices_first_K[0]`` (the fftwmpi3d layout); with the cyclic one they are ``rank::P`` and
    ``ky_indices_loc`` is the thing to use (``seq_indices_first_K`` then only names the first row)."""
    check_divisible(nz, ny, world)
    nyl, nzl, nk = ny // world, nz // world, nx // 2 + 1
    ky_idx = np.arange(rank, ny, world) if cyclic else np.arange(rank * nyl, (rank + 1) * nyl)

    def k_adim(n):
        k = np.fft.fftfreq(n, 1.0 / n)
        if n % 2 == 0:
            k[n // 2] = n // 2
        return k

    return dict(
        shapeX_seq=(nz, ny, nx), shapeX_loc=(nzl, ny, nx),
        shapeK_seq=(ny, nz, nk), shapeK_loc=(nyl, nz, nk),
        dimX_K=(1, 0, 2),
        seq_indices_first_X=(rank * nzl, 0, 0),
        seq_indices_first_K=(int(ky_idx[0]), 0, 0),
        ky_indices_loc=ky_idx,
        k_adim_loc=(k_adim(ny)[ky_idx], k_adim(nz), np.arange(nk, dtype=np.float64)),
    )


def check_divisible(nz, ny, world):
    if nz % world or ny % world:
        raise ValueError(f"slab decomposition needs nz={nz} and ny={ny} to be multiples of world={world}")


# ----------------------------------------------------------------------------------- GPU solver
class SlabSimul:
    """Distributed ns3d / ns3d.strat RK2 / RK4 stepping on a slab plan.

    ``params`` uses the reference's attribute names (see ``fluidsim_b200.params``).  The local
    state is ``state_spect`` of shape ``(nvar, ny_loc, nz, nk)``.
    """

    def __init__(self, solver, params, group=None, ky_distribution="auto", lean=None):
        import torch
        import torch.distributed as dist

        from ._lib import call, ptr

        if solver not in ("ns3d", "ns3d.strat"):
            raise NotImplementedError("slab decomposition is implemented for ns3d and ns3d.strat")
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (one process per GPU)")
        self.torch, self.dist = torch, dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.solver = solver
        self.params = params
        po = params.oper
        self.nx, self.ny, self.nz = int(po.nx), int(po.ny), int(po.nz)
        check_divisible(self.nz, self.ny, self.world)
        self.nyl, self.nzl = self.ny // self.world, self.nz // self.world
        self.nk = self.nx // 2 + 1
        self.device = torch.device("cuda", torch.cuda.current_device())
        import os

        ky_distribution = os.environ.get("B2_SLAB_KY", ky_distribution)
        if ky_distribution == "auto":
            # 2 ranks are balanced with blocks already (the dealiased band is centred); from 4 ranks
            # on, blocks leave the middle ranks with (almost) no kept rows
            ky_distribution = "cyclic" if self.world >= 4 else "block"
        if ky_distribution not in ("cyclic", "block"):
            raise ValueError("ky_distribution must be 'auto', 'cyclic' or 'block'")
        # "block" is the fftwmpi3d layout; "cyclic" deals the ky rows round-robin so that the kept
        # (non dealiased) rows -- hence the K-side work of the pruned transforms -- are balanced
        self.cyclic = ky_distribution == "cyclic"
        handle = C.c_void_p()
        call("b2_plan_create_slab", C.byref(handle), self.nz, self.ny, self.nx, float(po.Lz), float(po.Ly),
             float(po.Lx), self.rank, self.world, 1 if self.cyclic else 0)
        self.handle = handle
        self.shapeK_loc = (self.nyl, self.nz, self.nk)
        self.nvar = 4 if solver == "ns3d.strat" else 3
        self.nwork = self.nvar + 3
        self.nout = 3 if solver == "ns3d" else 6
        mk = lambda n: torch.zeros((n,) + self.shapeK_loc, dtype=torch.complex128, device=self.device)
        # Memory-lean mode (what makes 2048^3 fit 8 x 180 GB): exchange buffers sized for the PRUNED
        # exchange (allocated once the mask is known: ny * nz_loc * keepx and kept_ky * nz_loc * keepx
        # elements per field instead of a full K field each) and, for ns3d, the raw transform outputs
        # aliased with the stage buffer (work holds the 3 vorticity fields only).  Local K fields:
        # 27 -> 18.7 (ns3d).  The price: every state handed to the stepper must already be dealiased
        # (checked on the device), and tendencies_nonlin on arbitrary input is refused.
        if lean is None:
            lean = os.environ.get("B2_SLAB_LEAN", "0") not in ("0", "")
        self.lean = bool(lean)
        self.state_spect = mk(self.nvar)
        self._acc, self._stagebuf = mk(self.nvar), mk(self.nvar)
        self._alias = self.lean and solver == "ns3d"
        self._work = mk(3 if self._alias else self.nwork)
        if self.lean:
            self._xa = self._xb = None  # allocated by _alloc_exchange() when the mask is set
        else:
            self._xa, self._xb = mk(self.nwork), mk(self.nwork)
        self._xa_fs = self._xb_fs = int(np.prod(self.shapeK_loc))
        self.where_dealiased = None  # local uint8 mask (ny_loc, nz, nk); set_mask_from_global
        self.deltat = float(params.time_stepping.deltat0)
        self.scheme = params.time_stepping.type_time_scheme
        if self.scheme not in SCHEME_IDS:
            raise ValueError(f'Problem name time_scheme ("{self.scheme}")')
        self.it = 0
        self.t = 0.0
        self._scalar = torch.zeros(1, dtype=torch.float64, device=self.device)
        # dealias-pruned exchange (set up by set_mask...; used once the state is known dealiased)
        self.use_pruning = True
        self.use_cfl = bool(getattr(params.time_stepping, "USE_CFL", False))
        self.pipelined = True  # overlap the per-(field, chunk) all-to-alls with the FFT passes
        # z chunks of the exchange layout (pipelining granularity)
        import os

        want = int(os.environ.get("B2_SLAB_NCHUNK", "2"))
        while want > 1 and (self.nzl % want or self.nzl // want < 8):
            want //= 2
        self.nchunks = max(1, want)
        self._state_dealiased = False
        self._prune = None
        self._push()
        # native collectives: the all-to-alls are issued by the library itself (NCCL communicator owned
        # by the plan, one C call per time step).  B2_SLAB_NATIVE=0 keeps the torch.distributed
        # all_to_all_single orchestration below (also what the gloo CPU tests exercise).
        self.native = False
        if os.environ.get("B2_SLAB_NATIVE", "1") not in ("0", "") and dist.get_backend(group) == "nccl":
            self._init_native_comm()
        self._dt_dev = self._vmax_dev = None

    # ---- layout surface of a fluidfft MPI FFT class (operators3d.py:253-261) ------------------------
    def layout(self):
        return slab_layout(self.nx, self.ny, self.nz, self.rank, self.world, self.cyclic)

    def get_shapeX_loc(self):
        return self.layout()["shapeX_loc"]

    def get_shapeX_seq(self):
        return self.layout()["shapeX_seq"]

    def get_shapeK_loc(self):
        return self.layout()["shapeK_loc"]

    def get_shapeK_seq(self):
        return self.layout()["shapeK_seq"]

    def get_dimX_K(self):
        return self.layout()["dimX_K"]

    def get_seq_indices_first_X(self):
        return self.layout()["seq_indices_first_X"]

    def get_seq_indices_first_K(self):
        return self.layout()["seq_indices_first_K"]

    def get_k_adim_loc(self):
        return self.layout()["k_adim_loc"]

    def _init_native_comm(self):
        from ._lib import call, ptr

        tr = self.torch
        uid = tr.zeros(128, dtype=tr.uint8)
        if self.rank == 0:
            buf = (C.c_char * 128)()
            call("b2_nccl_unique_id", buf)
            uid = tr.frombuffer(bytearray(buf.raw), dtype=tr.uint8).clone()
        uid = uid.to(self.device)
        src = self.dist.get_global_rank(self.group, 0) if self.group is not None else 0
        self.dist.broadcast(uid, src=src, group=self.group)
        raw = bytes(uid.cpu().numpy().tobytes())
        call("b2_slab_comm_init", self.handle, C.c_char_p(raw))
        self.native = True

    def _push(self):
        from ._lib import call, ptr

        p = self.params
        f = getattr(p, "f", None)
        call("b2_set_physics", self.handle, SOLVER_IDS[self.solver], float(p.nu_2), float(p.nu_4),
             float(p.nu_8), float(p.nu_m4), 0 if f is None else 1, 0.0 if f is None else float(f),
             float(getattr(p, "N", 0.0)), 0.0, ptr(self.where_dealiased))
        call("b2_set_buffers", self.handle, ptr(self._acc), ptr(self._stagebuf), ptr(self._work))
        call("b2_set_aliasing", self.handle, 1 if self._alias else 0)
        if self._xa is not None:
            call("b2_slab_set_buffers", self.handle, ptr(self._xa), ptr(self._xb))
            call("b2_slab_set_buffer_strides", self.handle, self._xa_fs, self._xb_fs)

    # ---- data in / out ---------------------------------------------------------------------------
    def set_mask_from_global(self, mask_global):
        """``where_dealiased`` of the sequential operator ``(nz, ny, nk)`` -> local slab."""
        loc = local_from_global(np.asarray(mask_global, dtype=np.uint8), self.rank, self.world, self.cyclic)
        self.set_local_mask(self.torch.from_numpy(loc).to(self.device))

    def set_local_mask(self, mask_local):
        self.where_dealiased = mask_local.contiguous()
        self._push()
        self._setup_pruning()
        if self.lean:
            self._alloc_exchange()

    def _alloc_exchange(self):
        """Lean mode: exchange buffers of the pruned exchange only."""
        from ._lib import call, ptr

        if self._prune is None:
            raise ValueError("lean slab buffers need a dealiasing mask (pruned exchange)")
        keepx, kz_lo, kz_hi, yl_lo, yl_hi, gy_lo, gy_hi = self._prune["args"]
        nyk = self.ny - (gy_hi - gy_lo)
        self._xa_fs = self.ny * self.nzl * keepx
        self._xb_fs = nyk * self.nzl * keepx
        tr = self.torch
        self._xa = self._xb = None
        self._xa = tr.zeros(self.nwork * self._xa_fs, dtype=tr.complex128, device=self.device)
        self._xb = tr.zeros(self.nwork * self._xb_fs, dtype=tr.complex128, device=self.device)
        call("b2_slab_set_buffers", self.handle, ptr(self._xa), ptr(self._xb))
        call("b2_slab_set_buffer_strides", self.handle, self._xa_fs, self._xb_fs)

    def _field(self, buf, f, stride):
        """Field f of an exchange buffer (lean buffers are flat, the others (nwork, ny_loc, nz, nk))."""
        return buf[f] if buf.dim() > 1 else buf[f * stride:(f + 1) * stride]

    def set_state_from_global(self, state_global):
        loc = local_from_global(np.asarray(state_global), self.rank, self.world, self.cyclic)
        self.state_spect.copy_(self.torch.from_numpy(loc))
        self._state_dealiased = False

    def mark_spect_modified(self):
        self._state_dealiased = False

    # ---- pruning set-up -----------------------------------------------------------------------------
    @staticmethod
    def _band(keep):
        """[lo, hi): the contiguous run of non-kept indices around n/2 (``kept_range`` in api.cu).

        Only that run is pruned; other fully dealiased indices (ky = 0 with NO_KY0, custom masks)
        stay in the visited set, where the per-mode mask zeroes them."""
        keep = np.asarray(keep, dtype=bool)
        n = len(keep)
        c = n // 2
        if c >= n or keep[c]:
            best, bl, i = 0, n, 0
            while i < n:
                if keep[i]:
                    i += 1
                    continue
                j = i
                while j < n and not keep[j]:
                    j += 1
                if j - i > best:
                    best, bl = j - i, i
                i = j
            return (n, n) if best == 0 else (bl, bl + best)
        lo, hi = c, c + 1
        while lo > 0 and not keep[lo - 1]:
            lo -= 1
        while hi < n and not keep[hi]:
            hi += 1
        return lo, hi

    def _setup_pruning(self):
        """Agree on the kept ranges between the ranks and derive the all-to-all split sizes."""
        tr, dist = self.torch, self.dist
        if self.where_dealiased is None:
            self._prune = None
            return
        kept = self.where_dealiased == 0  # (ny_loc, nz, nk)
        kx = kept.any(dim=0).any(dim=0).to(tr.int32)
        kz = kept.any(dim=2).any(dim=0).to(tr.int32)
        kyl = kept.any(dim=2).any(dim=1).to(tr.int32)
        dist.all_reduce(kx, op=dist.ReduceOp.MAX, group=self.group)
        dist.all_reduce(kz, op=dist.ReduceOp.MAX, group=self.group)
        parts = [tr.empty_like(kyl) for _ in range(self.world)]
        dist.all_gather(parts, kyl, group=self.group)
        kx, kz, kyl = kx.cpu().numpy(), kz.cpu().numpy(), kyl.cpu().numpy()
        kyg = np.empty(self.ny, dtype=kyl.dtype)
        for r, part in enumerate(parts):
            if self.cyclic:
                kyg[r::self.world] = part.cpu().numpy()
            else:
                kyg[r * self.nyl:(r + 1) * self.nyl] = part.cpu().numpy()
        keepx = int(np.nonzero(kx)[0].max()) + 1 if kx.any() else 1
        kz_lo, kz_hi = self._band(kz)
        gy_lo, gy_hi = self._band(kyg)
        yl_lo, yl_hi = self._local_band(self.rank, gy_lo, gy_hi)
        self._prune = dict(args=(keepx, kz_lo, kz_hi, yl_lo, yl_hi, gy_lo, gy_hi))

    def _local_band(self, r, gy_lo, gy_hi):
        """Dealiased band of rank r's local ky rows (mirrors ``b2i_slab_local_band``)."""
        nyl, P = self.nyl, self.world
        clamp = lambda v: max(0, min(nyl, v))
        if self.cyclic:
            cd = lambda a: 0 if a <= 0 else (a + P - 1) // P
            lo, hi = clamp(cd(gy_lo - r)), clamp(cd(gy_hi - r))
        else:
            lo, hi = clamp(gy_lo - r * nyl), clamp(gy_hi - r * nyl)
        if hi <= lo:
            lo = hi = nyl
        return lo, hi

    def _kept_rows(self):
        """Kept local ky rows per rank for the pruning state last pushed to the library."""
        from ._lib import call

        arr = (C.c_int * self.world)()
        call("b2_slab_kept_rows", self.handle, arr)
        return list(arr)

    def gather_state(self):
        """Global sequential-layout state on every rank (testing / small grids only)."""
        parts = [self.torch.empty_like(self.state_spect) for _ in range(self.world)]
        self.dist.all_gather(parts, self.state_spect, group=self.group)
        return global_from_local([p.cpu().numpy() for p in parts], self.cyclic)

    # ---- collectives ------------------------------------------------------------------------------
    def _exchange_plan(self, pr):
        """Split sizes / chunk strides (in float64 elements) of the per-(field, chunk) all-to-alls."""
        P, zc = self.world, self.nzl // self.nchunks
        pitch = self.nk if pr is None else pr["args"][0]
        nkl = self._kept_rows()
        row = zc * pitch * 2
        mine = [nkl[self.rank] * row] * P      # K side: equal blocks, one per peer (its z range)
        theirs = [n * row for n in nkl]        # z-slab side: peer r contributes its kept ky rows
        return dict(cs_a=self.ny * row, cs_b=sum(nkl) * row, mine=mine, theirs=theirs)

    def _a2a(self, src, dst, src_off, dst_off, ins, outs, async_op):
        """One all-to-all: contiguous per-peer blocks (NCCL over NVLink; gloo on CPU tests)."""
        tr = self.torch
        d, s_ = tr.view_as_real(dst).view(-1), tr.view_as_real(src).view(-1)
        return self.dist.all_to_all_single(d[dst_off: dst_off + sum(outs)], s_[src_off: src_off + sum(ins)],
                                           outs, ins, group=self.group, async_op=async_op)

    # ---- stepping ---------------------------------------------------------------------------------
    def _run_stage(self, Sin, need_curl, scheme_id, stage, tout=None, prune=False):
        """One evaluation of the nonlinear term + RK epilogue.

        K side (z passes, epilogue) and z-slab side (y passes, fused x pass) are separated by two
        global transposes.  Pipelining: the exchange layout is cut in z chunks; the inverse
        all-to-all of (field f, chunk c) is issued right after the z-inverse of field f, the y / x
        passes of chunk c run while chunk c+1 is still in flight, and the forward all-to-all of
        chunk c overlaps the passes of chunk c+1 (NCCL stream vs. compute stream)."""
        from ._lib import call, ptr, stream_ptr

        h = self.handle
        pr = self._prune if (prune and self._prune is not None) else None
        if pr is None and self.lean:
            raise ValueError("lean slab buffers hold the pruned exchange only: the input must be dealiased "
                             "(SlabSimul(..., lean=False) handles arbitrary states)")
        if pr is None:
            call("b2_slab_set_pruning", h, 0, 0, 0, 0, 0, 0, 0, 0)
        else:
            call("b2_slab_set_pruning", h, 1, *pr["args"])
        nc = self.nchunks
        call("b2_slab_set_chunks", h, nc)
        ex = self._exchange_plan(pr)
        sp = stream_ptr()
        XA = lambda f: self._field(self._xa, f, self._xa_fs)
        XB = lambda f: self._field(self._xb, f, self._xb_fs)
        asyn = self.pipelined
        order = list(range(self.nwork))
        if need_curl:  # v (and b) first: they do not depend on the curl kernel
            order = [0, 1, 2] + list(range(6, self.nwork)) + [3, 4, 5]
        inv = {}
        curl_done = not need_curl
        for f in order:
            if not curl_done and 3 <= f < 6:
                call("b2_slab_curl", h, ptr(Sin), sp)
                curl_done = True
            call("b2_slab_zinv", h, ptr(Sin), f, f + 1, sp)
            inv[(f, 0)] = self._a2a(XA(f), XB(f), 0, 0, ex["mine"], ex["theirs"], asyn)
        for c in range(1, nc):
            for f in order:
                inv[(f, c)] = self._a2a(XA(f), XB(f), c * ex["cs_a"], c * ex["cs_b"], ex["mine"], ex["theirs"], asyn)
        fwd = {}
        for c in range(nc):
            for f in order:
                if asyn:
                    inv[(f, c)].wait()
                call("b2_slab_yinv", h, f, f + 1, c, sp)
            call("b2_slab_xpass", h, c, sp)
            for f in range(self.nout):
                call("b2_slab_yfwd", h, f, f + 1, c, sp)
                fwd[(f, c)] = self._a2a(XB(f), XA(f), c * ex["cs_b"], c * ex["cs_a"], ex["theirs"], ex["mine"], asyn)
        for f in range(self.nout):
            if asyn:
                for c in range(nc):
                    fwd[(f, c)].wait()
            call("b2_slab_zfwd", h, f, f + 1, sp)
        call("b2_slab_rk", h, scheme_id, stage, self.deltat, ptr(Sin), ptr(self.state_spect), ptr(tout), sp)

    def _set_stage_layout(self, prune):
        from ._lib import call

        pr = self._prune if (prune and self._prune is not None) else None
        if pr is None and self.lean:
            raise ValueError("lean slab buffers hold the pruned exchange only: the input must be dealiased "
                             "(SlabSimul(..., lean=False) handles arbitrary states)")
        if pr is None:
            call("b2_slab_set_pruning", self.handle, 0, 0, 0, 0, 0, 0, 0, 0)
        else:
            call("b2_slab_set_pruning", self.handle, 1, *pr["args"])
        call("b2_slab_set_chunks", self.handle, self.nchunks)

    def _native_step(self, sid, prune):
        """One C call per time step: FFT passes, NCCL all-to-alls and epilogues inside the library."""
        from ._lib import call, ptr, stream_ptr

        self._set_stage_layout(prune)
        tr = self.torch
        if self.use_cfl:
            # CFL on the device with the GLOBAL max |v| (base/time_stepping/base.py:320-354)
            if self._dt_dev is None:
                self._dt_dev = tr.full((1,), float(self.deltat), dtype=tr.float64, device=self.device)
                self._vmax_dev = tr.zeros(3, dtype=tr.float64, device=self.device)
            else:
                self._dt_dev.fill_(float(self.deltat))
            ts = self.params.time_stepping
            cfl = ts.cfl_coef if getattr(ts, "cfl_coef", None) else (1.0 if self.scheme == "RK4" else 0.4)
            call("b2_slab_time_step_cfl", self.handle, sid, float(cfl), float(ts.deltat_max), ptr(self._dt_dev),
                 ptr(self._vmax_dev), ptr(self.state_spect), stream_ptr())
            self.deltat = float(self._dt_dev.item())
        else:
            call("b2_slab_time_step", self.handle, sid, float(self.deltat), ptr(self.state_spect), stream_ptr())

    def tendencies_nonlin(self, state_spect=None, old=None):
        if self.lean:
            raise ValueError("tendencies_nonlin on arbitrary input needs the full-size buffers (lean=False)")
        src = self.state_spect if state_spect is None else state_spect
        out = self.torch.empty_like(self.state_spect) if old is None else old
        if self.native:
            from ._lib import call, ptr, stream_ptr

            self._set_stage_layout(False)
            call("b2_slab_tendencies", self.handle, ptr(src), ptr(out), stream_ptr())
        else:
            self._run_stage(src, True, SCHEME_IDS[self.scheme], -1, out)
        return out

    def one_time_step(self):
        sid = SCHEME_IDS[self.scheme]
        nstages = 4 if self.scheme == "RK4" else 2
        if self.use_pruning and not self._state_dealiased and self.where_dealiased is not None:
            self._state_dealiased = self._check_dealiased()
        prune = self.use_pruning and self._state_dealiased
        if self.native:
            self._native_step(sid, prune)
        else:
            if bool(getattr(self.params.time_stepping, "USE_CFL", False)) and self.use_cfl:
                raise NotImplementedError("CFL on slab plans needs the native collectives (NCCL)")
            for st in range(nstages):
                Sin = self.state_spect if st == 0 else self._stagebuf
                self._run_stage(Sin, st == 0, sid, st, prune=prune)
        self._state_dealiased = True  # the last stage projects and dealiases the state
        self.t += self.deltat
        self.it += 1

    def _check_dealiased(self):
        """Is the local state exactly zero wherever the mask dealiases, on every rank?"""
        from ._lib import call, ptr, stream_ptr

        tr = self.torch
        flag = tr.zeros(1, dtype=tr.int32, device=self.device)
        call("b2_check_dealiased", self.handle, ptr(self.state_spect), self.nvar, ptr(self.where_dealiased),
             ptr(flag), stream_ptr())
        self.dist.all_reduce(flag, op=self.dist.ReduceOp.MAX, group=self.group)
        return int(flag.item()) == 0

    def compute_observables(self):
        """Spatial means + 1-D / 3-D spectra of the velocity in one kernel pass per rank
        (``b2_observables``) and one all-reduce (SUM); same dictionary as
        ``OperatorsPseudoSpectral3D.compute_observables``."""
        import math

        from ._lib import call, lib, ptr, stream_ptr

        po = self.params.oper
        dks = [2 * math.pi / float(po.Lx), 2 * math.pi / float(po.Ly), 2 * math.pi / float(po.Lz)]
        deltak = max(dks)
        nks = int(math.sqrt((dks[0] * (self.nx // 2)) ** 2 + (dks[1] * (self.ny // 2)) ** 2
                            + (dks[2] * (self.nz // 2)) ** 2) / deltak) + 2
        n = int(lib.b2_observables_size(self.handle, 3, nks))
        out = self.torch.empty(n, dtype=self.torch.float64, device=self.device)
        call("b2_observables", self.handle, ptr(self.state_spect), 3, nks, deltak, ptr(out), stream_ptr())
        self.dist.all_reduce(out, group=self.group)
        o = out.cpu().numpy()
        res = {"Ex": float(o[0]), "Ey": float(o[1]), "Ez": float(o[2]), "E": float(o[0] + o[1] + o[2]),
               "epsK": float(o[4]), "epsK_hypo": float(o[5]), "epsK4": float(o[6]), "epsK8": float(o[7]),
               "enstrophy": float(o[8])}
        pos = 16
        for nm, ln in (("", nks), ("_kx", self.nx // 2 + 1), ("_ky", self.ny // 2 + 1), ("_kz", self.nz // 2 + 1)):
            for v in ("vx", "vy", "vz"):
                res[v + nm] = o[pos:pos + ln].copy()
                pos += ln
        res["E_spectrum3d"] = res["vx"] + res["vy"] + res["vz"]
        return res

    def compute_energy(self):
        """sum_wavenumbers(|v|^2)/2 over the velocity components, all-reduced."""
        from ._lib import call, ptr, stream_ptr

        call("b2_sum_wavenumbers_abs2", self.handle, ptr(self.state_spect), 3, ptr(self._scalar), stream_ptr())
        self.dist.all_reduce(self._scalar, group=self.group)
        return 0.5 * float(self._scalar.item())

    def __del__(self):
        try:
            from ._lib import lib

            if getattr(self, "handle", None):
                lib.b2_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass
