// Persistent, TMA-pipelined strided FFT pass (round 2).
//
// The round-1 strided kernel loads a tile with LDG, transforms it and stores it, one tile per CTA:
// ncu shows its warps waiting on the global loads (long_scoreboard 7-8 stalls per issue, DRAM
// 37-44 %).  Here a persistent CTA walks over tiles and the NEXT tile is fetched by the TMA unit
// (cp.async.bulk.tensor, 3-D tensor map over the (n0, n1, nk) complex array, one box of <= 256 rows
// x TK columns per instruction) into a shared-memory prefetch buffer while the current tile is
// transformed; an mbarrier (transaction bytes) hands the buffer over.  With dealias pruning only
// the kept rows of an inverse pass are fetched (boxes below and above the dealiased band), so the
// prefetch buffer is 2/3 of a tile and two CTAs fit one SM next to their exchange planes.
// The exchange plane is unpadded (XOR swizzle, fft_core.cuh PM = 1).
//
// Replaces the y / z parts of fluidfft's ifft_as_arg / fft_as_arg
// (/root/reference/fluidsim/solvers/ns3d/solver.py:210-241).
#pragma once
#include <cuda.h>

#include "passes.cuh"

#define B2_TMA_MAXBOX 8
struct TmaSet {
    CUtensorMap m[B2_MAXF];
};
struct TmaGeom {
    int axis_mid;   // 1: the FFT axis is the middle array axis (box over rows of axis 1), 0: outer axis
    int rb;         // rows per box
    int nbl, nbu;   // boxes covering rows [0, lo_rows) and [up_row0, up_row0 + up_rows)
    int up_row0;    // first row of the upper range (band_hi); unpruned: nbu = 0
    int pb_rows;    // rows of the prefetch buffer = (nbl + nbu) * rb
    int nct;        // column tiles
    int ntiles;     // nct * nouter * nf
};

B2_DEVINL unsigned st_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
B2_DEVINL void st_mbar_init(void* mbar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(st_smem_addr(mbar)), "r"(count) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
B2_DEVINL void st_mbar_expect_tx(void* mbar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(st_smem_addr(mbar)), "r"(bytes)
                 : "memory");
}
B2_DEVINL void st_mbar_wait(void* mbar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "ST_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra ST_DONE_%=;\n"
        "bra ST_WAIT_%=;\n"
        "ST_DONE_%=:\n"
        "}\n" ::"r"(st_smem_addr(mbar)),
        "r"(parity)
        : "memory");
}
B2_DEVINL void st_tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, void* mbar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
            "r"(st_smem_addr(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(st_smem_addr(mbar))
        : "memory");
}

struct PfSources {
    const cplx* src[B2_MAXF];
};
// InXf: value transform applied to the staged element (PlainIn: identity; Ns2dIn: ns2d prologue)
// TMA = true : the next tile is fetched by the TMA unit (tensor-map boxes), one elected thread issues
// TMA = false: the next tile is fetched with per-thread 16-byte cp.async (LDGSTS), all threads issue;
//              the prefetch buffer then holds exactly the kept rows (tg.rb = 1)
template <int N, int E, int TK, int DIR, bool TMA, class InXf, class StoreOp>
__global__ void __launch_bounds__(TK*(N / E))
    fft_strided_tma_kernel(const __grid_constant__ TmaSet maps, const __grid_constant__ PfSources srcs,
                           const __grid_constant__ Geom g, const __grid_constant__ TmaGeom tg,
                           const __grid_constant__ InXf in, const __grid_constant__ StoreOp st,
                           const cplx* __restrict__ tw) {
    extern __shared__ __align__(128) unsigned char b2_smem_raw[];
    constexpr int T = N / E;
    constexpr int PM = TK == 4 ? 1 : 0;
    cplx* plane = reinterpret_cast<cplx*>(b2_smem_raw);   // N x TK exchange plane
    cplx* pb = plane + (size_t)N * TK;                     // prefetch buffer, pb_rows x TK
    void* mbar = pb + (size_t)tg.pb_rows * TK;
    const int c = threadIdx.x % TK;
    const int t = threadIdx.x / TK;
    const unsigned box_bytes = (unsigned)tg.rb * TK * (unsigned)sizeof(cplx);
    auto issue = [&](int tile) {
        const int field = tile % g.nf;
        const int rest = tile / g.nf;
        const int ct = rest % tg.nct;
        const int oidx = rest / tg.nct + g.outer0;
        const int outer = oidx < g.outer_lo ? oidx : oidx + g.outer_gap;
        const int col0 = ct * TK * 2;  // in doubles
        st_mbar_expect_tx(mbar, box_bytes * (unsigned)(tg.nbl + tg.nbu));
        const CUtensorMap* mp = &maps.m[field];
        for (int b = 0; b < tg.nbl + tg.nbu; ++b) {
            const int row0 = b < tg.nbl ? b * tg.rb : tg.up_row0 + (b - tg.nbl) * tg.rb;
            cplx* dst = pb + (size_t)b * tg.rb * TK;
            if (tg.axis_mid) st_tma_load_3d(dst, mp, col0, row0, outer, mbar);
            else st_tma_load_3d(dst, mp, col0, outer, row0, mbar);
        }
    };
    // cp.async variant: thread i copies elements i, i + nthreads, ... of the compact (kept rows) tile
    auto issue_cpasync = [&](int tile) {
        const int field = tile % g.nf;
        const int rest = tile / g.nf;
        const int ct = rest % tg.nct;
        const int oidx = rest / tg.nct + g.outer0;
        const int outer = oidx < g.outer_lo ? oidx : oidx + g.outer_gap;
        const int cc = ct * TK + c;
        const cplx* src = srcs.src[field] + (long long)outer * g.os + cc;
        if (cc < g.dim_nk) {
            for (int rc = t; rc < tg.pb_rows; rc += T) {
                const int i = rc < tg.nbl ? rc : rc - (tg.nbl - tg.up_row0);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(st_smem_addr(pb + (size_t)rc * TK + c)),
                             "l"(src + (long long)i * g.es)
                             : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int tile = blockIdx.x;
    if constexpr (TMA) {
        if (threadIdx.x == 0) st_mbar_init(mbar, 1);
        __syncthreads();
        if (threadIdx.x == 0 && tile < tg.ntiles) issue(tile);
    } else {
        if (tile < tg.ntiles) issue_cpasync(tile);
    }
    unsigned ph = 0;
    const int up_shift = tg.nbl * tg.rb - tg.up_row0;  // compact row of an upper-range row i: i + up_shift
    for (; tile < tg.ntiles; tile += gridDim.x) {
        const int field = tile % g.nf;
        const int rest = tile / g.nf;
        const int ct = rest % tg.nct;
        const int oidx = rest / tg.nct + g.outer0;
        const int outer = oidx < g.outer_lo ? oidx : oidx + g.outer_gap;
        const int col = ct * TK + c;
        const bool active = col < g.ncols;
        const long long base = (long long)outer * g.os + col;
        if constexpr (TMA) {
            st_mbar_wait(mbar, ph);
            ph ^= 1u;
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
        }
        cplx x[E];
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int i = t + m * T;
            const bool zero = !active || (g.skip_load && i >= g.band_lo && i < g.band_hi);
            const int rc = (tg.nbu && i >= tg.up_row0) ? i + up_shift : i;
            x[m] = zero ? make_double2(0.0, 0.0) : in.xf(field, pb[(size_t)rc * TK + c], i, col, outer);
        }
        __syncthreads();  // prefetch buffer consumed
        if constexpr (TMA) {
            if (threadIdx.x == 0 && tile + (int)gridDim.x < tg.ntiles) issue(tile + (int)gridDim.x);
        } else {
            if (tile + (int)gridDim.x < tg.ntiles) issue_cpasync(tile + (int)gridDim.x);
        }
        fft_line<N, E, DIR, TK, 1, PM>(x, plane, t, c, tw, SyncBlock());
        if (active) {
#pragma unroll
            for (int m = 0; m < E; ++m) {
                const int i = t + m * T;
                if (!(g.skip_store && i >= g.band_lo && i < g.band_hi))
                    st(field, base + (long long)i * g.es, i, col, outer, x[m]);
            }
        }
    }
}
