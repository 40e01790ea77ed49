// Paired contiguous-axis pass (round 2): the fused c2r -> physical-space product -> r2c kernel
// rebuilt around COMPLEX transforms of TWO real fields each.
//
//   * inverse: the spectral lines of two real fields a, b (e.g. v_x and omega_x) are combined into
//     the Hermitian-extended spectrum of a + i b, ONE complex FFT of length N = nx gives both
//     physical lines (real part a, imaginary part b).  No half-length pre-processing pass, no
//     pre-processing twiddles.
//   * one thread group owns a whole line: after the three paired inverse transforms every thread
//     holds all six physical fields at the same 16 points in registers, so the products are formed
//     in registers (the previous kernel parked every field in shared memory and read it back).
//   * forward: two real product fields are paired again (f_x + i f_y); the unpaired one (f_z, ns2d's
//     F_rot) is paired with the same field of the NEXT line, which the group processes right after
//     (thread-private shared-memory slot in between).  Z(k) and Z(N-k) are separated into the two
//     spectra on the way out.
//   * every thread works in every phase; groups (one CTA of T = N/16 threads at N >= 1024) are
//     independent, so co-resident CTAs sit in different phases and overlap their global loads,
//     FP64 butterflies and shared-memory exchanges.
//
// Per line (ns3d) this is 4.5 complex length-N transforms instead of 9 half-length ones plus
// pre/post passes: ~15 % fewer FP64 instructions and ~35 % fewer shared-memory wavefronts, the two
// pipes that bounded the round-1 kernel (profiles/r1_final_bench_summary.md).
//
// Replaces: ifft_as_arg_destroy x6, vector_product, fft_as_arg x3
// (/root/reference/fluidsim/solvers/ns3d/solver.py:210-241), div_vb_fft_from_vb's products
// (solvers/ns3d/strat/solver.py:204-206), compute_Frot (solvers/ns2d/solver.py:34-38).
#pragma once
#include <type_traits>

#include "fft_core.cuh"
#include "xpair_op.h"

// ------------------------------------------------------------------------------- building blocks
// Z[j] = A[j] + i B[j] on the Hermitian-extended index j = t + m T in [0, N):
//   j <= N/2 : A[j] = a[j]            j > N/2 : A[j] = conj(a[N - j])
// Modes k >= nkeep are dealiased zeros (not stored, not loaded).  The imaginary parts of k = 0 and
// k = N/2 are ignored (c2r convention of FFTW / fluidfft).
template <int N, int E>
B2_DEVINL void xp_load_pair(cplx (&x)[E], const cplx* a, const cplx* b, int t, int nkeep) {
    constexpr int T = N / E, H = N / 2;
#pragma unroll
    for (int m = 0; m < E; ++m) {
        const int j = t + m * T;
        const bool mir = m >= E / 2;  // j >= N/2
        const int k = mir ? N - j : j;
        cplx av = make_double2(0.0, 0.0), bv = make_double2(0.0, 0.0);
        if (k < nkeep) {
            av = a[k];
            bv = b[k];
        }
        if (k == 0 || k == H) {
            av.y = 0.0;
            bv.y = 0.0;
        }
        x[m] = mir ? make_double2(av.x + bv.y, bv.x - av.y) : make_double2(av.x - bv.y, av.y + bv.x);
    }
}

// x[m] = Z[t + m T], Z = FFT(f1 + i f2):  F1[k] = (Z[k] + conj Z[N-k]) / 2, F2[k] = (Z[k] - conj Z[N-k]) / 2i
// for k < nkeep, scaled by 2 hs.  The mirrored half travels through the group's exchange plane.
template <int N, int E, class Sync>
B2_DEVINL void xp_separate_store(const cplx (&x)[E], cplx* plane, int t, Sync sync, cplx* o1, cplx* o2,
                                 bool st1, bool st2, double hs, int nkeep) {
    constexpr int T = N / E, H = N / 2;
    sync();  // WAR: the last exchange of the transform has been read by every thread of the group
#pragma unroll
    for (int m = E / 2; m < E; ++m) {
        const int j = t + m * T;
        if (j > H && N - j < nkeep) plane[j - H] = x[m];
    }
    sync();
#pragma unroll
    for (int m = 0; m <= E / 2; ++m) {
        const int k = t + m * T;
        if ((m < E / 2 || t == 0) && k < nkeep) {
            const cplx z = x[m];
            const cplx c = (k == 0 || k == H) ? z : plane[H - k];
            if (st1) o1[k] = make_double2((z.x + c.x) * hs, (z.y - c.y) * hs);
            if (st2) o2[k] = make_double2((z.y + c.y) * hs, (c.x - z.x) * hs);
        }
    }
}

// ------------------------------------------------------------------------------- bulk-async staging
// The two spectral lines of the NEXT transform are fetched with cp.async.bulk (TMA, 1-D) into a
// per-group shared-memory buffer while the current transform runs; completion is tracked by an
// mbarrier (transaction bytes).  The threads then combine the pair from shared memory, so the global
// load latency never sits on the critical path of a group (ncu: long_scoreboard was the top stall).
B2_DEVINL unsigned xp_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
B2_DEVINL void xp_mbar_init(void* mbar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(xp_smem_addr(mbar)), "r"(count) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
B2_DEVINL void xp_mbar_expect_tx(void* mbar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(xp_smem_addr(mbar)), "r"(bytes)
                 : "memory");
}
B2_DEVINL void xp_mbar_wait(void* mbar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "XP_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra XP_DONE_%=;\n"
        "bra XP_WAIT_%=;\n"
        "XP_DONE_%=:\n"
        "}\n" ::"r"(xp_smem_addr(mbar)),
        "r"(parity)
        : "memory");
}
B2_DEVINL void xp_bulk_g2s(void* dst, const void* src, unsigned bytes, void* mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     xp_smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(xp_smem_addr(mbar))
                 : "memory");
}

// xp_load_pair from a staged copy: a[k] at st[k], b[k] at st[nkeep + k], k < nkeep
template <int N, int E>
B2_DEVINL void xp_combine_staged(cplx (&x)[E], const cplx* st, int t, int nkeep, bool live) {
    constexpr int T = N / E, H = N / 2;
#pragma unroll
    for (int m = 0; m < E; ++m) {
        const int j = t + m * T;
        const bool mir = m >= E / 2;
        const int k = mir ? N - j : j;
        cplx av = make_double2(0.0, 0.0), bv = make_double2(0.0, 0.0);
        if (live && k < nkeep) {
            av = st[k];
            bv = st[nkeep + k];
        }
        if (k == 0 || k == H) {
            av.y = 0.0;
            bv.y = 0.0;
        }
        x[m] = mir ? make_double2(av.x + bv.y, bv.x - av.y) : make_double2(av.x - bv.y, av.y + bv.x);
    }
}

// max |u| of the real parts (velocity components) of a transformed pair -> atomicMax(dst)
template <int E, int W>
B2_DEVINL void xp_absmax_re(const cplx (&x)[E], double* dst) {
    double mx = 0.0;
#pragma unroll
    for (int m = 0; m < E; ++m) mx = fmax(mx, fabs(x[m].x));
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o, W));
    if ((threadIdx.x % W) == 0)
        atomicMax(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)__double_as_longlong(mx));
}

// ------------------------------------------------------------------------------- product ops
// KIND 0: ns3d   in = vx vy vz wx wy wz          out = (v x w)_xyz                   (W[0..2])
// KIND 1: strat  in = vx vy vz wx wy wz b        out = (v x w)_xyz, vx b, vy b, vz b (W[0..5])
// KIND 2: ns2d   in = ux uy d_x rot d_y rot      out = -u . grad rot - beta uy       (W[0])
template <int KIND>
struct XPTraits {
    // thread-private shared-memory parking per group, in doubles per line point
    static constexpr int PARK_D = KIND == 1 ? 2 : 1;
};

// One group of T = N/E threads processes pairs of consecutive lines.  G groups per CTA (G > 1 only
// when T <= 32: groups are then parts of one warp and synchronise with __syncwarp()).
template <int N, int E, int G, int KIND, bool VMAX, bool PARK0, int MAXREG, bool PF>
__global__ void __launch_bounds__(G*(N / E)) __maxnreg__(MAXREG)
    xpass_pair_kernel(PairOp op, long long nlines, const cplx* __restrict__ tw, double scale, int nkeep,
                      int pitch, long long line0, int ppg) {
    extern __shared__ double b2_smem[];
    constexpr int T = N / E, PS = PlaneSize<N, 1>::value;
    constexpr int PARK_D = XPTraits<KIND>::PARK_D + (PARK0 ? 2 : 0);
    constexpr int IPP = KIND == 0 ? 6 : (KIND == 1 ? 7 : 4);  // staged transforms per pair of lines
    // doubles per group: [mbarrier (2)] exchange plane, thread-private parking, [staging 2 x nkeep cplx]
    const int per_g = (PF ? 2 : 0) + PS * 2 + PARK_D * N + (PF ? 4 * nkeep : 0);
    static_assert(G == 1 || T <= 32, "several groups per CTA only when a group is part of one warp");
    using Sync = typename std::conditional<(T > 32), SyncBlock, SyncWarp>::type;
    constexpr int W = T < 32 ? T : 32;
    const Sync sync{};
    const int g = threadIdx.x / T, t = threadIdx.x % T;
    double* gbase = b2_smem + (size_t)g * per_g;
    void* mbar = gbase;
    cplx* plane = reinterpret_cast<cplx*>(gbase + (PF ? 2 : 0));
    double* park = gbase + (PF ? 2 : 0) + PS * 2;                // [PARK_D][E][T] thread-private slots
    cplx* park0 = reinterpret_cast<cplx*>(park + XPTraits<KIND>::PARK_D * N);  // PARK0: first pair
    cplx* stage = reinterpret_cast<cplx*>(park + PARK_D * N);    // PF: staged (a, b) lines
    const long long npairs = (nlines + 1) / 2;
    const long long pair0 = ((long long)blockIdx.x * G + g) * ppg;
    long long pair = pair0;
    if ((long long)blockIdx.x * G * ppg >= npairs) return;
    const double hs = 0.5 * scale;
    // staged item q of this group: transform k = q % IPP of pair pair0 + q / IPP
    auto issue = [&](int q) {
        const int it = q / IPP, k = q - it * IPP;
        long long pr = pair0 + it;
        if (pr >= npairs) pr = npairs - 1;
        const long long la = 2 * pr, lb = (2 * pr + 1 < nlines) ? 2 * pr + 1 : la;
        const cplx *a, *b;
        if (KIND == 1 && k == 0) {
            a = op.in[6] + (la + line0) * pitch;
            b = op.in[6] + (lb + line0) * pitch;
        } else {
            const int kk = KIND == 1 ? k - 1 : k;
            constexpr int NP = KIND == 2 ? 2 : 3;
            const int h = kk / NP, pp = kk - h * NP;
            const long long off = ((h ? lb : la) + line0) * pitch;
            a = op.in[pp] + off;
            b = op.in[pp + NP] + off;
        }
        const unsigned bytes = (unsigned)nkeep * (unsigned)sizeof(cplx);
        xp_mbar_expect_tx(mbar, 2 * bytes);
        xp_bulk_g2s(stage, a, bytes, mbar);
        xp_bulk_g2s(stage + nkeep, b, bytes, mbar);
    };
    int q = 0;
    unsigned ph = 0;
    if constexpr (PF) {
        if (t == 0) xp_mbar_init(mbar, 1);
        sync();
        if (t == 0) issue(0);
    }
    // next (a + i b) spectrum of the sequence into x: staged copy (PF) or direct global loads
    auto fetch = [&](cplx (&x)[E], const cplx* a, const cplx* b, bool live) {
        if constexpr (PF) {
            xp_mbar_wait(mbar, ph);
            ph ^= 1u;
            xp_combine_staged<N, E>(x, stage, t, nkeep, live);
            sync();  // the staging buffer has been read by the whole group
            ++q;
            if (t == 0 && q < ppg * IPP) issue(q);
        } else {
            xp_load_pair<N, E>(x, a, b, t, live ? nkeep : 0);
        }
    };
#pragma unroll 1
    for (int it = 0; it < ppg; ++it, ++pair) {
        const bool pv = pair < npairs;                      // invalid groups redo the last pair, stores off
        const long long pr = pv ? pair : npairs - 1;
        const long long la = 2 * pr, lb_raw = 2 * pr + 1;
        const bool bv = lb_raw < nlines;
        const long long lb = bv ? lb_raw : la;
        const long long offa = (la + line0) * pitch, offb = (lb + line0) * pitch;
        if constexpr (KIND == 1) {
            // buoyancy of both lines in one transform, parked (re: line a, im: line b)
            cplx xb[E];
            fetch(xb, op.in[6] + offa, op.in[6] + offb, pv);
            fft_line<N, E, +1, 1, 1>(xb, plane, t, 0, tw, sync);
            cplx* bp = reinterpret_cast<cplx*>(park);
#pragma unroll
            for (int m = 0; m < E; ++m) bp[m * T + t] = xb[m];
        }
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const long long off = h ? offb : offa;
            const bool lv = pv && (h == 0 || bv);
            cplx P0[E], P1[E];
            if constexpr (KIND == 2) {
                fetch(P0, op.in[0] + off, op.in[2] + off, lv);  // (ux, d_x rot)
                fft_line<N, E, +1, 1, 1>(P0, plane, t, 0, tw, sync);
                if constexpr (VMAX) xp_absmax_re<E, W>(P0, op.vmax + 0);
                fetch(P1, op.in[1] + off, op.in[3] + off, lv);  // (uy, d_y rot)
                fft_line<N, E, +1, 1, 1>(P1, plane, t, 0, tw, sync);
                if constexpr (VMAX) xp_absmax_re<E, W>(P1, op.vmax + 1);
                const double beta = op.beta;
#pragma unroll
                for (int m = 0; m < E; ++m) {
                    const double f = beta == 0.0 ? -P0[m].x * P0[m].y - P1[m].x * P1[m].y
                                                 : -P0[m].x * P0[m].y - P1[m].x * (P1[m].y + beta);
                    if (h == 0) park[m * T + t] = f;
                    else P0[m] = make_double2(park[m * T + t], f);
                }
                if (h == 1) {
                    fft_line<N, E, -1, 1, 1>(P0, plane, t, 0, tw, sync);
                    xp_separate_store<N, E>(P0, plane, t, sync, op.out[0] + offa, op.out[0] + offb, pv, pv && bv,
                                            hs, nkeep);
                }
            } else {
                cplx P2[E];
                fetch(P0, op.in[0] + off, op.in[3] + off, lv);  // (vx, wx)
                fft_line<N, E, +1, 1, 1>(P0, plane, t, 0, tw, sync);
                if constexpr (VMAX) xp_absmax_re<E, W>(P0, op.vmax + 0);
                if constexpr (PARK0) {
#pragma unroll
                    for (int m = 0; m < E; ++m) park0[m * T + t] = P0[m];
                }
                fetch(P1, op.in[1] + off, op.in[4] + off, lv);  // (vy, wy)
                fft_line<N, E, +1, 1, 1>(P1, plane, t, 0, tw, sync);
                if constexpr (VMAX) xp_absmax_re<E, W>(P1, op.vmax + 1);
                fetch(P2, op.in[2] + off, op.in[5] + off, lv);  // (vz, wz)
                fft_line<N, E, +1, 1, 1>(P2, plane, t, 0, tw, sync);
                if constexpr (VMAX) xp_absmax_re<E, W>(P2, op.vmax + 2);
                if constexpr (PARK0) {
#pragma unroll
                    for (int m = 0; m < E; ++m) P0[m] = park0[m * T + t];
                }
#pragma unroll
                for (int m = 0; m < E; ++m) {
                    const double vx = P0[m].x, wx = P0[m].y, vy = P1[m].x, wy = P1[m].y, vz = P2[m].x,
                                 wz = P2[m].y;
                    const double fx = vy * wz - vz * wy;
                    const double fy = vz * wx - vx * wz;
                    const double fz = vx * wy - vy * wx;
                    P0[m] = make_double2(fx, fy);
                    if constexpr (KIND == 1) {
                        const cplx bb = reinterpret_cast<const cplx*>(park)[m * T + t];
                        const double b = h ? bb.y : bb.x;
                        P1[m] = make_double2(fz, vx * b);
                        P2[m] = make_double2(vy * b, vz * b);
                    } else {
                        if (h == 0) park[m * T + t] = fz;
                        else P1[m] = make_double2(park[m * T + t], fz);
                    }
                }
                fft_line<N, E, -1, 1, 1>(P0, plane, t, 0, tw, sync);
                xp_separate_store<N, E>(P0, plane, t, sync, op.out[0] + off, op.out[1] + off, lv, lv, hs, nkeep);
                if constexpr (KIND == 1) {
                    fft_line<N, E, -1, 1, 1>(P1, plane, t, 0, tw, sync);
                    xp_separate_store<N, E>(P1, plane, t, sync, op.out[2] + off, op.out[3] + off, lv, lv, hs,
                                            nkeep);
                    fft_line<N, E, -1, 1, 1>(P2, plane, t, 0, tw, sync);
                    xp_separate_store<N, E>(P2, plane, t, sync, op.out[4] + off, op.out[5] + off, lv, lv, hs,
                                            nkeep);
                } else if (h == 1) {
                    fft_line<N, E, -1, 1, 1>(P1, plane, t, 0, tw, sync);
                    xp_separate_store<N, E>(P1, plane, t, sync, op.out[2] + offa, op.out[2] + offb, pv, pv && bv,
                                            hs, nkeep);
                }
            }
        }
    }
}
