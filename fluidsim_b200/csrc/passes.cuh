// FFT pass kernels (float64 / complex128) built on fft_core.cuh.
//
//  * fft_strided_kernel   batched complex FFT along a strided axis (y or z pass).  A CTA owns a
//                         tile of TK adjacent columns (contiguous in memory -> every global access
//                         is a TK*16-byte segment) and the whole line along the FFT axis.
//                         LoadOp / StoreOp functors fuse k-space prologues / epilogues.
//  * xpass_*_kernel       contiguous-axis pass: c2r, r2c and the fused
//                         c2r x NI -> pointwise product -> r2c x NO kernel.
//  * fft_generic_kernel   any line length (mixed radix, O(N * sum(prime factors)) in shared
//                         memory) -- the non power-of-two path used by the operator-level API.
//
// Replaces fluidfft's fft_as_arg / ifft_as_arg / vector_product call chain
// (/root/reference/fluidsim/solvers/ns3d/solver.py:199-241).
#pragma once
#include "fft_core.cuh"

#define B2_MAXR 8  // ranks of a slab decomposition (one node)

// Row map of the slab y passes.  The exchanged (z-slab side) array of one z chunk is grouped by owning
// rank: [rank r][z in chunk][kept local ky row of r][kx] -- rank r's block holds only its kept rows
// (its local dealiased band [lo, lo + gap) is not stored) and starts blk[r] elements into the chunk.
// Logical ky row i belongs to rank r = i / nyl (block) or i % P (cyclic ky distribution).
struct RowMap {
    int P;  // 0: identity
    int nyl, cyclic;
    int shift;  // log2 of the divisor (P if cyclic, nyl otherwise) when it is a power of two, else -1
    int lo[B2_MAXR], gap[B2_MAXR], nkr[B2_MAXR];
    long long blk[B2_MAXR];
    // element offset (without the kx column) of logical row i, plane z of the chunk.  32-bit
    // arithmetic inside a rank block (a chunk of one field stays far below 2^31 elements).
    B2_DEVINL long long xoff(int i, int z, int pitch) const {
        const int d = cyclic ? P : nyl;
        const int q = shift >= 0 ? (i >> shift) : i / d;
        const int m = i - q * d;
        const int r = cyclic ? m : q;
        const int yl = cyclic ? q : m;
        const int ylc = yl < lo[r] ? yl : yl - gap[r];
        return blk[r] + (long long)((unsigned)(z * nkr[r] + ylc) * (unsigned)pitch);
    }
};

struct Geom {
    int ncols;      // number of columns (contiguous index)
    int nouter;     // number of outer batches
    long long es;   // stride (elements) between successive points of a line
    long long os;   // stride (elements) between outer batches
    long long cs;   // stride between columns (1 for the fast path)
    int nf;         // number of fields handled by the launch
    // dealias-pruned transforms: the kept (non dealiased) wavenumbers of an axis are the index
    // ranges [0, lo) and [lo + gap, n).  `outer` runs over the kept outer indices only and is
    // mapped to memory by o -> o < outer_lo ? o : o + outer_gap.  Along the FFT axis the rows
    // i in [band_lo, band_hi) are known zeros: they are not loaded (skip_load, inverse passes) or
    // not stored (skip_store, forward passes).
    int outer_lo, outer_gap;
    int band_lo, band_hi;
    int skip_load, skip_store;
    int wide;       // prefer the wide-tile configuration (plane-strided lines)
    // row storage of the source (map_load) / destination (map_store) array: rows.operator()(i)
    // (band rows are never touched)
    RowMap rows;
    int map_load, map_store;
    int xpitch;     // kx pitch of the exchanged array (row map)
    int outer0;     // first outer index of this launch (chunked launches)
    // shape of the (n0, n1, nk) complex array the pass works on (TMA tensor maps, strided_tma.cuh);
    // dim_nk == 0: unknown -> LDG kernels only
    int dim_nk, dim_n1, dim_n0;
    // L2 prefetch distance in CTAs (0: off): every CTA first asks L2 for the tile of the CTA l2pf
    // positions ahead in launch order (prefetch.global.L2: no registers, no shared memory), so that
    // the demand loads of that CTA hit L2 instead of paying the DRAM latency
    int l2pf;
};
static inline Geom geom_init() {
    Geom g;
    g.ncols = 0; g.nouter = 1; g.es = 0; g.os = 0; g.cs = 1; g.nf = 1;
    g.outer_lo = 1 << 30; g.outer_gap = 0; g.band_lo = 0; g.band_hi = 0; g.skip_load = 0; g.skip_store = 0;
    g.wide = 0;
    g.rows.P = 0; g.rows.nyl = 1; g.rows.cyclic = 0; g.rows.shift = -1; g.map_load = 0; g.map_store = 0;
    g.xpitch = 0;
    for (int r = 0; r < B2_MAXR; ++r) { g.rows.blk[r] = 0; g.rows.lo[r] = 0; g.rows.gap[r] = 0; g.rows.nkr[r] = 0; }
    g.outer0 = 0;
    g.dim_nk = 0; g.dim_n1 = 0; g.dim_n0 = 0;
    g.l2pf = 0;
    return g;
}

#define B2_MAXF 8
// ------------------------------------------------------------------------------- load/store ops
B2_DEVINL void b2_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
struct PlainLoad {
    const cplx* in[B2_MAXF];
    B2_DEVINL cplx operator()(int f, long long off, int i, int col, int outer) const {
        return in[f][off];
    }
    B2_DEVINL void prefetch(int f, long long off, int i, int col, int outer) const { b2_prefetch_l2(in[f] + off); }
};
struct PlainStore {
    cplx* out[B2_MAXF];
    B2_DEVINL void operator()(int f, long long off, int i, int col, int outer, cplx v) const {
        out[f][off] = v;
    }
};
struct ScaleStore {
    cplx* out[B2_MAXF];
    double scale;
    B2_DEVINL void operator()(int f, long long off, int i, int col, int outer, cplx v) const {
        out[f][off] = cscale(v, scale);
    }
};

// ------------------------------------------------------------------------------- strided pass
// ROWMAP: compile the RowMap row translation in (slab y passes only; it costs ~10 % when merely
// tested at run time in the single-GPU passes).
template <int N, int E, int TK, int DIR, bool ROWMAP, class LoadOp, class StoreOp>
__global__ void __launch_bounds__(TK*(N / E), (TK * (N / E) <= 256 ? 2 : 1))
    fft_strided_kernel(const __grid_constant__ Geom g, const __grid_constant__ LoadOp ld,
                       const __grid_constant__ StoreOp st, const cplx* __restrict__ tw) {
    extern __shared__ double b2_smem[];
    constexpr int T = N / E;
    cplx* plane = reinterpret_cast<cplx*>(b2_smem);
    const int c = threadIdx.x % TK;
    const int t = threadIdx.x / TK;
    // field is the fastest-varying block coordinate: the CTAs that (re-)read the same input tile
    // for different output fields (curl prologue) run at the same time and share it through L2
    const int field = blockIdx.x % g.nf;
    const int col = (blockIdx.x / g.nf) * TK + c;
    const int oidx = (int)blockIdx.y + g.outer0;
    const int outer = oidx < g.outer_lo ? oidx : oidx + g.outer_gap;
    const bool active = col < g.ncols;
    const long long base = (long long)outer * g.os + col;
    cplx x[E];
#pragma unroll
    for (int m = 0; m < E; ++m) {
        const int i = t + m * T;
        const bool zero = !active || (g.skip_load && i >= g.band_lo && i < g.band_hi);
        long long addr = base + (long long)i * g.es;
        if constexpr (ROWMAP) {
            if (g.map_load && !zero) addr = g.rows.xoff(i, outer, g.xpitch) + col;
        }
        x[m] = zero ? make_double2(0.0, 0.0) : ld(field, addr, i, col, outer);
    }
    if (g.l2pf > 0) {
        // tile of the CTA g.l2pf launches ahead (x fastest, then y)
        const long long lin = (long long)blockIdx.y * gridDim.x + blockIdx.x + g.l2pf;
        const int by = (int)(lin / gridDim.x), bx = (int)(lin - (long long)by * gridDim.x);
        if (by < (int)gridDim.y) {
            const int pfield = bx % g.nf;
            const int pcol = (bx / g.nf) * TK + c;
            const int poidx = by + g.outer0;
            const int pouter = poidx < g.outer_lo ? poidx : poidx + g.outer_gap;
            const long long pbase = (long long)pouter * g.os + pcol;
            // one request per 128-byte line: lanes c with (c * 16) % 128 == 0
            if (pcol < g.ncols && (c % (TK < 8 ? TK : 8)) == 0) {
#pragma unroll
                for (int m = 0; m < E; ++m) {
                    const int i = t + m * T;
                    if (!(g.skip_load && i >= g.band_lo && i < g.band_hi)) {
                        long long paddr = pbase + (long long)i * g.es;
                        if constexpr (ROWMAP) {
                            if (g.map_load) paddr = g.rows.xoff(i, pouter, g.xpitch) + pcol;
                        }
                        ld.prefetch(pfield, paddr, i, pcol, pouter);
                    }
                }
            }
        }
    }
    fft_line<N, E, DIR, TK, 1>(x, plane, t, c, tw, SyncBlock());
    if (active) {
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int i = t + m * T;
            if (!(g.skip_store && i >= g.band_lo && i < g.band_hi)) {
                long long addr = base + (long long)i * g.es;
                if constexpr (ROWMAP) {
                    if (g.map_store) addr = g.rows.xoff(i, outer, g.xpitch) + col;
                }
                st(field, addr, i, col, outer, x[m]);
            }
        }
    }
}


// ------------------------------------------------------------------------------- input ops
// InOp interface (shared by the fast and the generic kernels):
//     const cplx* ptr(int field, long long off, int i, int col, int outer) const  -- source address
//     cplx xf(int field, cplx v, int i, int col, int outer) const                 -- value transform
struct PlainIn {
    const cplx* in[B2_MAXF];
    B2_DEVINL const cplx* ptr(int f, long long off, int i, int col, int outer) const { return in[f] + off; }
    B2_DEVINL cplx xf(int f, cplx v, int i, int col, int outer) const { return v; }
};

// ------------------------------------------------------------------------------- x pass pieces
// ---- experimental (off by default, knob B2_XTWC): pre/post-processing twiddles from ONE table entry
// per thread: exp(-2 pi i (t + m T)/N) = tw[t] * exp(-i pi m / E), the second factor being a
// compile-time constant (multiples of pi/16).  Measured: no gain (profiles/r1_tuning.md).
constexpr double b2_cos16_ce(int j) {
    return j == 0 ? 1.0 : j == 1 ? 0.98078528040323044913 : j == 2 ? 0.92387953251128675613
         : j == 3 ? 0.83146961230254523708 : j == 4 ? 0.70710678118654752440
         : j == 5 ? 0.55557023301960222474 : j == 6 ? 0.38268343236508977173
         : j == 7 ? 0.19509032201612826785 : j == 8 ? 0.0 : -b2_cos16_ce(16 - j);
}
template <int J>
struct B2C16 {  // constant-evaluated: usable as immediates in device code
    static constexpr double c = b2_cos16_ce(J);
    static constexpr double s = J <= 8 ? b2_cos16_ce(8 - J) : b2_cos16_ce(J - 8);
};
template <int N, int E, int m>
B2_DEVINL void c2r_pre_twc(cplx (&x)[E], const cplx* __restrict__ K, int t, cplx wt, int nkeep) {
    if constexpr (m < E) {
        constexpr int M = N / 2, T = M / E;
        const cplx zero = make_double2(0.0, 0.0);
        const int k = t + m * T;
        const cplx a = k < nkeep ? K[k] : zero;
        const cplx b = M - k < nkeep ? K[M - k] : zero;
        if (k == 0) {
            x[m] = make_double2(a.x + b.x, a.x - b.x);
        } else {
            const cplx s = make_double2(a.x + b.x, a.y - b.y);
            const cplx d = make_double2(a.x - b.x, a.y + b.y);
            constexpr double cm = B2C16<m*(16 / E)>::c, sm = B2C16<m*(16 / E)>::s;
            // exp(+2 pi i k / N) = conj(wt) * (cm + i sm)
            const cplx w = make_double2(wt.x * cm + wt.y * sm, wt.x * sm - wt.y * cm);
            const cplx e = cmul(d, w);
            x[m] = make_double2(s.x - e.y, s.y + e.x);
        }
        c2r_pre_twc<N, E, m + 1>(x, K, t, wt, nkeep);
    }
}
template <int N, int E, int m>
B2_DEVINL void r2c_post_twc(const cplx (&x)[E], cplx* __restrict__ K, const cplx* plane, int t, cplx wt,
                            double scale, bool do_store, int nkeep) {
    if constexpr (m < E) {
        constexpr int M = N / 2, T = M / E;
        const int k = t + m * T;
        if (k == 0) {
            if (do_store) {
                K[0] = make_double2((x[m].x + x[m].y) * scale, 0.0);
                if (M < nkeep) K[M] = make_double2((x[m].x - x[m].y) * scale, 0.0);
            }
        } else {
            const cplx zc = cconj(plane[b2_pad<1>(M - k)]);
            const cplx s = cadd(x[m], zc);
            const cplx d = csub(x[m], zc);
            constexpr double cm = B2C16<m*(16 / E)>::c, sm = B2C16<m*(16 / E)>::s;
            // exp(-2 pi i k / N) = wt * (cm - i sm)
            const cplx w = make_double2(wt.x * cm + wt.y * sm, wt.y * cm - wt.x * sm);
            const cplx e = cmul(d, w);
            const double hs = 0.5 * scale;
            if (do_store && k < nkeep) K[k] = make_double2((s.x + e.y) * hs, (s.y - e.x) * hs);
        }
        r2c_post_twc<N, E, m + 1>(x, K, plane, t, wt, scale, do_store, nkeep);
    }
}

// c2r along a contiguous line of N reals (M = N/2 complex FFT).  On exit x[m] = (u[2n], u[2n+1]),
// n = t + m*T.  Unnormalised (FFTW c2r convention).  Imaginary parts of k=0 and k=N/2 are ignored.
template <int N, int E, bool TWC = false, class Sync>
B2_DEVINL void c2r_line(cplx (&x)[E], const cplx* __restrict__ K, cplx* plane, int t,
                        const cplx* __restrict__ twN, Sync sync, int nkeep) {
    constexpr int M = N / 2, T = M / E;
    if constexpr (TWC) {
        c2r_pre_twc<N, E, 0>(x, K, t, __ldg(twN + t), nkeep);
        fft_line<M, E, +1, 1, 2>(x, plane, t, 0, twN, sync);
        return;
    }
    const cplx zero = make_double2(0.0, 0.0);
#pragma unroll
    for (int m = 0; m < E; ++m) {
        const int k = t + m * T;
        // modes k >= nkeep are dealiased zeros (never stored): not loaded
        const cplx a = k < nkeep ? K[k] : zero;
        const cplx b = M - k < nkeep ? K[M - k] : zero;
        if (k == 0) {
            x[m] = make_double2(a.x + b.x, a.x - b.x);
        } else {
            const cplx s = make_double2(a.x + b.x, a.y - b.y);  // a + conj(b)
            const cplx d = make_double2(a.x - b.x, a.y + b.y);  // a - conj(b)
            cplx w = __ldg(twN + k);
            w.y = -w.y;  // exp(+2 pi i k / N)
            const cplx e = cmul(d, w);
            x[m] = make_double2(s.x - e.y, s.y + e.x);  // s + i e
        }
    }
    fft_line<M, E, +1, 1, 2>(x, plane, t, 0, twN, sync);
}

// r2c: x[m] = (u[2n], u[2n+1]) on entry; writes K[0..M] scaled by `scale`.
template <int N, int E, bool TWC = false, class Sync>
B2_DEVINL void r2c_line(cplx (&x)[E], cplx* __restrict__ K, cplx* plane, int t,
                        const cplx* __restrict__ twN, Sync sync, double scale, bool do_store, int nkeep) {
    constexpr int M = N / 2, T = M / E;
    fft_line<M, E, -1, 1, 2>(x, plane, t, 0, twN, sync);
    sync();
#pragma unroll
    for (int m = 0; m < E; ++m) plane[b2_pad<1>(t + m * T)] = x[m];
    sync();
    if constexpr (TWC) {
        r2c_post_twc<N, E, 0>(x, K, plane, t, __ldg(twN + t), scale, do_store, nkeep);
        return;
    }
    const double hs = 0.5 * scale;
#pragma unroll
    for (int m = 0; m < E; ++m) {
        const int k = t + m * T;
        if (k == 0) {
            if (do_store) {
                K[0] = make_double2((x[m].x + x[m].y) * scale, 0.0);
                if (M < nkeep) K[M] = make_double2((x[m].x - x[m].y) * scale, 0.0);
            }
        } else {
            const cplx zc = cconj(plane[b2_pad<1>(M - k)]);  // conj(Z[M-k])
            const cplx s = cadd(x[m], zc);
            const cplx d = csub(x[m], zc);
            const cplx w = __ldg(twN + k);
            const cplx e = cmul(d, w);
            if (do_store && k < nkeep) K[k] = make_double2((s.x + e.y) * hs, (s.y - e.x) * hs);
        }
    }
}

template <int N, int E>
struct XSync {
    static constexpr int T = (N / 2) / E;
};

// plain c2r : K (nlines, M+1) -> X (nlines, N)
template <int N, int E, int LPB>
__global__ void __launch_bounds__(LPB*((N / 2) / E))
    xpass_c2r_kernel(const cplx* __restrict__ K, double* __restrict__ X, long long nlines,
                     const cplx* __restrict__ twN) {
    extern __shared__ double b2_smem[];
    constexpr int M = N / 2, T = M / E, PS = PlaneSize<M, 1>::value;
    const int ls = threadIdx.x / T, t = threadIdx.x % T;
    long long line = (long long)blockIdx.x * LPB + ls;
    const bool active = line < nlines;
    if (!active) line = nlines - 1;
    cplx* plane = reinterpret_cast<cplx*>(b2_smem) + (size_t)ls * PS;
    cplx x[E];
    c2r_line<N, E>(x, K + line * (M + 1), plane, t, twN, SyncBlock(), M + 1);
    if (active) {
        double2* out = reinterpret_cast<double2*>(X + line * N);
#pragma unroll
        for (int m = 0; m < E; ++m) out[t + m * T] = x[m];
    }
}

// plain r2c : X (nlines, N) -> K (nlines, M+1), scaled
template <int N, int E, int LPB>
__global__ void __launch_bounds__(LPB*((N / 2) / E))
    xpass_r2c_kernel(const double* __restrict__ X, cplx* __restrict__ K, long long nlines,
                     const cplx* __restrict__ twN, double scale) {
    extern __shared__ double b2_smem[];
    constexpr int M = N / 2, T = M / E, PS = PlaneSize<M, 1>::value;
    const int ls = threadIdx.x / T, t = threadIdx.x % T;
    long long line = (long long)blockIdx.x * LPB + ls;
    const bool active = line < nlines;
    if (!active) line = nlines - 1;
    cplx* plane = reinterpret_cast<cplx*>(b2_smem) + (size_t)ls * PS;
    cplx x[E];
    const double2* in = reinterpret_cast<const double2*>(X + line * N);
#pragma unroll
    for (int m = 0; m < E; ++m) x[m] = in[t + m * T];
    r2c_line<N, E>(x, K + line * (M + 1), plane, t, twN, SyncBlock(), scale, active, M + 1);
}

// side output of the fused x pass: max |u| of the leading op.nvmax physical fields (the velocity),
// accumulated with atomicMax into op.vmax[f] -- the reductions of
// _compute_time_increment_CLF_uxuyuz (/root/reference/fluidsim/base/time_stepping/base.py:320-339)
// at no extra pass over memory.  W = number of consecutive lanes holding one line (<= 32).
template <int E, int W>
B2_DEVINL void b2_line_absmax(const cplx (&x)[E], double* dst, bool wanted) {
    double mx = 0.0;
#pragma unroll
    for (int m = 0; m < E; ++m) mx = fmax(mx, fmax(fabs(x[m].x), fabs(x[m].y)));
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o, W));
    if (wanted && (threadIdx.x % W) == 0)
        atomicMax(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)__double_as_longlong(mx));
}

// fused: NI spectral lines -> c2r -> pointwise Op -> r2c -> NO spectral lines.
// Op: struct with static NI, NO; in[NI], out[NO] pointers (field bases);
//     __device__ void point(const double* u /*NI*/, double* r /*NO*/) const.
// Physical values are parked in thread-private shared-memory slots between transforms.
template <int N, int E, int LPB, bool VMAX, class Op>
__global__ void __launch_bounds__(LPB*((N / 2) / E))
    xpass_fused_kernel(Op op, long long nlines, const cplx* __restrict__ twN, double scale, int nkeep,
                       int pitch, long long line0) {
    extern __shared__ double b2_smem[];
    constexpr int M = N / 2, T = M / E, PS = PlaneSize<M, 1>::value;
    constexpr int NI = Op::NI, NO = Op::NO;
    constexpr int PER_LS = PS + NI * M;  // complex elements per line-set
    const int ls = threadIdx.x / T, t = threadIdx.x % T;
    long long line = (long long)blockIdx.x * LPB + ls;
    const bool active = line < nlines;
    if (!active) line = nlines - 1;
    cplx* plane = reinterpret_cast<cplx*>(b2_smem) + (size_t)ls * PER_LS;
    cplx* park = plane + PS;
    const long long loff = (line + line0) * pitch;
    cplx x[E];
#pragma unroll 1
    for (int f = 0; f < NI; ++f) {
        c2r_line<N, E>(x, op.in[f] + loff, plane, t, twN, SyncBlock(), nkeep);
        if constexpr (VMAX) b2_line_absmax<E, (T < 32 ? T : 32)>(x, op.vmax + f, f < op.nvmax);
#pragma unroll
        for (int m = 0; m < E; ++m) park[(f * E + m) * T + t] = x[m];
    }
#pragma unroll
    for (int m = 0; m < E; ++m) {
        double ue[NI], uo[NI], re[NO], ro[NO];
#pragma unroll
        for (int f = 0; f < NI; ++f) {
            const cplx v = park[(f * E + m) * T + t];
            ue[f] = v.x;
            uo[f] = v.y;
        }
        op.point(ue, re);
        op.point(uo, ro);
#pragma unroll
        for (int o = 0; o < NO; ++o) park[(o * E + m) * T + t] = make_double2(re[o], ro[o]);
    }
#pragma unroll 1
    for (int o = 0; o < NO; ++o) {
#pragma unroll
        for (int m = 0; m < E; ++m) x[m] = park[(o * E + m) * T + t];
        r2c_line<N, E>(x, op.out[o] + loff, plane, t, twN, SyncBlock(), scale, active, nkeep);
    }
}


// Field-parallel variant (used when T = (N/2)/E >= 32): one CTA per line, NI thread groups of T
// threads.  Group g does the c2r of input field g; after one CTA barrier groups o < NO form output
// o of the pointwise product from the parked physical lines and do its r2c.  Compared with the
// single-group kernel this multiplies the number of resident warps per line by NI, which is what
// the latency hiding of this pass needs (see profiles/).
// Op additionally provides out_of_group(g), needs(g, f) and point_g(g, u): the output formed by
// group g (chosen so that it uses the group's own register-resident field).
template <int N, int E, int MINB, bool VMAX, class Op, bool TWC = false>
__global__ void __launch_bounds__(Op::NI*((N / 2) / E), MINB)
    xpass_fused_fp_kernel(Op op, long long nlines, const cplx* __restrict__ twN, double scale, int nkeep,
                          int pitch, long long line0) {
    extern __shared__ double b2_smem[];
    constexpr int M = N / 2, T = M / E, PS = PlaneSize<M, 1>::value;
    constexpr int NI = Op::NI, NO = Op::NO;
    static_assert(T % 32 == 0 || T == 16, "field-parallel x pass: groups are whole warps or half warps");
    const int g = threadIdx.x / T, t = threadIdx.x % T;
    const long long line = (long long)blockIdx.x + line0;
    cplx* park = reinterpret_cast<cplx*>(b2_smem);    // [NI][E][T] complex
    cplx* plane = park + (size_t)NI * M + (size_t)g * PS;  // per-group exchange plane
    const long long loff = line * pitch;
    cplx x[E];
    if constexpr (T <= 32) {
        c2r_line<N, E, TWC>(x, op.in[g] + loff, plane, t, twN, SyncWarp(), nkeep);
    } else {
        c2r_line<N, E, TWC>(x, op.in[g] + loff, plane, t, twN, SyncNamed<T>{g + 1}, nkeep);
    }
    if constexpr (VMAX) b2_line_absmax<E, (T < 32 ? T : 32)>(x, op.vmax + g, g < op.nvmax);
#pragma unroll
    for (int m = 0; m < E; ++m) park[(g * E + m) * T + t] = x[m];
    __syncthreads();
    if (g >= NO) return;
    // group g still holds its own physical line (field g) in registers; the other operands of
    // output Op::out_of_group(g) come from the parked lines.
#pragma unroll
    for (int m = 0; m < E; ++m) {
        double ue[NI], uo[NI];
#pragma unroll
        for (int f = 0; f < NI; ++f) {
            if (!op.needs(g, f)) continue;
            const cplx v = (f == g) ? x[m] : park[(f * E + m) * T + t];
            ue[f] = v.x;
            uo[f] = v.y;
        }
        x[m] = make_double2(op.point_g(g, ue), op.point_g(g, uo));
    }
    cplx* const outp = op.out[op.out_of_group(g)] + loff;
    if constexpr (T <= 32) {
        r2c_line<N, E, TWC>(x, outp, plane, t, twN, SyncWarp(), scale, true, nkeep);
    } else {
        r2c_line<N, E, TWC>(x, outp, plane, t, twN, SyncNamed<T>{g + 1}, scale, true, nkeep);
    }
}

// ------------------------------------------------------------------------------- generic length
struct GenericFactors {
    int nfac;
    int fac[24];
};

// dynamic smem: 2 * N * TK complex
template <int DIR, class LoadOp, class StoreOp>
__global__ void fft_generic_kernel(int N, int TK, Geom g, LoadOp ld, StoreOp st,
                                   const cplx* __restrict__ tw, GenericFactors gf) {
    extern __shared__ double b2_smem[];
    cplx* a = reinterpret_cast<cplx*>(b2_smem);
    cplx* b = a + (size_t)N * TK;
    const int oidx = (int)blockIdx.y + g.outer0;
    const int outer = oidx < g.outer_lo ? oidx : oidx + g.outer_gap;
    const int field = blockIdx.z;
    const int col0 = blockIdx.x * TK;
    for (int idx = threadIdx.x; idx < N * TK; idx += blockDim.x) {
        const int i = idx / TK, c = idx % TK, col = col0 + c;
        cplx v = make_double2(0.0, 0.0);
        const bool inband_l = g.skip_load && i >= g.band_lo && i < g.band_hi;
        long long addr_l = (long long)outer * g.os + (long long)i * g.es + (long long)col * g.cs;
        if (g.map_load && !inband_l) addr_l = g.rows.xoff(i, outer, g.xpitch) + col;
        if (col < g.ncols && !inband_l) v = ld(field, addr_l, i, col, outer);
        a[idx] = v;
    }
    __syncthreads();
    int Ns = 1;
    for (int s = 0; s < gf.nfac; ++s) {
        const int r = gf.fac[s];
        const int nb = N / r;
        const int tws = N / (Ns * r);
        for (int idx = threadIdx.x; idx < nb * TK; idx += blockDim.x) {
            const int j = idx / TK, c = idx % TK;
            const int jm = j % Ns;
            const int base_out = (j / Ns) * Ns * r + jm;
            for (int n = 0; n < r; ++n) {
                cplx acc = make_double2(0.0, 0.0);
                for (int m = 0; m < r; ++m) {
                    const long long ph = ((long long)jm * m * tws + (long long)m * n * nb) % N;
                    cplx w = tw[ph];
                    if (DIR > 0) w.y = -w.y;
                    const cplx v = a[(size_t)(j + m * nb) * TK + c];
                    acc.x += v.x * w.x - v.y * w.y;
                    acc.y += v.x * w.y + v.y * w.x;
                }
                b[(size_t)(base_out + n * Ns) * TK + c] = acc;
            }
        }
        __syncthreads();
        cplx* tmp = a;
        a = b;
        b = tmp;
        Ns *= r;
    }
    for (int idx = threadIdx.x; idx < N * TK; idx += blockDim.x) {
        const int i = idx / TK, c = idx % TK, col = col0 + c;
        const bool inband_s = g.skip_store && i >= g.band_lo && i < g.band_hi;
        long long addr_s = (long long)outer * g.os + (long long)i * g.es + (long long)col * g.cs;
        if (g.map_store && !inband_s) addr_s = g.rows.xoff(i, outer, g.xpitch) + col;
        if (col < g.ncols && !inband_s) st(field, addr_s, i, col, outer, a[idx]);
    }
}
