// Strided-axis (y / z) FFT passes and the fused first inverse pass (curl / ns2d prologue on load).
#include "internal.h"
#include <stdlib.h>
#include <string.h>

#include "passes.cuh"
#include "strided_tma.cuh"

// per-size configuration of the strided pass: E points per thread, TK columns per CTA tile
template <int N> struct SCfg;
template <> struct SCfg<2048> { static constexpr int E = 16, TK = 4; };
template <> struct SCfg<1024> { static constexpr int E = 16, TK = 4; };
template <> struct SCfg<512>  { static constexpr int E = 16, TK = 8; };
template <> struct SCfg<256>  { static constexpr int E = 8,  TK = 8; };
template <> struct SCfg<128>  { static constexpr int E = 8,  TK = 8; };
template <> struct SCfg<64>   { static constexpr int E = 8,  TK = 16; };
template <> struct SCfg<32>   { static constexpr int E = 4,  TK = 16; };
template <> struct SCfg<16>   { static constexpr int E = 4,  TK = 16; };
template <> struct SCfg<8>    { static constexpr int E = 2,  TK = 16; };

// tuning knob (development): B2_SVAR=<variant> selects alternative (E, TK) for N = 512 / 1024
static int strided_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B2_SVAR");
        v = e ? atoi(e) : 0;
    }
    return v;
}

template <int N, int E, int TK, int DIR, bool ROWMAP, class L, class S>
static int launch_strided_cfg(Geom g, int nf, L ld, S st, const cplx* tw, cudaStream_t s) {
    constexpr size_t smem = (size_t)PlaneSize<N, TK>::value * TK * sizeof(cplx);
    auto kern = fft_strided_kernel<N, E, TK, DIR, ROWMAP, L, S>;
    static bool attr_done = false;
    if (!attr_done) {
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_done = true;
    }
    g.nf = nf;
    dim3 grid(((g.ncols + TK - 1) / TK) * nf, g.nouter, 1);
    // L2 prefetch distance (CTAs ahead); development knob B2_L2PF, default: about half a wave of
    // resident CTAs for the long lines (measured best, profiles/r2_tuning.md), off for the short ones (tiles small, many CTAs per SM)
    static const int l2pf_env = getenv("B2_L2PF") ? atoi(getenv("B2_L2PF")) : -1;
    g.l2pf = l2pf_env >= 0 ? l2pf_env : (N >= 1024 ? 150 : 0);
    kern<<<grid, TK*(N / E), smem, s>>>(g, ld, st, tw);
    B2_LAUNCH_CHECK("fft_strided_kernel");
    return 0;
}

static GenericFactors factorize(int n) {
    GenericFactors gf;
    gf.nfac = 0;
    // prefer radix 4 for powers of two, then primes
    while (n % 4 == 0) { gf.fac[gf.nfac++] = 4; n /= 4; }
    for (int pr = 2; n > 1; ++pr)
        while (n % pr == 0) { gf.fac[gf.nfac++] = pr; n /= pr; }
    return gf;
}

template <int DIR, class L, class S>
static int launch_generic(int N, Geom g, int nf, L ld, S st, const cplx* tw, cudaStream_t s) {
    int TK = 8;
    while (TK > 1 && (size_t)2 * N * TK * sizeof(cplx) > 160 * 1024) TK /= 2;
    size_t smem = (size_t)2 * N * TK * sizeof(cplx);
    if (smem > 220 * 1024) return b2i_set_error("generic FFT: line length %d too long", N);
    auto kern = fft_generic_kernel<DIR, L, S>;
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        attr_smem = 220 * 1024;
    }
    dim3 grid((g.ncols + TK - 1) / TK, g.nouter, nf);
    kern<<<grid, 256, smem, s>>>(N, TK, g, ld, st, tw, factorize(N));
    B2_LAUNCH_CHECK("fft_generic_kernel");
    return 0;
}

// adapter: InOp -> LoadOp for the small-size / generic kernels
template <class In>
struct LoadFromIn {
    In in;
    B2_DEVINL cplx operator()(int f, long long off, int i, int col, int outer) const {
        return in.xf(f, *in.ptr(f, off, i, col, outer), i, col, outer);
    }
    B2_DEVINL void prefetch(int f, long long off, int i, int col, int outer) const {
        b2_prefetch_l2(in.ptr(f, off, i, col, outer));
    }
};

// ------------------------------------------------------------------------------- TMA-pipelined path
typedef CUresult (*b2_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                       const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                       CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                       CUtensorMapFloatOOBfill);
static b2_encode_tiled_fn encode_tiled() {
    static b2_encode_tiled_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (b2_encode_tiled_fn)ptr;
        else
            cudaGetLastError();
    }
    return fn;
}
static int strided_tma_mode() {  // development knob: B2_STMA=0 disables the TMA-pipelined kernels
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B2_STMA");
        v = e ? atoi(e) : 0;
    }
    return v;
}
template <class In> struct TmaSource { static constexpr bool ok = false; };
template <> struct TmaSource<PlainIn> {
    static constexpr bool ok = true;
    static const cplx* base(const PlainIn& in, int f) { return in.in[f]; }
};

// returns 1 if the pass was launched, 0 if this path does not apply (caller falls back), < 0 on error
template <int N, int E, int TK, int DIR, bool TMA, class In, class S>
static int try_launch_strided_tma(Geom g, int nf, const In& in, S st, const cplx* tw, cudaStream_t s) {
    if (g.dim_nk == 0 || g.cs != 1 || g.map_load || g.map_store) return 0;
    b2_encode_tiled_fn enc = TMA ? encode_tiled() : nullptr;
    if (TMA && !enc) return 0;
    TmaGeom tg;
    const long long plane_stride = (long long)g.dim_n1 * g.dim_nk;
    if (g.es == g.dim_nk && g.os == plane_stride) tg.axis_mid = 1;
    else if (g.es == plane_stride && g.os == g.dim_nk) tg.axis_mid = 0;
    else return 0;
    const int lo_rows = g.skip_load ? g.band_lo : N;
    const int up_rows = g.skip_load ? N - g.band_hi : 0;
    const int maxr = lo_rows > up_rows ? lo_rows : up_rows;
    if (maxr <= 0) return 0;
    const int nb = (maxr + 255) / 256;
    tg.rb = (maxr + nb - 1) / nb;
    constexpr int RALIGN = TK * (int)sizeof(cplx) >= 128 ? 1 : 128 / (TK * (int)sizeof(cplx));
    tg.rb = (tg.rb + RALIGN - 1) / RALIGN * RALIGN;  // every box lands on a 128-byte boundary
    if (tg.rb > 256) return 0;
    if (!TMA) tg.rb = 1;  // cp.async: the buffer holds exactly the kept rows
    tg.nbl = (lo_rows + tg.rb - 1) / tg.rb;
    tg.nbu = up_rows > 0 ? (up_rows + tg.rb - 1) / tg.rb : 0;
    tg.up_row0 = g.skip_load ? g.band_hi : N;
    tg.pb_rows = (tg.nbl + tg.nbu) * tg.rb;
    tg.nct = (g.ncols + TK - 1) / TK;
    const long long ntiles = (long long)tg.nct * g.nouter * nf;
    if (ntiles <= 0 || ntiles > 0x7fffffffLL) return 0;
    tg.ntiles = (int)ntiles;
    const size_t smem = ((size_t)N * TK + (size_t)tg.pb_rows * TK) * sizeof(cplx) + 16;
    if (smem > 227 * 1024) return 0;
    TmaSet maps;
    const cuuint64_t gdim[3] = {(cuuint64_t)2 * g.dim_nk, (cuuint64_t)g.dim_n1, (cuuint64_t)g.dim_n0};
    const cuuint64_t gstr[2] = {(cuuint64_t)g.dim_nk * sizeof(cplx), (cuuint64_t)plane_stride * sizeof(cplx)};
    const cuuint32_t box_mid[3] = {2 * TK, (cuuint32_t)tg.rb, 1}, box_out[3] = {2 * TK, 1, (cuuint32_t)tg.rb};
    const cuuint32_t estr[3] = {1, 1, 1};
    static const int l2p = getenv("B2_STMA_L2") ? atoi(getenv("B2_STMA_L2")) : 1;
    PfSources srcs;
    for (int f = 0; f < B2_MAXF; ++f) srcs.src[f] = TmaSource<In>::base(in, f < nf ? f : 0);
    for (int f = 0; TMA && f < nf; ++f) {
        const cplx* base = TmaSource<In>::base(in, f);
        if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return 0;
        CUresult r = enc(&maps.m[f], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)base, gdim, gstr,
                         tg.axis_mid ? box_mid : box_out, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE,
                         l2p == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                  : (l2p == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE),
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return 0;
    }
    for (int f = nf; TMA && f < B2_MAXF; ++f) maps.m[f] = maps.m[0];
    if (!TMA) memset(&maps, 0, sizeof(maps));
    auto kern = fft_strided_tma_kernel<N, E, TK, DIR, TMA, In, S>;
    cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce != cudaSuccess) return b2i_set_error("fft_strided_tma_kernel: %s", cudaGetErrorString(ce));
    int dev = 0, nsm = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TK * (N / E), smem);
    if (occ < 1) return 0;
    long long grid = (long long)nsm * occ;
    if (grid > ntiles) grid = ntiles;
    g.nf = nf;
    kern<<<(unsigned)grid, TK*(N / E), smem, s>>>(maps, srcs, g, tg, in, st, tw);
    B2_LAUNCH_CHECK("fft_strided_tma_kernel");
    return 1;
}

template <int N, int DIR, bool ROWMAP, class L, class S>
static int launch_strided_n(Geom g, int nf, L ld, S st, const cplx* tw, cudaStream_t s) {
    if constexpr (N == 1024) {
        // measured (profiles/r1_tuning.md): the plane-strided z pass wants 128-byte row segments
        // (TK = 8, one 512-thread CTA per SM), the y pass prefers two 256-thread CTAs (TK = 4)
        const bool zlike = g.nouter == 1 || g.wide;
        switch (strided_variant()) {
            case 1: return launch_strided_cfg<N, 16, 8, DIR, ROWMAP>(g, nf, ld, st, tw, s);
            case 2: return launch_strided_cfg<N, 16, 4, DIR, ROWMAP>(g, nf, ld, st, tw, s);
        }
        if (zlike) return launch_strided_cfg<N, 16, 8, DIR, ROWMAP>(g, nf, ld, st, tw, s);
        return launch_strided_cfg<N, 16, 4, DIR, ROWMAP>(g, nf, ld, st, tw, s);
    }
    if constexpr (N == 512) {
        switch (strided_variant()) {
            case 1: return launch_strided_cfg<N, 16, 4, DIR, ROWMAP>(g, nf, ld, st, tw, s);
            case 2: return launch_strided_cfg<N, 8, 8, DIR, ROWMAP>(g, nf, ld, st, tw, s);
            case 3: return launch_strided_cfg<N, 8, 4, DIR, ROWMAP>(g, nf, ld, st, tw, s);
        }
    }
    return launch_strided_cfg<N, SCfg<N>::E, SCfg<N>::TK, DIR, ROWMAP>(g, nf, ld, st, tw, s);
}

template <int DIR, bool ROWMAP = false, class In, class S>
static int launch_strided(bool fast, int N, Geom g, int nf, In in, S st, const cplx* tw, cudaStream_t s) {
    LoadFromIn<In> ld{in};
    if constexpr (!ROWMAP && TmaSource<In>::ok) {
        if (fast && (N == 512 || N == 1024 || N == 2048)) {
            // development knob B2_STMA: 0 LDG kernels, 1 TMA tensor-map prefetch (TK 4 / B2_STMA_TK=8),
            // 2 cp.async prefetch (TK 4)
            int r = 0;
            const int mode = strided_tma_mode();
            static const int tk8 = getenv("B2_STMA_TK") ? atoi(getenv("B2_STMA_TK")) == 8 : 0;
            if (mode == 1) {
                if (N == 512) r = tk8 ? try_launch_strided_tma<512, 16, 8, DIR, true>(g, nf, in, st, tw, s)
                                      : try_launch_strided_tma<512, 16, 4, DIR, true>(g, nf, in, st, tw, s);
                else if (N == 1024) r = tk8 ? try_launch_strided_tma<1024, 16, 8, DIR, true>(g, nf, in, st, tw, s)
                                            : try_launch_strided_tma<1024, 16, 4, DIR, true>(g, nf, in, st, tw, s);
                else r = try_launch_strided_tma<2048, 16, 4, DIR, true>(g, nf, in, st, tw, s);
            } else if (mode == 2) {
                if (N == 512) r = try_launch_strided_tma<512, 16, 4, DIR, false>(g, nf, in, st, tw, s);
                else if (N == 1024) r = try_launch_strided_tma<1024, 16, 4, DIR, false>(g, nf, in, st, tw, s);
                else r = try_launch_strided_tma<2048, 16, 4, DIR, false>(g, nf, in, st, tw, s);
            }
            if (r != 0) return r < 0 ? r : 0;
        }
    }
    if (fast) {
        switch (N) {
#define B2_CASE(n) case n: return launch_strided_n<n, DIR, ROWMAP>(g, nf, ld, st, tw, s);
            B2_CASE(8) B2_CASE(16) B2_CASE(32) B2_CASE(64) B2_CASE(128) B2_CASE(256) B2_CASE(512)
            B2_CASE(1024) B2_CASE(2048)
#undef B2_CASE
        }
    }
    return launch_generic<DIR>(N, g, nf, ld, st, tw, s);
}

// pruned geometries: only kept kx columns / kept outer rows are visited, the dealiased band of the
// transformed axis is skipped on load (inverse) or on store (forward)
static Geom geom_pruned(const b2_plan* p, int axis, int dir) {
    Geom g = geom_init();
    g.ncols = p->keepx;
    if (axis == 0) {  // z pass over kept (ky, kx) columns
        g.nouter = p->keep1_lo + (p->n1 - p->keep1_hi);
        g.outer_lo = p->keep1_lo;
        g.outer_gap = p->keep1_hi - p->keep1_lo;
        g.es = (long long)p->n1 * p->nk;
        g.os = p->nk;
        g.band_lo = p->keep0_lo;
        g.band_hi = p->keep0_hi;
        g.wide = 1;
    } else {  // y pass over kept kx columns, all z
        g.nouter = p->n0;
        g.es = p->nk;
        g.os = (long long)p->n1 * p->nk;
        g.band_lo = p->keep1_lo;
        g.band_hi = p->keep1_hi;
    }
    g.skip_load = dir > 0;
    g.skip_store = dir < 0;
    g.dim_nk = p->nk; g.dim_n1 = p->n1; g.dim_n0 = p->n0;
    return g;
}

static Geom geom_for_axis(const b2_plan* p, int axis) {
    Geom g = geom_init();
    if (axis == 0) {  // z pass: columns = flattened (i1, kx)
        g.ncols = p->n1 * p->nk;  // < 2^31 for all supported sizes
        g.nouter = 1;
        g.es = (long long)p->n1 * p->nk;
        g.os = 0;
    } else {  // y pass
        g.ncols = p->nk;
        g.nouter = p->n0;
        g.es = p->nk;
        g.os = (long long)p->n1 * p->nk;
    }
    g.cs = 1;
    g.nf = 1;
    g.dim_nk = p->nk; g.dim_n1 = p->n1; g.dim_n0 = p->n0;
    return g;
}

int b2i_strided_plain(b2_plan* p, int axis, int dir, const cplx* const* in, cplx* const* out, int nf,
                      double scale, cudaStream_t s, bool pruned, int outer0, int nouter) {
    if (axis == 0 && p->n0 == 1) return 0;
    if (nf > B2_MAXF) return b2i_set_error("too many fields");
    Geom g = pruned ? geom_pruned(p, axis, dir) : geom_for_axis(p, axis);
    if (nouter >= 0) {  // chunked launch over a range of outer indices (y pass: z planes)
        g.outer0 = outer0;
        g.nouter = nouter;
    }
    const int N = axis == 0 ? p->n0 : p->n1;
    const bool fast = axis == 0 ? p->fast0 : p->fast1;
    const cplx* tw = axis == 0 ? p->tw0 : p->tw1;
    PlainIn ld;
    for (int f = 0; f < nf; ++f) ld.in[f] = in[f];
    if (scale == 1.0) {
        PlainStore st;
        for (int f = 0; f < nf; ++f) st.out[f] = out[f];
        return dir < 0 ? launch_strided<-1>(fast, N, g, nf, ld, st, tw, s)
                       : launch_strided<+1>(fast, N, g, nf, ld, st, tw, s);
    }
    ScaleStore st;
    st.scale = scale;
    for (int f = 0; f < nf; ++f) st.out[f] = out[f];
    return dir < 0 ? launch_strided<-1>(fast, N, g, nf, ld, st, tw, s)
                   : launch_strided<+1>(fast, N, g, nf, ld, st, tw, s);
}

// ------------------------------------------------------------------------------- first inverse pass
// ns2d: vecfft_from_rotfft + gradfft_from_fft applied to the prefetched value
// (/root/reference/fluidsim/solvers/ns2d/solver.py:158,165): four outputs from one input field.
struct Ns2dIn {
    const cplx* rot;
    const double *k1, *kx;
    B2_DEVINL const cplx* ptr(int f, long long off, int i, int col, int outer) const { return rot + off; }
    B2_DEVINL cplx xf(int fld, cplx r, int i, int col, int outer) const {
        const double Ky = __ldg(k1 + i), Kx = __ldg(kx + col);
        double K2 = Kx * Kx + Ky * Ky;
        if (i == 0 && col == 0) K2 = 1e-14;
        const double inv = 1.0 / K2;
        double cf;
        switch (fld) {
            case 0: cf = Ky * inv; break;     // ux =  i KY / K2 rot
            case 1: cf = -(Kx * inv); break;  // uy = -i KX / K2 rot
            case 2: cf = Kx; break;           // d_x rot = i KX rot
            default: cf = Ky; break;          // d_y rot = i KY rot
        }
        return make_double2(-cf * r.y, cf * r.x);
    }
};

template <> struct TmaSource<Ns2dIn> {
    static constexpr bool ok = true;
    static const cplx* base(const Ns2dIn& in, int) { return in.rot; }
};

// ns3d / strat: in[] = nvar stage-input fields (v, [b]); the vorticity has already been written to
// out[3..5] by the RK epilogue (or by b2i_curl for the first stage).  One plain launch transforms
// v: in -> out[0..2], omega: out[3..5] in place, b: in[3] -> out[6].
int b2i_first_inverse_pass(b2_plan* p, const cplx* const* in, cplx* const* out, cudaStream_t s) {
    if (p->solver == B2_SOLVER_NS2D) {
        Ns2dIn ld;
        ld.rot = in[0];
        ld.k1 = p->k1;
        ld.kx = p->kx;
        PlainStore st;
        for (int f = 0; f < 4; ++f) st.out[f] = out[f];
        return launch_strided<+1>(p->fast1, p->n1, p->prune ? geom_pruned(p, 1, +1) : geom_for_axis(p, 1), 4, ld,
                                  st, p->tw1, s);
    }
    const int nv = p->solver == B2_SOLVER_NS3D_STRAT ? 4 : 3;
    const int nout = nv + 3;
    PlainIn ld;
    PlainStore st;
    for (int f = 0; f < nout; ++f) st.out[f] = out[f];
    for (int f = 0; f < 3; ++f) ld.in[f] = in[f];
    for (int f = 3; f < 6; ++f) ld.in[f] = out[f];
    if (nv == 4) ld.in[6] = in[3];
    const int axis = p->n0 > 1 ? 0 : 1;
    return launch_strided<+1>(axis == 0 ? p->fast0 : p->fast1, axis == 0 ? p->n0 : p->n1,
                              p->prune ? geom_pruned(p, axis, +1) : geom_for_axis(p, axis), nout, ld, st,
                              axis == 0 ? p->tw0 : p->tw1, s);
}

// ------------------------------------------------------------------------------- slab passes
// Index map between K-layout coordinates (yl = outer, z = i, kx = col) and the exchange layout of one
// field: [z chunk c][peer q = z / nzl][z within chunk][kept local ky row][kx < pitch].  The z range of
// every peer is cut in `nc` chunks so that the all-to-all of chunk c+1 overlaps the y / x passes of
// chunk c; chunk regions are `cstride` elements apart.  Inside a peer block z is OUTSIDE ky: on the
// receiving side the y passes then walk rows that are `pitch` elements apart (the access pattern of the
// single-GPU y pass) instead of whole planes apart.  With pruning only the kept local ky rows (compact
// index) and the kept kx columns are exchanged.
struct SlabMapper {
    int nzl, nkl, pitch, y_lo, y_gap, zc;
    int sh_nzl, sh_zc;  // log2 of nzl / zc when they are powers of two (-1 otherwise): no integer
                        // divisions in front of the loads of the z-forward pass (measured: the
                        // passes that LOAD through a mapper ran 1.45x slower per byte than plain ones)
    long long cstride;
    B2_DEVINL long long operator()(int z, int kx, int yl) const {
        const int q = sh_nzl >= 0 ? (z >> sh_nzl) : z / nzl;
        const int zl = z - q * nzl;
        const int c = sh_zc >= 0 ? (zl >> sh_zc) : zl / zc;
        const int zlc = zl - c * zc;
        const int ylc = yl < y_lo ? yl : yl - y_gap;
        return c * cstride + (((long long)q * zc + zlc) * nkl + ylc) * pitch + kx;
    }
};
struct SlabStore {
    cplx* out[B2_MAXF];
    SlabMapper map;
    B2_DEVINL void operator()(int f, long long off, int i, int col, int outer, cplx v) const {
        out[f][map(i, col, outer)] = v;
    }
};
struct SlabIn {
    const cplx* in[B2_MAXF];
    SlabMapper map;
    B2_DEVINL const cplx* ptr(int f, long long off, int i, int col, int outer) const {
        return in[f] + map(i, col, outer);
    }
    B2_DEVINL cplx xf(int f, cplx v, int i, int col, int outer) const { return v; }
};

int b2i_slab_zpass(b2_plan* p, int dir, const cplx* const* in, cplx* const* out, int nf, cudaStream_t s) {
    if (nf > B2_MAXF) return b2i_set_error("too many fields");
    Geom g = geom_init();
    g.es = p->nk;
    g.os = (long long)p->n1 * p->nk;
    g.nf = nf;
    g.dim_nk = p->nk; g.dim_n1 = p->n1; g.dim_n0 = p->n0;
    SlabMapper map;
    map.nzl = p->nzl;
    map.zc = p->nzl / p->slab_nc;
    if (p->prune) {
        g.ncols = p->keepx;
        g.nouter = p->keep0_lo + (p->n0 - p->keep0_hi);  // kept local ky rows
        g.outer_lo = p->keep0_lo;
        g.outer_gap = p->keep0_hi - p->keep0_lo;
        g.band_lo = p->keep1_lo;  // kz band
        g.band_hi = p->keep1_hi;
        g.skip_load = dir > 0;
        g.skip_store = dir < 0;
        map.nkl = g.nouter; map.pitch = p->keepx; map.y_lo = p->keep0_lo; map.y_gap = g.outer_gap;
    } else {
        g.ncols = p->nk;
        g.nouter = p->n0;
        map.nkl = p->n0; map.pitch = p->nk; map.y_lo = 1 << 30; map.y_gap = 0;
    }
    map.cstride = (long long)p->gy * map.zc * map.pitch;
    auto log2_or_m1 = [](int v) {
        if (v <= 0 || (v & (v - 1))) return -1;
        int sh = 0;
        while ((1 << sh) < v) ++sh;
        return sh;
    };
    map.sh_nzl = log2_or_m1(map.nzl);
    map.sh_zc = log2_or_m1(map.zc);
    if (g.nouter == 0) return 0;  // this rank owns only dealiased ky rows
    if (dir > 0) {
        PlainIn ld;
        SlabStore st;
        st.map = map;
        for (int f = 0; f < nf; ++f) { ld.in[f] = in[f]; st.out[f] = out[f]; }
        return launch_strided<+1>(p->fast1, p->n1, g, nf, ld, st, p->tw1, s);
    }
    SlabIn ld;
    ld.map = map;
    PlainStore st;
    for (int f = 0; f < nf; ++f) { ld.in[f] = in[f]; st.out[f] = out[f]; }
    return launch_strided<-1>(p->fast1, p->n1, g, nf, ld, st, p->tw1, s);
}

// y pass on the z-slab side, z chunk `chunk` (sub-array of zc = nz_loc / nc planes).  One side of the
// pass is the exchanged array (rows grouped by owning rank, only kept ky rows and kx < pitch columns:
// RowMap), the other the natural array (zc, ny, pitch) the fused x pass works on: the inverse pass
// expands exchanged -> natural (in -> out), the forward pass stores the kept rows back.  `in` / `out`
// are the field bases; the chunk offsets (exchanged: kept rows, natural: ny rows) are applied here.
int b2i_slab_ypass(b2_plan* p, int dir, const cplx* const* in, cplx* const* out, int nf, int chunk,
                   cudaStream_t s) {
    if (nf > B2_MAXF) return b2i_set_error("too many fields");
    Geom g = geom_init();
    const int pitch = p->prune ? p->keepx : p->nk;
    const int zc = p->nzl / p->slab_nc;
    g.ncols = pitch;
    g.nouter = zc;
    g.es = pitch;
    g.os = (long long)p->gy * pitch;
    g.nf = nf;
    g.xpitch = pitch;
    g.rows.P = p->nranks;
    g.rows.nyl = p->nyl;
    g.rows.cyclic = p->ky_cyclic ? 1 : 0;
    g.rows.shift = -1;
    const int divisor = p->ky_cyclic ? p->nranks : p->nyl;
    if ((divisor & (divisor - 1)) == 0) {
        int sh = 0;
        while ((1 << sh) < divisor) ++sh;
        g.rows.shift = sh;
    }
    long long start = 0;
    for (int r = 0; r < p->nranks; ++r) {
        int lo, hi;
        b2i_slab_local_band(p, r, &lo, &hi);
        g.rows.lo[r] = lo;
        g.rows.gap[r] = hi - lo;
        g.rows.nkr[r] = p->nyl - (hi - lo);
        g.rows.blk[r] = start;
        start += (long long)g.rows.nkr[r] * zc * pitch;
    }
    const long long cs_x = start;                       // exchanged chunk
    const long long cs_n = (long long)p->gy * zc * pitch;  // natural chunk
    long long in_off, out_off;
    if (dir > 0) { g.map_load = 1; in_off = chunk * cs_x; out_off = chunk * cs_n; }
    else { g.map_store = 1; in_off = chunk * cs_n; out_off = chunk * cs_x; }
    if (p->prune) {
        g.band_lo = p->gyk_lo;
        g.band_hi = p->gyk_hi;
        if (dir > 0) g.skip_load = 1; else g.skip_store = 1;
    }
    PlainIn ld;
    PlainStore st;
    for (int f = 0; f < nf; ++f) { ld.in[f] = in[f] + in_off; st.out[f] = out[f] + out_off; }
    return dir < 0 ? launch_strided<-1, true>(p->fasty, p->gy, g, nf, ld, st, p->twy, s)
                   : launch_strided<+1, true>(p->fasty, p->gy, g, nf, ld, st, p->twy, s);
}
