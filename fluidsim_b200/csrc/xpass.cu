// Contiguous-axis (x) passes: c2r, r2c and the fused c2r -> physical-space product -> r2c kernel.
#include <stdlib.h>

#include "internal.h"
#include "passes.cuh"
#include "xpair_op.h"

static bool g_xpass_share_sm = false;
void b2i_xpass_share_sm(bool on) { g_xpass_share_sm = on; }

// per-size configuration: E points (of the N/2 complex FFT) per thread
template <int N> struct XCfg;
template <> struct XCfg<2048> { static constexpr int E = 16; };  // M=1024, T=64
template <> struct XCfg<1024> { static constexpr int E = 8; };   // M=512,  T=64
template <> struct XCfg<512>  { static constexpr int E = 8; };   // M=256,  T=32
template <> struct XCfg<256>  { static constexpr int E = 8; };   // M=128,  T=16
template <> struct XCfg<128>  { static constexpr int E = 8; };   // M=64,   T=8
template <> struct XCfg<64>   { static constexpr int E = 4; };   // M=32,   T=8
template <> struct XCfg<32>   { static constexpr int E = 4; };   // M=16,   T=4
template <> struct XCfg<16>   { static constexpr int E = 2; };   // M=8,    T=4
template <> struct XCfg<8>    { static constexpr int E = 2; };   // M=4,    T=2

// lines per block: as many as fit `budget` bytes of smem, at most `maxthr` threads, block size a
// multiple of 32
constexpr int lpb_for(int T, size_t per_ls_bytes, size_t budget, int maxthr) {
    int l = (int)(budget / per_ls_bytes);
    if (l < 1) l = 1;
    if (l * T > maxthr) l = maxthr / T;
    const int q = T >= 32 ? 1 : 32 / T;
    l = (l / q) * q;
    if (l < q) l = q;
    return l;
}

// ------------------------------------------------------------------------------- product ops
// ns3d: f = v x omega  (vector_product, /root/reference/fluidsim/solvers/ns3d/solver.py:226)
struct OpNS3D {
    static constexpr int NI = 6, NO = 3;
    const cplx* in[NI];
    cplx* out[NO];
    double* vmax;  // optional max|u| side output (CFL), nvmax leading fields
    int nvmax;
    B2_DEVINL void point(const double* u, double* r) const {
        r[0] = u[1] * u[5] - u[2] * u[4];
        r[1] = u[2] * u[3] - u[0] * u[5];
        r[2] = u[0] * u[4] - u[1] * u[3];
    }
    // group 0 (holds vx) forms fy, group 1 (vy) forms fz, group 2 (vz) forms fx
    B2_DEVINL int out_of_group(int g) const { return g == 0 ? 1 : (g == 1 ? 2 : 0); }
    B2_DEVINL bool needs(int g, int f) const {
        if (g == 0) return f == 0 || f == 2 || f == 3 || f == 5;
        if (g == 1) return f == 0 || f == 1 || f == 3 || f == 4;
        return f == 1 || f == 2 || f == 4 || f == 5;
    }
    B2_DEVINL double point_g(int g, const double* u) const {
        if (g == 0) return u[2] * u[3] - u[0] * u[5];
        if (g == 1) return u[0] * u[4] - u[1] * u[3];
        return u[1] * u[5] - u[2] * u[4];
    }
};
// ns3d.strat: f = v x omega and v*b (div_vb_fft_from_vb, strat/solver.py:206)
struct OpStrat {
    static constexpr int NI = 7, NO = 6;
    const cplx* in[NI];
    cplx* out[NO];
    double* vmax;  // optional max|u| side output (CFL), nvmax leading fields
    int nvmax;
    B2_DEVINL void point(const double* u, double* r) const {
        r[0] = u[1] * u[5] - u[2] * u[4];
        r[1] = u[2] * u[3] - u[0] * u[5];
        r[2] = u[0] * u[4] - u[1] * u[3];
        r[3] = u[0] * u[6];
        r[4] = u[1] * u[6];
        r[5] = u[2] * u[6];
    }
    // groups 0..2 as ns3d; groups 3..5 (holding omega) form vx b, vy b, vz b
    B2_DEVINL int out_of_group(int g) const { return g == 0 ? 1 : (g == 1 ? 2 : (g == 2 ? 0 : g)); }
    B2_DEVINL bool needs(int g, int f) const {
        if (g == 0) return f == 0 || f == 2 || f == 3 || f == 5;
        if (g == 1) return f == 0 || f == 1 || f == 3 || f == 4;
        if (g == 2) return f == 1 || f == 2 || f == 4 || f == 5;
        return f == g - 3 || f == 6;
    }
    B2_DEVINL double point_g(int g, const double* u) const {
        if (g == 0) return u[2] * u[3] - u[0] * u[5];
        if (g == 1) return u[0] * u[4] - u[1] * u[3];
        if (g == 2) return u[1] * u[5] - u[2] * u[4];
        return u[g - 3] * u[6];
    }
};
// ns2d: Frot = -ux d_x rot - uy (d_y rot + beta)  (compute_Frot, solvers/ns2d/solver.py:34-38)
struct OpNS2D {
    static constexpr int NI = 4, NO = 1;
    const cplx* in[NI];
    cplx* out[NO];
    double* vmax;  // optional max|u| side output (CFL), nvmax leading fields
    int nvmax;
    double beta;
    B2_DEVINL void point(const double* u, double* r) const {
        r[0] = beta == 0.0 ? -u[0] * u[2] - u[1] * u[3] : -u[0] * u[2] - u[1] * (u[3] + beta);
    }
    B2_DEVINL int out_of_group(int g) const { return 0; }
    B2_DEVINL bool needs(int g, int f) const { return true; }
    B2_DEVINL double point_g(int g, const double* u) const {
        return beta == 0.0 ? -u[0] * u[2] - u[1] * u[3] : -u[0] * u[2] - u[1] * (u[3] + beta);
    }
};

template <int N>
static int launch_c2r_n(const cplx* K, double* X, long long nlines, const cplx* tw, cudaStream_t s) {
    constexpr int E = XCfg<N>::E, M = N / 2, T = M / E;
    constexpr size_t per_ls = (size_t)PlaneSize<M, 1>::value * sizeof(cplx);
    constexpr int LPB = lpb_for(T, per_ls, 48 * 1024, 256);
    auto kern = xpass_c2r_kernel<N, E, LPB>;
    const unsigned grid = (unsigned)((nlines + LPB - 1) / LPB);
    kern<<<grid, LPB * T, LPB * per_ls, s>>>(K, X, nlines, tw);
    B2_LAUNCH_CHECK("xpass_c2r_kernel");
    return 0;
}

template <int N>
static int launch_r2c_n(const double* X, cplx* K, long long nlines, const cplx* tw, double scale,
                        cudaStream_t s) {
    constexpr int E = XCfg<N>::E, M = N / 2, T = M / E;
    constexpr size_t per_ls = (size_t)PlaneSize<M, 1>::value * sizeof(cplx);
    constexpr int LPB = lpb_for(T, per_ls, 48 * 1024, 256);
    auto kern = xpass_r2c_kernel<N, E, LPB>;
    const unsigned grid = (unsigned)((nlines + LPB - 1) / LPB);
    kern<<<grid, LPB * T, LPB * per_ls, s>>>(X, K, nlines, tw, scale);
    B2_LAUNCH_CHECK("xpass_r2c_kernel");
    return 0;
}

template <int N, class Op>
static int launch_fused_fp_n(Op op, long long nlines, const cplx* tw, double scale, int nkeep, int pitch, long long line0,
                  cudaStream_t s) {
    constexpr int E = XCfg<N>::E, M = N / 2, T = M / E;
    constexpr size_t smem = ((size_t)Op::NI * M + (size_t)Op::NI * PlaneSize<M, 1>::value) * sizeof(cplx);
    // b2i_xpass_share_sm(true): run with <= 80 registers and a shared-memory request padded above
    // half an SM, so that exactly one x-pass CTA plus one strided-pass CTA are co-resident per SM
    // (the x pass is L1/FP64 bound, the y passes are HBM bound: they overlap, see api.cu)
    const bool share = g_xpass_share_sm;
    const bool vm = op.vmax != nullptr;
    // experimental, off by default: B2_XTWC=1 selects the computed pre/post twiddle variant
    static const bool twc = getenv("B2_XTWC") != nullptr;
    auto kern = share ? xpass_fused_fp_kernel<N, E, 2, false, Op>
                      : (vm ? xpass_fused_fp_kernel<N, E, 1, true, Op>
                            : (twc ? xpass_fused_fp_kernel<N, E, 1, false, Op, true>
                                   : xpass_fused_fp_kernel<N, E, 1, false, Op>));
    if (share && vm) return b2i_set_error("x pass: CFL side output is not available in SM-sharing mode");
    size_t smem_req = smem;
    if (share && smem_req < 118 * 1024) smem_req = 118 * 1024;
    static bool attr_done[4] = {false, false, false, false};
    const int slot = share ? 1 : (vm ? 2 : (twc ? 3 : 0));
    if (!attr_done[slot]) {
        if (smem_req > 48 * 1024)
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_req);
        attr_done[slot] = true;
    }
    kern<<<(unsigned)nlines, Op::NI * T, smem_req, s>>>(op, nlines, tw, scale, nkeep, pitch, line0);
    B2_LAUNCH_CHECK("xpass_fused_fp_kernel");
    return 0;
}

template <int N, class Op>
static int launch_fused_n(Op op, long long nlines, const cplx* tw, double scale, int nkeep, int pitch, long long line0,
                  cudaStream_t s) {
    constexpr int E = XCfg<N>::E, M = N / 2, T = M / E;
    constexpr size_t smem_fp = ((size_t)Op::NI * M + (size_t)Op::NI * PlaneSize<M, 1>::value) * sizeof(cplx);
    if constexpr ((T % 32 == 0 || (T == 16 && (Op::NI * T) % 32 == 0)) && smem_fp <= 227 * 1024)
        return launch_fused_fp_n<N>(op, nlines, tw, scale, nkeep, pitch, line0, s);
    constexpr size_t per_ls = ((size_t)PlaneSize<M, 1>::value + (size_t)Op::NI * M) * sizeof(cplx);
    constexpr int LPB = lpb_for(T, per_ls, 100 * 1024, 256);
    constexpr size_t smem = LPB * per_ls;
    static_assert(smem <= 227 * 1024, "x-pass tile does not fit shared memory");
    const bool vm = op.vmax != nullptr;
    auto kern = vm ? xpass_fused_kernel<N, E, LPB, true, Op> : xpass_fused_kernel<N, E, LPB, false, Op>;
    static bool attr_done[2] = {false, false};
    if (!attr_done[vm]) {
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_done[vm] = true;
    }
    const unsigned grid = (unsigned)((nlines + LPB - 1) / LPB);
    kern<<<grid, LPB * T, smem, s>>>(op, nlines, tw, scale, nkeep, pitch, line0);
    B2_LAUNCH_CHECK("xpass_fused_kernel");
    return 0;
}

#define B2_XSIZES(X) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048)

// ------------------------------------------------------------------------------- generic length
// any nx (odd included): full complex FFT of the Hermitian-extended line.
struct GenC2RLoad {
    const cplx* K;
    int nx, nk;
    B2_DEVINL cplx operator()(int f, long long off, int i, int col, int outer) const {
        const long long base = (long long)col * nk;
        if (i < nk) return K[base + i];
        return cconj(K[base + nx - i]);
    }
};
struct GenC2RStore {
    double* X;
    int nx;
    B2_DEVINL void operator()(int f, long long off, int i, int col, int outer, cplx v) const {
        X[(long long)col * nx + i] = v.x;
    }
};
struct GenR2CLoad {
    const double* X;
    int nx;
    B2_DEVINL cplx operator()(int f, long long off, int i, int col, int outer) const {
        return make_double2(X[(long long)col * nx + i], 0.0);
    }
};
struct GenR2CStore {
    cplx* K;
    int nk;
    double scale;
    B2_DEVINL void operator()(int f, long long off, int i, int col, int outer, cplx v) const {
        if (i < nk) K[(long long)col * nk + i] = cscale(v, scale);
    }
};

static GenericFactors factorize(int n) {
    GenericFactors gf;
    gf.nfac = 0;
    while (n % 4 == 0) { gf.fac[gf.nfac++] = 4; n /= 4; }
    for (int pr = 2; n > 1; ++pr)
        while (n % pr == 0) { gf.fac[gf.nfac++] = pr; n /= pr; }
    return gf;
}

template <int DIR, class L, class S>
static int launch_generic_x(int N, long long nlines, L ld, S st, const cplx* tw, cudaStream_t s) {
    int TK = 8;
    while (TK > 1 && (size_t)2 * N * TK * sizeof(cplx) > 160 * 1024) TK /= 2;
    size_t smem = (size_t)2 * N * TK * sizeof(cplx);
    if (smem > 220 * 1024) return b2i_set_error("generic FFT: line length %d too long", N);
    auto kern = fft_generic_kernel<DIR, L, S>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        attr_done = true;
    }
    Geom g = geom_init();
    g.ncols = (int)nlines;
    g.cs = 0;
    dim3 grid((unsigned)((nlines + TK - 1) / TK), 1, 1);
    kern<<<grid, 256, smem, s>>>(N, TK, g, ld, st, tw, factorize(N));
    B2_LAUNCH_CHECK("fft_generic_kernel(x)");
    return 0;
}

int b2i_xpass_c2r(b2_plan* p, const cplx* K, double* X, cudaStream_t s) {
    const long long nlines = (long long)p->n0 * p->n1;
    if (p->fast2) {
        switch (p->n2) {
#define B2_CASE(n) case n: return launch_c2r_n<n>(K, X, nlines, p->tw2, s);
            B2_XSIZES(B2_CASE)
#undef B2_CASE
        }
    }
    GenC2RLoad ld{K, p->n2, p->nk};
    GenC2RStore st{X, p->n2};
    return launch_generic_x<+1>(p->n2, nlines, ld, st, p->tw2, s);
}

int b2i_xpass_r2c(b2_plan* p, const double* X, cplx* K, double scale, cudaStream_t s) {
    const long long nlines = (long long)p->n0 * p->n1;
    if (p->fast2) {
        switch (p->n2) {
#define B2_CASE(n) case n: return launch_r2c_n<n>(X, K, nlines, p->tw2, scale, s);
            B2_XSIZES(B2_CASE)
#undef B2_CASE
        }
    }
    GenR2CLoad ld{X, p->n2};
    GenR2CStore st{K, p->nk, scale};
    return launch_generic_x<-1>(p->n2, nlines, ld, st, p->tw2, s);
}

template <class Op>
static int launch_fused(b2_plan* p, Op op, long long nlines, double scale, int nkeep, int pitch, long long line0,
                  cudaStream_t s) {
    switch (p->n2) {
#define B2_CASE(n) case n: return launch_fused_n<n>(op, nlines, p->tw2, scale, nkeep, pitch, line0, s);
        B2_XSIZES(B2_CASE)
#undef B2_CASE
    }
    return b2i_set_error("fused x pass: nx=%d not supported (power of two in [8, 2048])", p->n2);
}

int b2i_xpass_fused(b2_plan* p, cplx* const* W, long long nlines, double scale, int nkeep, int pitch, long long line0,
                    cudaStream_t s, double* vmax) {
    if (!p->fast2) return b2i_set_error("fused x pass needs a power-of-two nx");
    static const bool use_old = getenv("B2_XPASS_OLD") != nullptr;  // development: round-1 kernels for A/B runs
    if (!use_old) {
        PairOp op;
        op.vmax = vmax;
        op.beta = p->beta;
        const int nin = p->solver == B2_SOLVER_NS3D ? 6 : (p->solver == B2_SOLVER_NS3D_STRAT ? 7 : 4);
        const int nout = p->solver == B2_SOLVER_NS3D ? 3 : (p->solver == B2_SOLVER_NS3D_STRAT ? 6 : 1);
        for (int f = 0; f < 7; ++f) op.in[f] = W[f < nin ? f : 0];
        for (int f = 0; f < 6; ++f) op.out[f] = W[f < nout ? f : 0];
        if (p->solver == B2_SOLVER_NS3D) return b2i_xpair_ns3d(p, op, nlines, scale, nkeep, pitch, line0, s);
        if (p->solver == B2_SOLVER_NS3D_STRAT) return b2i_xpair_strat(p, op, nlines, scale, nkeep, pitch, line0, s);
        return b2i_xpair_ns2d(p, op, nlines, scale, nkeep, pitch, line0, s);
    }
    if (p->solver == B2_SOLVER_NS3D) {
        OpNS3D op;
        op.vmax = vmax; op.nvmax = 3;
        for (int f = 0; f < 6; ++f) op.in[f] = W[f];
        for (int f = 0; f < 3; ++f) op.out[f] = W[f];
        return launch_fused(p, op, nlines, scale, nkeep, pitch, line0, s);
    }
    if (p->solver == B2_SOLVER_NS3D_STRAT) {
        OpStrat op;
        op.vmax = vmax; op.nvmax = 3;
        for (int f = 0; f < 7; ++f) op.in[f] = W[f];
        for (int f = 0; f < 6; ++f) op.out[f] = W[f];
        return launch_fused(p, op, nlines, scale, nkeep, pitch, line0, s);
    }
    OpNS2D op;
    op.vmax = vmax; op.nvmax = 2;
    for (int f = 0; f < 4; ++f) op.in[f] = W[f];
    op.out[0] = W[0];
    op.beta = p->beta;
    return launch_fused(p, op, nlines, scale, nkeep, pitch, line0, s);
}
