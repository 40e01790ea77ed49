// Strided-axis (y / z) FFT passes and the fused first inverse pass (curl / ns2d prologue on load).
#include "internal.h"
#include <stdlib.h>

#include "passes.cuh"

// per-size configuration of the strided pass: E points per thread, TK columns per CTA tile
template <int N> struct SCfg;
template <> struct SCfg<2048> { static constexpr int E = 16, TK = 4; };
template <> struct SCfg<1024> { static constexpr int E = 16, TK = 4; };
template <> struct SCfg<512>  { static constexpr int E = 8,  TK = 4; };
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

template <int N, int E, int TK, int DIR, class L, class S>
static int launch_strided_cfg(Geom g, int nf, L ld, S st, const cplx* tw, cudaStream_t s) {
    constexpr size_t smem = 2 * (size_t)PlaneSize<N>::value * TK * sizeof(double);
    auto kern = fft_strided_kernel<N, E, TK, DIR, L, S>;
    static bool attr_done = false;
    if (!attr_done) {
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_done = true;
    }
    g.nf = nf;
    dim3 grid(((g.ncols + TK - 1) / TK) * nf, g.nouter, 1);
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

template <int N, int DIR, class L, class S>
static int launch_strided_n(Geom g, int nf, L ld, S st, const cplx* tw, cudaStream_t s) {
    if constexpr (N == 1024) {
        switch (strided_variant()) {
            case 1: return launch_strided_cfg<N, 8, 4, DIR>(g, nf, ld, st, tw, s);
            case 2: return launch_strided_cfg<N, 16, 8, DIR>(g, nf, ld, st, tw, s);
            case 3: return launch_strided_cfg<N, 8, 8, DIR>(g, nf, ld, st, tw, s);
            case 4: return launch_strided_cfg<N, 16, 2, DIR>(g, nf, ld, st, tw, s);
        }
    }
    if constexpr (N == 512) {
        switch (strided_variant()) {
            case 1: return launch_strided_cfg<N, 8, 8, DIR>(g, nf, ld, st, tw, s);
            case 2: return launch_strided_cfg<N, 16, 8, DIR>(g, nf, ld, st, tw, s);
            case 3: return launch_strided_cfg<N, 16, 4, DIR>(g, nf, ld, st, tw, s);
            case 4: return launch_strided_cfg<N, 4, 4, DIR>(g, nf, ld, st, tw, s);
        }
    }
    return launch_strided_cfg<N, SCfg<N>::E, SCfg<N>::TK, DIR>(g, nf, ld, st, tw, s);
}

template <int DIR, class L, class S>
static int launch_strided(bool fast, int N, Geom g, int nf, L ld, S st, const cplx* tw, cudaStream_t s) {
    if (fast) {
        switch (N) {
#define B2_CASE(n) case n: return launch_strided_n<n, DIR>(g, nf, ld, st, tw, s);
            B2_CASE(8) B2_CASE(16) B2_CASE(32) B2_CASE(64) B2_CASE(128) B2_CASE(256) B2_CASE(512)
            B2_CASE(1024) B2_CASE(2048)
#undef B2_CASE
        }
    }
    return launch_generic<DIR>(N, g, nf, ld, st, tw, s);
}

static Geom geom_for_axis(const b2_plan* p, int axis) {
    Geom g;
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
    return g;
}

int b2i_strided_plain(b2_plan* p, int axis, int dir, const cplx* const* in, cplx* const* out, int nf,
                      double scale, cudaStream_t s) {
    if (axis == 0 && p->n0 == 1) return 0;
    if (nf > B2_MAXF) return b2i_set_error("too many fields");
    Geom g = geom_for_axis(p, axis);
    const int N = axis == 0 ? p->n0 : p->n1;
    const bool fast = axis == 0 ? p->fast0 : p->fast1;
    const cplx* tw = axis == 0 ? p->tw0 : p->tw1;
    PlainLoad ld;
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

// ------------------------------------------------------------------------------- fused prologues
// ns3d / ns3d.strat: rotfft_from_vecfft_outin (+ Coriolis f on the k=0 mode) computed on load.
// /root/reference/fluidsim/solvers/ns3d/solver.py:199-204, strat/solver.py:154-160.
template <int AXIS>
struct CurlLoad {
    const cplx* in[4];
    const double *k0, *k1, *kx;
    int nk;
    int has_f;
    double f;
    B2_DEVINL cplx operator()(int fld, long long off, int i, int col, int outer) const {
        if (fld < 3) return in[fld][off];
        if (fld == 6) return in[3][off];
        int i0, i1, ikx;
        if (AXIS == 0) {
            i0 = i;
            i1 = col / nk;
            ikx = col - i1 * nk;
        } else {
            i0 = outer;
            i1 = i;
            ikx = col;
        }
        const double Kz = __ldg(k0 + i0), Ky = __ldg(k1 + i1), Kx = __ldg(kx + ikx);
        cplx a, b;
        double ka, kb;
        if (fld == 3) {  // i (Ky vz - Kz vy)
            a = in[2][off]; b = in[1][off]; ka = Ky; kb = Kz;
        } else if (fld == 4) {  // i (Kz vx - Kx vz)
            a = in[0][off]; b = in[2][off]; ka = Kz; kb = Kx;
        } else {  // i (Kx vy - Ky vx)
            a = in[1][off]; b = in[0][off]; ka = Kx; kb = Ky;
        }
        const double tr = ka * a.x - kb * b.x;
        const double ti = ka * a.y - kb * b.y;
        cplx r = make_double2(-ti, tr);
        if (fld == 5 && has_f && i0 == 0 && i1 == 0 && ikx == 0) r.x += f;
        return r;
    }
};

// ns2d: vecfft_from_rotfft + gradfft_from_fft on load.
// /root/reference/fluidsim/solvers/ns2d/solver.py:158,165.
struct Ns2dLoad {
    const cplx* rot;
    const double *k1, *kx;
    B2_DEVINL cplx operator()(int fld, long long off, int i, int col, int outer) const {
        const cplx r = rot[off];
        const double Ky = __ldg(k1 + i), Kx = __ldg(kx + col);
        double K2 = Kx * Kx + Ky * Ky;
        if (i == 0 && col == 0) K2 = 1e-14;
        const double inv = 1.0 / K2;
        double cf;
        switch (fld) {
            case 0: cf = Ky * inv; break;    // ux =  i KY / K2 rot
            case 1: cf = -(Kx * inv); break; // uy = -i KX / K2 rot
            case 2: cf = Kx; break;          // d_x rot = i KX rot
            default: cf = Ky; break;         // d_y rot = i KY rot
        }
        return make_double2(-cf * r.y, cf * r.x);
    }
};

int b2i_first_inverse_pass(b2_plan* p, const cplx* const* in, cplx* const* out, cudaStream_t s) {
    if (p->solver == B2_SOLVER_NS2D) {
        Ns2dLoad ld;
        ld.rot = in[0];
        ld.k1 = p->k1;
        ld.kx = p->kx;
        PlainStore st;
        for (int f = 0; f < 4; ++f) st.out[f] = out[f];
        return launch_strided<+1>(p->fast1, p->n1, geom_for_axis(p, 1), 4, ld, st, p->tw1, s);
    }
    const int nv = p->solver == B2_SOLVER_NS3D_STRAT ? 4 : 3;
    const int nout = nv + 3;
    PlainStore st;
    for (int f = 0; f < nout; ++f) st.out[f] = out[f];
    if (p->n0 > 1) {
        CurlLoad<0> ld;
        for (int f = 0; f < nv; ++f) ld.in[f] = in[f];
        ld.k0 = p->k0; ld.k1 = p->k1; ld.kx = p->kx; ld.nk = p->nk;
        ld.has_f = p->has_f; ld.f = p->f;
        return launch_strided<+1>(p->fast0, p->n0, geom_for_axis(p, 0), nout, ld, st, p->tw0, s);
    }
    CurlLoad<1> ld;
    for (int f = 0; f < nv; ++f) ld.in[f] = in[f];
    ld.k0 = p->k0; ld.k1 = p->k1; ld.kx = p->kx; ld.nk = p->nk;
    ld.has_f = p->has_f; ld.f = p->f;
    return launch_strided<+1>(p->fast1, p->n1, geom_for_axis(p, 1), nout, ld, st, p->tw1, s);
}
