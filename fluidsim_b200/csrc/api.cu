// C ABI of libb200spectral.so: plan management, transforms, k-space / x-space operator kernels,
// the fused tendencies_nonlin and the RK2 / RK4 time step.  See include/b200spectral.h.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <dlfcn.h>

#include <atomic>
#include <vector>

#include "internal.h"

// ------------------------------------------------------------------------------- errors / counters
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int b2i_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return -1;
}
int b2i_check_launch(const char* what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        return b2i_set_error("%s: %s", what, cudaGetErrorString(e));
    }
    return 0;
}
void b2i_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

extern "C" const char* b2_last_error(void) { return g_err; }
extern "C" int b2_version(void) { return 100; }
extern "C" long long b2_launch_count(void) { return g_launches.load(); }

#define CUDA_TRY(expr)                                                              \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) return b2i_set_error(#expr ": %s", cudaGetErrorString(_e)); \
    } while (0)

// ------------------------------------------------------------------------------- plan
static bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

static int upload_twiddles(int n, cplx** out) {
    std::vector<cplx> h(n > 0 ? n : 1);
    const long double tau = 6.283185307179586476925286766559005768L;
    for (int k = 0; k < n; ++k) {
        long double a = tau * (long double)k / (long double)n;
        h[k] = make_double2((double)cosl(a), (double)-sinl(a));
    }
    CUDA_TRY(cudaMalloc((void**)out, sizeof(cplx) * h.size()));
    CUDA_TRY(cudaMemcpy(*out, h.data(), sizeof(cplx) * h.size(), cudaMemcpyHostToDevice));
    return 0;
}

static int upload_wavenumbers(int n, int nkeep, double L, bool half, double** out, int first = 0,
                              int stride = 1) {
    // fluidfft k_adim ordering: [0..n/2, -n/2+1..-1] (np.fft.fftfreq*n with +n/2 for even n)
    std::vector<double> h(nkeep > 0 ? nkeep : 1, 0.0);
    const double dk = L > 0 ? 2.0 * M_PI / L : 0.0;
    for (int j = 0; j < nkeep; ++j) {
        const int i = first + j * stride;
        int k = i;
        if (!half && i > n / 2) k = i - n;
        h[j] = dk * (double)k;
    }
    CUDA_TRY(cudaMalloc((void**)out, sizeof(double) * h.size()));
    CUDA_TRY(cudaMemcpy(*out, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int b2_plan_create(b2_plan** out, int ndim, int n0, int n1, int n2, double L0, double L1,
                              double L2) {
    if (!out) return b2i_set_error("b2_plan_create: out is NULL");
    if (const char* gr = getenv("B2_L2GRAN")) {  // tuning knob: L2 fetch granularity hint (32/64/128 B)
        cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(gr));
    }
    if (ndim != 2 && ndim != 3) return b2i_set_error("b2_plan_create: ndim must be 2 or 3");
    b2_plan* p = new b2_plan();
    memset(p, 0, sizeof(*p));
    p->ndim = ndim;
    if (ndim == 2) {
        p->n0 = 1; p->n1 = n0; p->n2 = n1;
        p->L0 = 0.0; p->L1 = L0; p->L2 = L1;
    } else {
        p->n0 = n0; p->n1 = n1; p->n2 = n2;
        p->L0 = L0; p->L1 = L1; p->L2 = L2;
    }
    if (p->n0 < 1 || p->n1 < 1 || p->n2 < 2) {
        delete p;
        return b2i_set_error("b2_plan_create: bad shape (%d, %d, %d)", n0, n1, n2);
    }
    p->nk = p->n2 / 2 + 1;
    p->fast0 = is_pow2(p->n0) && p->n0 >= 8 && p->n0 <= 2048;
    p->fast1 = is_pow2(p->n1) && p->n1 >= 8 && p->n1 <= 2048;
    p->fast2 = is_pow2(p->n2) && p->n2 >= 8 && p->n2 <= 2048;
    int e = 0;
    e |= upload_twiddles(p->n0, &p->tw0);
    e |= upload_twiddles(p->n1, &p->tw1);
    e |= upload_twiddles(p->n2, &p->tw2);
    e |= upload_wavenumbers(p->n0, p->n0, p->L0, false, &p->k0);
    e |= upload_wavenumbers(p->n1, p->n1, p->L1, false, &p->k1);
    e |= upload_wavenumbers(p->n2, p->nk, p->L2, true, &p->kx);
    if (e) {
        b2_plan_destroy(p);
        return -1;
    }
    p->solver = -1;
    *out = p;
    return 0;
}

// Slab-decomposed plan (one rank of `nranks`): X space is split along z, K space along ky with the
// local K layout (ny_loc, nz, nx/2+1), dimX_K = (1, 0, 2) -- the layout of fluidfft's
// fft3d.mpi_with_fftwmpi3d that fluidsim already handles
// (/root/reference/fluidsim/operators/operators3d.py:384-391).
void b2i_slab_local_band(const b2_plan* p, int r, int* lo, int* hi) {
    const int nyl = p->nyl, P = p->nranks;
    const int glo = p->prune ? p->gyk_lo : p->gy, ghi = p->prune ? p->gyk_hi : p->gy;
    auto clampi = [&](long long v) { return (int)(v < 0 ? 0 : (v > nyl ? nyl : v)); };
    int l, h;
    if (p->ky_cyclic) {  // global row = yl * P + r
        auto ceil_div = [&](long long a) { return a <= 0 ? 0 : (a + P - 1) / P; };
        l = clampi(ceil_div((long long)glo - r));
        h = clampi(ceil_div((long long)ghi - r));
    } else {  // global row = r * nyl + yl
        l = clampi((long long)glo - (long long)r * nyl);
        h = clampi((long long)ghi - (long long)r * nyl);
    }
    if (h <= l) l = h = nyl;
    *lo = l;
    *hi = h;
}

extern "C" int b2_plan_create_slab(b2_plan** out, int nz, int ny, int nx, double Lz, double Ly, double Lx,
                                   int rank, int nranks, int ky_cyclic) {
    if (!out) return b2i_set_error("b2_plan_create_slab: out is NULL");
    if (nranks < 1 || rank < 0 || rank >= nranks) return b2i_set_error("b2_plan_create_slab: bad rank");
    if (nranks > 8) return b2i_set_error("b2_plan_create_slab: at most 8 ranks (one node)");
    if (nz % nranks || ny % nranks)
        return b2i_set_error("b2_plan_create_slab: nz=%d and ny=%d must be multiples of nranks=%d", nz, ny,
                             nranks);
    b2_plan* p = new b2_plan();
    memset(p, 0, sizeof(*p));
    p->ndim = 3;
    p->slab = true;
    p->rank = rank;
    p->nranks = nranks;
    p->gy = ny;
    p->nyl = ny / nranks;
    p->nzl = nz / nranks;
    p->n0 = p->nyl;  // local K layout (ny_loc, nz, nk)
    p->n1 = nz;
    p->n2 = nx;
    p->L0 = Ly; p->L1 = Lz; p->L2 = Lx;
    p->nk = nx / 2 + 1;
    p->fast0 = false;
    p->fast1 = is_pow2(nz) && nz >= 8 && nz <= 2048;
    p->fast2 = is_pow2(nx) && nx >= 8 && nx <= 2048;
    p->fasty = is_pow2(ny) && ny >= 8 && ny <= 2048;
    p->slab_nc = 1;
    p->ky_cyclic = ky_cyclic ? 1 : 0;
    int e = 0;
    e |= upload_twiddles(1, &p->tw0);
    e |= upload_twiddles(nz, &p->tw1);
    e |= upload_twiddles(nx, &p->tw2);
    e |= upload_twiddles(ny, &p->twy);
    if (ky_cyclic) e |= upload_wavenumbers(ny, p->nyl, Ly, false, &p->k0, rank, nranks);
    else e |= upload_wavenumbers(ny, p->nyl, Ly, false, &p->k0, rank * p->nyl);
    e |= upload_wavenumbers(nz, nz, Lz, false, &p->k1);
    e |= upload_wavenumbers(nx, p->nk, Lx, true, &p->kx);
    if (e) {
        b2_plan_destroy(p);
        return -1;
    }
    p->solver = -1;
    *out = p;
    return 0;
}

extern "C" int b2_slab_comm_destroy(b2_plan* p);
extern "C" int b2_plan_destroy(b2_plan* p) {
    if (!p) return 0;
    b2_slab_comm_destroy(p);
    if (p->streams_ready) {
        cudaStreamDestroy(p->sy1); cudaStreamDestroy(p->sx); cudaStreamDestroy(p->sy2);
        cudaEventDestroy(p->ev_begin); cudaEventDestroy(p->ev_end);
        for (int i = 0; i < 32; ++i) { cudaEventDestroy(p->ev_y[i]); cudaEventDestroy(p->ev_x[i]); }
    }
    cudaFree(p->twy);
    cudaFree(p->tw0); cudaFree(p->tw1); cudaFree(p->tw2);
    cudaFree(p->k0); cudaFree(p->k1); cudaFree(p->kx);
    delete p;
    return 0;
}

extern "C" int b2_plan_shapes(const b2_plan* p, int* shapeX, int* shapeK) {
    shapeX[0] = p->n0; shapeX[1] = p->n1; shapeX[2] = p->n2;
    shapeK[0] = p->n0; shapeK[1] = p->n1; shapeK[2] = p->nk;
    return 0;
}

extern "C" int b2_plan_is_fast(const b2_plan* p) {
    if (p->slab) return p->fast1 && p->fast2 && p->fasty;
    return (p->n0 == 1 || p->fast0) && p->fast1 && p->fast2;
}

// ------------------------------------------------------------------------------- transforms
extern "C" int b2_fft_r2c(b2_plan* p, const double* X, double* K, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    cplx* k = reinterpret_cast<cplx*>(K);
    const double scale = 1.0 / ((double)p->n0 * p->n1 * p->n2);
    int e = b2i_xpass_r2c(p, X, k, scale, s);
    if (e) return e;
    const cplx* in[1] = {k};
    cplx* out[1] = {k};
    if ((e = b2i_strided_plain(p, 1, -1, in, out, 1, 1.0, s))) return e;
    return b2i_strided_plain(p, 0, -1, in, out, 1, 1.0, s);
}

extern "C" int b2_ifft_c2r(b2_plan* p, const double* K, double* X, double* work, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const cplx* k = reinterpret_cast<const cplx*>(K);
    cplx* w = work ? reinterpret_cast<cplx*>(work) : const_cast<cplx*>(k);
    const cplx* in[1] = {k};
    cplx* out[1] = {w};
    const cplx* inw[1] = {w};
    int e;
    if (p->n0 > 1) {
        if ((e = b2i_strided_plain(p, 0, +1, in, out, 1, 1.0, s))) return e;
        if ((e = b2i_strided_plain(p, 1, +1, inw, out, 1, 1.0, s))) return e;
    } else {
        if ((e = b2i_strided_plain(p, 1, +1, in, out, 1, 1.0, s))) return e;
    }
    return b2i_xpass_c2r(p, w, X, s);
}

// ------------------------------------------------------------------------------- pointwise kernels
// One CTA per (i0, i1) row, threads run over kx: wavenumbers along axes 0 / 1 are per-block
// constants, accesses are coalesced 16-byte elements.
struct KGrid {
    const double *k0, *k1, *kx;
    int n0, n1, nk;
    int swap01;      // slab K layout (ky_loc, kz, kx): axis 0 carries ky, axis 1 carries kz
    int has_origin;  // the k = 0 mode lives on this rank (row 0)
    // visited rows / columns (dealias-pruned fused path; identity otherwise): compact row index
    // (j0, j1) -> (i0, i1) by j < lo ? j : j + gap ; kx < nkx
    int n1c, r0_lo, r0_gap, r1_lo, r1_gap, nkx;
};
static KGrid kgrid(const b2_plan* p) {
    KGrid g;
    g.k0 = p->k0; g.k1 = p->k1; g.kx = p->kx;
    g.n0 = p->n0; g.n1 = p->n1; g.nk = p->nk;
    g.swap01 = p->slab ? 1 : 0;
    g.has_origin = (!p->slab || p->rank == 0) ? 1 : 0;
    g.n1c = p->n1; g.r0_lo = 1 << 30; g.r0_gap = 0; g.r1_lo = 1 << 30; g.r1_gap = 0; g.nkx = p->nk;
    return g;
}
// grid of the fused path: only the bounding box of the non-dealiased modes when pruning is on
static KGrid kgrid_fused(const b2_plan* p) {
    KGrid g = kgrid(p);
    if (p->prune) {
        g.r0_lo = p->keep0_lo; g.r0_gap = p->keep0_hi - p->keep0_lo;
        g.r1_lo = p->keep1_lo; g.r1_gap = p->keep1_hi - p->keep1_lo;
        g.n1c = p->keep1_lo + (p->n1 - p->keep1_hi);
        g.nkx = p->keepx;
    }
    return g;
}
static inline unsigned nrows_fused(const b2_plan* p) {
    if (!p->prune) return (unsigned)((long long)p->n0 * p->n1);
    return (unsigned)((long long)(p->keep0_lo + (p->n0 - p->keep0_hi)) * (p->keep1_lo + (p->n1 - p->keep1_hi)));
}
#define B2_ROW_SETUP                                                     \
    const long long row = blockIdx.x;                                    \
    const int j0_ = (int)(row / g.n1c);                                  \
    const int j1_ = (int)(row - (long long)j0_ * g.n1c);                 \
    const int i0 = j0_ < g.r0_lo ? j0_ : j0_ + g.r0_gap;                 \
    const int i1 = j1_ < g.r1_lo ? j1_ : j1_ + g.r1_gap;                 \
    const double K0v_ = g.k0[i0], K1v_ = g.k1[i1];                       \
    const double Kz = g.swap01 ? K1v_ : K0v_;                            \
    const double Ky = g.swap01 ? K0v_ : K1v_;                            \
    const bool row_origin = g.has_origin && i0 == 0 && i1 == 0;          \
    (void)row_origin;                                                    \
    const long long rbase = ((long long)i0 * g.n1 + i1) * g.nk;
static inline unsigned nrows(const b2_plan* p) { return (unsigned)((long long)p->n0 * p->n1); }
#define B2_ROW_THREADS 128

// omega = i k x v  (rotfft_from_vecfft_outin); `f` is added to omega_z at k = 0
// (_modif_omegafft_with_f, /root/reference/fluidsim/solvers/ns3d/solver.py:176-178)
B2_DEVINL void curl3(double Kx, double Ky, double Kz, cplx a, cplx b, cplx c, cplx& rx, cplx& ry, cplx& rz) {
    rx = make_double2(-(Ky * c.y - Kz * b.y), Ky * c.x - Kz * b.x);
    ry = make_double2(-(Kz * a.y - Kx * c.y), Kz * a.x - Kx * c.x);
    rz = make_double2(-(Kx * b.y - Ky * a.y), Kx * b.x - Ky * a.x);
}
__global__ void rot_kernel(KGrid g, const cplx* vx, const cplx* vy, const cplx* vz, cplx* rx, cplx* ry,
                           cplx* rz, double f) {
    B2_ROW_SETUP
    for (int ikx = threadIdx.x; ikx < g.nkx; ikx += blockDim.x) {
        const double Kx = g.kx[ikx];
        const long long i = rbase + ikx;
        cplx ox, oy, oz;
        curl3(Kx, Ky, Kz, vx[i], vy[i], vz[i], ox, oy, oz);
        if (row_origin && ikx == 0) oz.x += f;
        rx[i] = ox; ry[i] = oy; rz[i] = oz;
    }
}

__global__ void div_kernel(KGrid g, const cplx* vx, const cplx* vy, const cplx* vz, cplx* d) {
    B2_ROW_SETUP
    for (int ikx = threadIdx.x; ikx < g.nkx; ikx += blockDim.x) {
        const double Kx = g.kx[ikx];
        const long long i = rbase + ikx;
        const cplx a = vx[i], b = vy[i], c = vz[i];
        const double tr = Kx * a.x + Ky * b.x + Kz * c.x;
        const double ti = Kx * a.y + Ky * b.y + Kz * c.y;
        d[i] = make_double2(-ti, tr);
    }
}

B2_DEVINL double inv_k2_nozero(double K2, bool is_origin) { return 1.0 / (is_origin ? 1e-14 : K2); }

B2_DEVINL void project3(double Kx, double Ky, double Kz, double invK2, cplx& a, cplx& b, cplx& c) {
    // project_perpk3d: tmp = (Kx vx + Ky vy + Kz vz) * inv_K_square_nozero ; v -= K tmp
    const double tr = (Kx * a.x + Ky * b.x + Kz * c.x) * invK2;
    const double ti = (Kx * a.y + Ky * b.y + Kz * c.y) * invK2;
    a.x -= Kx * tr; a.y -= Kx * ti;
    b.x -= Ky * tr; b.y -= Ky * ti;
    c.x -= Kz * tr; c.y -= Kz * ti;
}

// project_toroidal / project_poloidal (/root/reference/fluidsim/operators/operators3d.py:911-958,
// 788-856), same operation order as the reference's array expressions
B2_DEVINL void project_toroidal3(double Kx, double Ky, cplx& a, cplx& b, cplx& c) {
    double Kh2 = Kx * Kx + Ky * Ky;
    if (Kh2 == 0.0) Kh2 = 1e-14;
    const double tmp = sqrt(1.0 / Kh2);
    const double cphi = Kx * tmp, sphi = Ky * tmp;
    const double tr = -sphi * a.x + cphi * b.x, ti = -sphi * a.y + cphi * b.y;
    a = make_double2(-sphi * tr, -sphi * ti);
    b = make_double2(cphi * tr, cphi * ti);
    c = make_double2(0.0, 0.0);
}
B2_DEVINL void project_poloidal3(double Kx, double Ky, double Kz, cplx& a, cplx& b, cplx& c) {
    const double Kh2 = Kx * Kx + Ky * Ky;
    double K2nz = Kh2 + Kz * Kz, Kh2nz = Kh2;
    if (Kh2nz == 0.0) Kh2nz = 1e-14;
    if (K2nz == 0.0) K2nz = 1e-14;
    const double invKh = 1.0 / Kh2nz, invK = 1.0 / K2nz;
    const double cth = Kz * sqrt(invK), sth = sqrt(Kh2 * invK);
    const double cphi = Kx * sqrt(invKh), sphi = Ky * sqrt(invKh);
    const double cc = cth * cphi, cs = cth * sphi;
    const double tr = cc * a.x + cs * b.x - sth * c.x, ti = cc * a.y + cs * b.y - sth * c.y;
    a = make_double2(cc * tr, cc * ti);
    b = make_double2(cs * tr, cs * ti);
    c = make_double2(-sth * tr, -sth * ti);
}
// params.projection dispatch (solvers/ns3d/solver.py:158-174)
B2_DEVINL void project_any(int projection, double Kx, double Ky, double Kz, double invK2, cplx& a, cplx& b, cplx& c) {
    if (projection == 0) project3(Kx, Ky, Kz, invK2, a, b, c);
    else if (projection == 1) project_toroidal3(Kx, Ky, a, b, c);
    else project_poloidal3(Kx, Ky, Kz, a, b, c);
}

__global__ void project_tp_kernel(KGrid g, cplx* vx, cplx* vy, cplx* vz, int projection) {
    B2_ROW_SETUP
    for (int ikx = threadIdx.x; ikx < g.nkx; ikx += blockDim.x) {
        const double Kx = g.kx[ikx];
        const long long i = rbase + ikx;
        cplx a = vx[i], b = vy[i], c = vz[i];
        project_any(projection, Kx, Ky, Kz, 0.0, a, b, c);
        vx[i] = a; vy[i] = b; vz[i] = c;
    }
}

__global__ void project_kernel(KGrid g, cplx* vx, cplx* vy, cplx* vz) {
    B2_ROW_SETUP
    for (int ikx = threadIdx.x; ikx < g.nkx; ikx += blockDim.x) {
        const double Kx = g.kx[ikx];
        const long long i = rbase + ikx;
        cplx a = vx[i], b = vy[i], c = vz[i];
        const double K2 = Kx * Kx + Ky * Ky + Kz * Kz;
        project3(Kx, Ky, Kz, inv_k2_nozero(K2, row_origin && ikx == 0), a, b, c);
        vx[i] = a; vy[i] = b; vz[i] = c;
    }
}

__global__ void dealias_kernel(cplx* f, long long fsize, int nvar, const uint8_t* mask) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= fsize) return;
    if (mask[i]) {
        for (int v = 0; v < nvar; ++v) f[v * fsize + i] = make_double2(0.0, 0.0);
    }
}

__global__ void vecprod_kernel(const double* ax, const double* ay, const double* az, double* bx,
                               double* by, double* bz, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double a0 = ax[i], a1 = ay[i], a2 = az[i], b0 = bx[i], b1 = by[i], b2 = bz[i];
    bx[i] = a1 * b2 - a2 * b1;
    by[i] = a2 * b0 - a0 * b2;
    bz[i] = a0 * b1 - a1 * b0;
}

__global__ void mul_real_kernel(const double* a, const double* b, double* o, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = a[i] * b[i];
}

__global__ void frot_kernel(const double* ux, const double* uy, const double* px, const double* py,
                            double beta, double* o, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    o[i] = beta == 0.0 ? -ux[i] * px[i] - uy[i] * py[i] : -ux[i] * px[i] - uy[i] * (py[i] + beta);
}

// tendencies_nonlin_ns2dstrat / _ns2dbouss (solvers/ns2d/strat/solver.py:21-27, bouss/solver.py:21-27)
__global__ void ns2d_buoyancy_kernel(const double* ux, const double* uy, const double* px_rot,
                                     const double* py_rot, const double* px_b, const double* py_b, double N2,
                                     int bouss, double* f_rot, double* f_b, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double u = ux[i], v = uy[i], pxb = px_b[i], pyb = py_b[i];
    const double adv_rot = -u * px_rot[i] - v * py_rot[i];
    const double adv_b = -u * pxb - v * pyb;
    f_rot[i] = bouss ? adv_rot + pxb : adv_rot;
    f_b[i] = bouss ? adv_b : adv_b - N2 * v;
}

__global__ void fb_kernel(cplx* d, double N2, const cplx* vz, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const cplx a = d[i], b = vz[i];
    d[i] = make_double2(-a.x - N2 * b.x, -a.y - N2 * b.y);
}

__global__ void add_kernel(cplx* a, const cplx* b, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    cplx x = a[i];
    const cplx y = b[i];
    x.x += y.x; x.y += y.y;
    a[i] = x;
}

// 2-D operators (plans stored with n0 == 1: Ky = k1, Kx = kx)
__global__ void vec_from_rot2d_kernel(KGrid g, const cplx* rot, cplx* ux, cplx* uy) {
    B2_ROW_SETUP
    (void)Kz;
    for (int ikx = threadIdx.x; ikx < g.nkx; ikx += blockDim.x) {
        const double Kx = g.kx[ikx];
        const long long i = rbase + ikx;
        double K2 = Kx * Kx + Ky * Ky;
        if (row_origin && ikx == 0) K2 = 1e-14;
        const double inv = 1.0 / K2;
        const cplx r = rot[i];
        const double cy = Ky * inv, cx = Kx * inv;
        ux[i] = make_double2(-cy * r.y, cy * r.x);
        uy[i] = make_double2(cx * r.y, -(cx * r.x));
    }
}
__global__ void grad2d_kernel(KGrid g, const cplx* f, cplx* px, cplx* py) {
    B2_ROW_SETUP
    (void)Kz;
    for (int ikx = threadIdx.x; ikx < g.nkx; ikx += blockDim.x) {
        const double Kx = g.kx[ikx];
        const long long i = rbase + ikx;
        const cplx r = f[i];
        px[i] = make_double2(-Kx * r.y, Kx * r.x);
        py[i] = make_double2(-Ky * r.y, Ky * r.x);
    }
}
__global__ void rot2d_kernel(KGrid g, const cplx* ux, const cplx* uy, cplx* rot) {
    B2_ROW_SETUP
    (void)Kz;
    for (int ikx = threadIdx.x; ikx < g.nkx; ikx += blockDim.x) {
        const double Kx = g.kx[ikx];
        const long long i = rbase + ikx;
        const cplx a = ux[i], b = uy[i];
        rot[i] = make_double2(-(Kx * b.y - Ky * a.y), Kx * b.x - Ky * a.x);
    }
}

#define B2_1D_GRID(n) (unsigned)(((n) + 255) / 256), 256

extern "C" int b2_rotfft_from_vecfft(b2_plan* p, const double* vx, const double* vy, const double* vz,
                                     double* rx, double* ry, double* rz, void* stream) {
    rot_kernel<<<nrows(p), B2_ROW_THREADS, 0, (cudaStream_t)stream>>>(
        kgrid(p), (const cplx*)vx, (const cplx*)vy, (const cplx*)vz, (cplx*)rx, (cplx*)ry, (cplx*)rz, 0.0);
    B2_LAUNCH_CHECK("rot_kernel");
    return 0;
}
extern "C" int b2_divfft_from_vecfft(b2_plan* p, const double* vx, const double* vy, const double* vz,
                                     double* d, void* stream) {
    div_kernel<<<nrows(p), B2_ROW_THREADS, 0, (cudaStream_t)stream>>>(
        kgrid(p), (const cplx*)vx, (const cplx*)vy, (const cplx*)vz, (cplx*)d);
    B2_LAUNCH_CHECK("div_kernel");
    return 0;
}
extern "C" int b2_project_perpk3d(b2_plan* p, double* vx, double* vy, double* vz, void* stream) {
    project_kernel<<<nrows(p), B2_ROW_THREADS, 0, (cudaStream_t)stream>>>(kgrid(p), (cplx*)vx, (cplx*)vy,
                                                                           (cplx*)vz);
    B2_LAUNCH_CHECK("project_kernel");
    return 0;
}
extern "C" int b2_project_toroidal(b2_plan* p, double* vx, double* vy, double* vz, void* stream) {
    project_tp_kernel<<<nrows(p), B2_ROW_THREADS, 0, (cudaStream_t)stream>>>(kgrid(p), (cplx*)vx, (cplx*)vy,
                                                                              (cplx*)vz, 1);
    B2_LAUNCH_CHECK("project_tp_kernel");
    return 0;
}
extern "C" int b2_project_poloidal(b2_plan* p, double* vx, double* vy, double* vz, void* stream) {
    project_tp_kernel<<<nrows(p), B2_ROW_THREADS, 0, (cudaStream_t)stream>>>(kgrid(p), (cplx*)vx, (cplx*)vy,
                                                                              (cplx*)vz, 2);
    B2_LAUNCH_CHECK("project_tp_kernel");
    return 0;
}
extern "C" int b2_vector_product(const double* ax, const double* ay, const double* az, double* bx,
                                 double* by, double* bz, long long n, void* stream) {
    vecprod_kernel<<<B2_1D_GRID(n), 0, (cudaStream_t)stream>>>(ax, ay, az, bx, by, bz, n);
    B2_LAUNCH_CHECK("vecprod_kernel");
    return 0;
}
extern "C" int b2_mul_real(const double* a, const double* b, double* out, long long n, void* stream) {
    mul_real_kernel<<<B2_1D_GRID(n), 0, (cudaStream_t)stream>>>(a, b, out, n);
    B2_LAUNCH_CHECK("mul_real_kernel");
    return 0;
}
extern "C" int b2_dealias(b2_plan* p, double* fields, int nvar, const uint8_t* mask, void* stream) {
    const long long n = p->fsize();
    dealias_kernel<<<B2_1D_GRID(n), 0, (cudaStream_t)stream>>>((cplx*)fields, n, nvar, mask);
    B2_LAUNCH_CHECK("dealias_kernel");
    return 0;
}
extern "C" int b2_vecfft_from_rotfft2d(b2_plan* p, const double* rot, double* ux, double* uy,
                                       void* stream) {
    vec_from_rot2d_kernel<<<nrows(p), B2_ROW_THREADS, 0, (cudaStream_t)stream>>>(
        kgrid(p), (const cplx*)rot, (cplx*)ux, (cplx*)uy);
    B2_LAUNCH_CHECK("vec_from_rot2d_kernel");
    return 0;
}
extern "C" int b2_gradfft_from_fft2d(b2_plan* p, const double* f, double* px, double* py, void* stream) {
    grad2d_kernel<<<nrows(p), B2_ROW_THREADS, 0, (cudaStream_t)stream>>>(kgrid(p), (const cplx*)f,
                                                                          (cplx*)px, (cplx*)py);
    B2_LAUNCH_CHECK("grad2d_kernel");
    return 0;
}
extern "C" int b2_rotfft_from_vecfft2d(b2_plan* p, const double* ux, const double* uy, double* rot,
                                       void* stream) {
    rot2d_kernel<<<nrows(p), B2_ROW_THREADS, 0, (cudaStream_t)stream>>>(kgrid(p), (const cplx*)ux,
                                                                         (const cplx*)uy, (cplx*)rot);
    B2_LAUNCH_CHECK("rot2d_kernel");
    return 0;
}
extern "C" int b2_compute_frot(const double* ux, const double* uy, const double* px, const double* py,
                               double beta, double* out, long long n, void* stream) {
    frot_kernel<<<B2_1D_GRID(n), 0, (cudaStream_t)stream>>>(ux, uy, px, py, beta, out, n);
    B2_LAUNCH_CHECK("frot_kernel");
    return 0;
}
extern "C" int b2_tendencies_ns2d_buoyancy(const double* ux, const double* uy, const double* px_rot,
                                          const double* py_rot, const double* px_b, const double* py_b,
                                          double N, int bouss, double* f_rot, double* f_b, long long n,
                                          void* stream) {
    ns2d_buoyancy_kernel<<<B2_1D_GRID(n), 0, (cudaStream_t)stream>>>(ux, uy, px_rot, py_rot, px_b, py_b, N * N,
                                                                     bouss, f_rot, f_b, n);
    B2_LAUNCH_CHECK("ns2d_buoyancy_kernel");
    return 0;
}
extern "C" int b2_compute_fb_fft(double* div_vb, double N, const double* vz, long long nk, void* stream) {
    fb_kernel<<<B2_1D_GRID(nk), 0, (cudaStream_t)stream>>>((cplx*)div_vb, N * N, (const cplx*)vz, nk);
    B2_LAUNCH_CHECK("fb_kernel");
    return 0;
}
extern "C" int b2_add_inplace(double* a, const double* b, long long nk, void* stream) {
    add_kernel<<<B2_1D_GRID(nk), 0, (cudaStream_t)stream>>>((cplx*)a, (const cplx*)b, nk);
    B2_LAUNCH_CHECK("add_kernel");
    return 0;
}

// ------------------------------------------------------------------------------- linear term
struct Visc {
    double nu2, nu4, nu8, num4;
    double k2_hypo_origin;  // K2 at index [0,0,1] (2-D: [0,1]) used to patch the K=0 mode
};
static Visc visc_of(const b2_plan* p, double nu2, double nu4, double nu8, double num4) {
    const double dkx = 2.0 * M_PI / p->L2;
    return Visc{nu2, nu4, nu8, num4, dkx * dkx};
}
// compute_freq_diss, /root/reference/fluidsim/base/solvers/pseudo_spect.py:161-189
B2_DEVINL double freq_diss(const Visc& v, double K2, bool is_origin) {
    double fd = v.nu2 > 0.0 ? v.nu2 * K2 : 0.0;
    const double K4 = K2 * K2;
    if (v.nu4 > 0.0) fd += v.nu4 * K4;
    if (v.nu8 > 0.0) fd += v.nu8 * (K4 * K4);
    if (v.num4 != 0.0) {
        const double k2n = is_origin ? v.k2_hypo_origin : K2;
        fd += v.num4 / (k2n * k2n);
    }
    return fd;
}

__global__ void exact_coefs_kernel(KGrid g, Visc v, double dt, double* exact, double* exact2) {
    B2_ROW_SETUP
    for (int ikx = threadIdx.x; ikx < g.nkx; ikx += blockDim.x) {
        const double Kx = g.kx[ikx];
        const double K2 = Kx * Kx + Ky * Ky + Kz * Kz;
        const double fd = freq_diss(v, K2, row_origin && ikx == 0);
        exact[rbase + ikx] = exp(-dt * fd);
        exact2[rbase + ikx] = exp(-dt / 2 * fd);
    }
}
extern "C" int b2_exact_coefs(b2_plan* p, double nu2, double nu4, double nu8, double num4, double dt,
                              double* exact, double* exact2, void* stream) {
    exact_coefs_kernel<<<nrows(p), B2_ROW_THREADS, 0, (cudaStream_t)stream>>>(
        kgrid(p), visc_of(p, nu2, nu4, nu8, num4), dt, exact, exact2);
    B2_LAUNCH_CHECK("exact_coefs_kernel");
    return 0;
}

// elementwise RK kernels with explicit diss arrays (operator-level API)
template <int MODE>
__global__ void rk_elem_kernel(long long fsize, int nvar, cplx* S, cplx* acc, cplx* O, const cplx* T,
                               const double* diss, const double* diss2, double dt) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= fsize) return;
    const double d1 = diss ? diss[i] : 1.0, d2 = diss2 ? diss2[i] : 1.0;
    for (int v = 0; v < nvar; ++v) {
        const long long j = v * fsize + i;
        const cplx t = T[j];
        if (MODE == 0) {  // step_Euler: O = (S + dt T) diss
            const cplx s = S[j];
            O[j] = make_double2((s.x + dt * t.x) * d1, (s.y + dt * t.y) * d1);
        } else if (MODE == 1) {  // step_like_RK2: S = S diss + dt diss2 T
            const cplx s = S[j];
            const double c = dt * d2;
            S[j] = make_double2(s.x * d1 + c * t.x, s.y * d1 + c * t.y);
        } else if (MODE == 2) {  // rk4_step1: acc += dt/3 diss2 T ; O = S diss2 + dt/2 T
            const cplx s = S[j];
            cplx a = acc[j];
            const double c = dt / 3 * d2;
            a.x += c * t.x; a.y += c * t.y;
            acc[j] = a;
            const double h = dt / 2;
            O[j] = make_double2(s.x * d2 + h * t.x, s.y * d2 + h * t.y);
        } else if (MODE == 3) {  // rk4_step2: acc += dt/3 diss2 T ; O = S diss + dt diss2 T
            const cplx s = S[j];
            cplx a = acc[j];
            const double c = dt / 3 * d2;
            a.x += c * t.x; a.y += c * t.y;
            acc[j] = a;
            const double c2 = dt * d2;
            O[j] = make_double2(s.x * d1 + c2 * t.x, s.y * d1 + c2 * t.y);
        } else {  // rk4_step3: S = acc + dt/6 T
            const cplx a = acc[j];
            const double c = dt / 6;
            S[j] = make_double2(a.x + c * t.x, a.y + c * t.y);
        }
    }
}

extern "C" int b2_step_euler(b2_plan* p, const double* S, double dt, const double* T, const double* diss,
                             double* out, int nvar, void* stream) {
    const long long n = p->fsize();
    rk_elem_kernel<0><<<B2_1D_GRID(n), 0, (cudaStream_t)stream>>>(n, nvar, (cplx*)S, nullptr, (cplx*)out,
                                                                  (const cplx*)T, diss, nullptr, dt);
    B2_LAUNCH_CHECK("rk_elem_kernel<0>");
    return 0;
}
extern "C" int b2_step_like_rk2(b2_plan* p, double* S, double dt, const double* T, const double* diss,
                                const double* diss2, int nvar, void* stream) {
    const long long n = p->fsize();
    rk_elem_kernel<1><<<B2_1D_GRID(n), 0, (cudaStream_t)stream>>>(n, nvar, (cplx*)S, nullptr, nullptr,
                                                                  (const cplx*)T, diss, diss2, dt);
    B2_LAUNCH_CHECK("rk_elem_kernel<1>");
    return 0;
}
extern "C" int b2_rk4_step1(b2_plan* p, const double* S, double* acc, double* S12, const double* T,
                            const double* diss2, double dt, int nvar, void* stream) {
    const long long n = p->fsize();
    rk_elem_kernel<2><<<B2_1D_GRID(n), 0, (cudaStream_t)stream>>>(n, nvar, (cplx*)S, (cplx*)acc, (cplx*)S12,
                                                                  (const cplx*)T, nullptr, diss2, dt);
    B2_LAUNCH_CHECK("rk_elem_kernel<2>");
    return 0;
}
extern "C" int b2_rk4_step2(b2_plan* p, const double* S, double* acc, double* S1, const double* T,
                            const double* diss, const double* diss2, double dt, int nvar, void* stream) {
    const long long n = p->fsize();
    rk_elem_kernel<3><<<B2_1D_GRID(n), 0, (cudaStream_t)stream>>>(n, nvar, (cplx*)S, (cplx*)acc, (cplx*)S1,
                                                                  (const cplx*)T, diss, diss2, dt);
    B2_LAUNCH_CHECK("rk_elem_kernel<3>");
    return 0;
}
extern "C" int b2_rk4_step3(b2_plan* p, double* S, const double* acc, const double* T, double dt, int nvar,
                            void* stream) {
    const long long n = p->fsize();
    rk_elem_kernel<4><<<B2_1D_GRID(n), 0, (cudaStream_t)stream>>>(n, nvar, (cplx*)S, (cplx*)acc, nullptr,
                                                                  (const cplx*)T, nullptr, nullptr, dt);
    B2_LAUNCH_CHECK("rk_elem_kernel<4>");
    return 0;
}

// ------------------------------------------------------------------------------- reductions
B2_DEVINL double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
B2_DEVINL double warp_max(double v) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}
B2_DEVINL void atomic_max_double(double* addr, double v) {  // v >= 0
    atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

// sum_wavenumbers(|f|^2): weight 1 for kx=0 and (even nx) kx=nx/2, else 2
__global__ void sumk_abs2_kernel(KGrid g, const cplx* f, long long fsize, int nvar, int nx_even,
                                 double* out) {
    __shared__ double sh[B2_ROW_THREADS / 32];
    const long long rbase = (long long)blockIdx.x * g.nk;
    double acc = 0.0;
    for (int ikx = threadIdx.x; ikx < g.nk; ikx += blockDim.x) {
        const double w = (ikx == 0 || (nx_even && ikx == g.nk - 1)) ? 1.0 : 2.0;
        double e = 0.0;
        for (int v = 0; v < nvar; ++v) {
            const cplx a = f[v * fsize + rbase + ikx];
            e += a.x * a.x + a.y * a.y;
        }
        acc += w * e;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < B2_ROW_THREADS / 32; ++i) t += sh[i];
        atomicAdd(out, t);
    }
}
extern "C" int b2_sum_wavenumbers_abs2(b2_plan* p, const double* fields, int nvar, double* out_dev,
                                       void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(out_dev, 0, sizeof(double), s));
    sumk_abs2_kernel<<<nrows(p), B2_ROW_THREADS, 0, s>>>(kgrid(p), (const cplx*)fields, p->fsize(), nvar,
                                                         p->n2 % 2 == 0, out_dev);
    B2_LAUNCH_CHECK("sumk_abs2_kernel");
    return 0;
}

template <int OP>  // 0: max |x| ; 1: sum
__global__ void reduce_kernel(const double* x, long long n, double* out) {
    __shared__ double sh[8];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const double v = x[i];
        if (OP == 0) {
            // NaN-propagating max so that a blown-up field is visible
            acc = (v != v) ? v : fmax(acc, fabs(v));
        } else {
            acc += v;
        }
    }
    acc = OP == 0 ? warp_max(acc) : warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = sh[0];
        for (int i = 1; i < 8; ++i) t = OP == 0 ? fmax(t, sh[i]) : t + sh[i];
        if (OP == 0) atomic_max_double(out, t);
        else atomicAdd(out, t);
    }
}
extern "C" int b2_max_abs(const double* x, long long n, double* out_dev, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(out_dev, 0, sizeof(double), s));
    const unsigned grid = (unsigned)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    reduce_kernel<0><<<grid ? grid : 1, 256, 0, s>>>(x, n, out_dev);
    B2_LAUNCH_CHECK("reduce_kernel<max>");
    return 0;
}
extern "C" int b2_sum(const double* x, long long n, double* out_dev, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(out_dev, 0, sizeof(double), s));
    const unsigned grid = (unsigned)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    reduce_kernel<1><<<grid ? grid : 1, 256, 0, s>>>(x, n, out_dev);
    B2_LAUNCH_CHECK("reduce_kernel<sum>");
    return 0;
}


// ------------------------------------------------------------------------------- observables
// One pass over the state for everything the reference's periodic outputs reduce from it:
// component energies, dissipation rates, enstrophy (SpatialMeansNS3D._save_one_time,
// /root/reference/fluidsim/solvers/ns3d/output/spatial_means.py:23-73), the shell-binned 3-D spectra
// and the 1-D spectra of every component (SpectraNS3D.compute, output/spectra.py:15-60, through
// fluidfft's compute_3dspectrum / compute_1dspectra [EXT]: r2c weights, linear sharing between
// adjacent shells, +-k folded in the 1-D spectra).  A few persistent CTAs walk over the (i0, i1)
// rows; histograms live in shared memory and are flushed once per CTA with atomicAdd.
//
// out (doubles):  [0..3]  E of variable 0..3 (sum' |a|^2 / 2)        [4] epsK   [5] epsK_hypo
//                 [6] epsK4   [7] epsK8   [8] enstrophy (sum' |k x v|^2 / 2, 3-D)   [9..15] reserved
//                 then spec3d[nvar][nks], s_kx[nvar][nkx1], s_ky[nvar][nky1], s_kz[nvar][nkz1]
//                 (already divided by deltak / deltakx / deltaky / deltakz)
#define B2_OBS_SCALARS 16
struct ObsArgs {
    KGrid g;
    Visc visc;
    const cplx* S;
    long long fsize;
    int nvar, nks, nkx1, nky1, nkz1, nx_even, is3d;
    double inv_dk, inv_dkx, inv_dky, inv_dkz;
    long long nrows;
    double* out;
};
__global__ void __launch_bounds__(B2_ROW_THREADS) observables_kernel(ObsArgs a) {
    extern __shared__ double obs_sm[];
    const KGrid& g = a.g;
    const int nv = a.nvar;
    double* sp3 = obs_sm;                       // [nv][nks]
    double* skx = sp3 + (size_t)nv * a.nks;     // [nv][nkx1]
    double* sky = skx + (size_t)nv * a.nkx1;    // [nv][nky1]
    double* skz = sky + (size_t)nv * a.nky1;    // [nv][nkz1]
    const int ntot = nv * (a.nks + a.nkx1 + a.nky1 + a.nkz1);
    for (int i = threadIdx.x; i < ntot; i += blockDim.x) obs_sm[i] = 0.0;
    __syncthreads();
    double sc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (long long row = blockIdx.x; row < a.nrows; row += gridDim.x) {
        const int i0 = (int)(row / g.n1), i1 = (int)(row - (long long)i0 * g.n1);
        const double K0v = g.k0[i0], K1v = g.k1[i1];
        const double Kz = a.is3d ? (g.swap01 ? K1v : K0v) : 0.0;
        const double Ky = g.swap01 ? K0v : K1v;
        const bool row_origin = g.has_origin && i0 == 0 && i1 == 0;
        const long long rbase = ((long long)i0 * g.n1 + i1) * g.nk;
        const int iky = (int)lrint(fabs(Ky) * a.inv_dky), ikz = (int)lrint(fabs(Kz) * a.inv_dkz);
        double rowE[4] = {0, 0, 0, 0};
        for (int ikx = threadIdx.x; ikx < g.nk; ikx += blockDim.x) {
            const double Kx = g.kx[ikx];
            const double w = (ikx == 0 || (a.nx_even && ikx == g.nk - 1)) ? 1.0 : 2.0;
            const double K2 = Kx * Kx + Ky * Ky + Kz * Kz;
            const double kappa = sqrt(K2) * a.inv_dk;
            int ik = (int)floor(kappa);
            double share = kappa - ik;
            if (ik >= a.nks - 1) { ik = a.nks - 1; share = 0.0; }
            cplx v3[3] = {make_double2(0, 0), make_double2(0, 0), make_double2(0, 0)};
            double etot = 0.0;
            for (int v = 0; v < nv; ++v) {
                const cplx q = a.S[v * a.fsize + rbase + ikx];
                if (v < 3) v3[v] = q;
                const double e = 0.5 * w * (q.x * q.x + q.y * q.y);
                if (e != 0.0) {
                    atomicAdd(&sp3[v * a.nks + ik], (1.0 - share) * e);
                    if (share != 0.0) atomicAdd(&sp3[v * a.nks + ik + 1], share * e);
                    skx[v * a.nkx1 + ikx] += e;  // slot owned by this thread
                }
                rowE[v] += e;
                sc[v] += e;
                if (v < 3) etot += e;
            }
            const bool origin = row_origin && ikx == 0;
            const Visc& vs = a.visc;
            // compute_freq_diss splits f_d (hyper) and f_d_hypo (base/solvers/pseudo_spect.py:161-189)
            double fd = vs.nu2 > 0.0 ? vs.nu2 * K2 : 0.0;
            const double K4 = K2 * K2;
            if (vs.nu4 > 0.0) { fd += vs.nu4 * K4; sc[6] += vs.nu4 * K4 * 2 * etot; }
            if (vs.nu8 > 0.0) { fd += vs.nu8 * K4 * K4; sc[7] += vs.nu8 * K4 * K4 * 2 * etot; }
            sc[4] += fd * 2 * etot;
            if (vs.num4 != 0.0) {
                const double k2n = origin ? vs.k2_hypo_origin : K2;
                sc[5] += vs.num4 / (k2n * k2n) * 2 * etot;
            }
            if (a.is3d && nv >= 3) {
                cplx ox, oy, oz;
                curl3(Kx, Ky, Kz, v3[0], v3[1], v3[2], ox, oy, oz);
                sc[8] += 0.5 * w * (ox.x * ox.x + ox.y * ox.y + oy.x * oy.x + oy.y * oy.y + oz.x * oz.x + oz.y * oz.y);
            }
        }
        for (int v = 0; v < nv; ++v) {
            const double r = warp_sum(rowE[v]);
            if ((threadIdx.x & 31) == 0 && r != 0.0) {
                atomicAdd(&sky[v * a.nky1 + iky], r);
                if (a.is3d) atomicAdd(&skz[v * a.nkz1 + ikz], r);
            }
        }
    }
    __syncthreads();
    double* o = a.out + B2_OBS_SCALARS;
    for (int i = threadIdx.x; i < ntot; i += blockDim.x) {
        const double val = obs_sm[i];
        if (val != 0.0) {
            double scale;
            const int j = i;
            if (j < nv * a.nks) scale = a.inv_dk;
            else if (j < nv * (a.nks + a.nkx1)) scale = a.inv_dkx;
            else if (j < nv * (a.nks + a.nkx1 + a.nky1)) scale = a.inv_dky;
            else scale = a.inv_dkz;
            atomicAdd(&o[i], val * scale);
        }
    }
    for (int k = 0; k < 9; ++k) {
        const double r = warp_sum(sc[k]);
        if ((threadIdx.x & 31) == 0 && r != 0.0) atomicAdd(&a.out[k], r);
    }
}

extern "C" long long b2_observables_size(const b2_plan* p, int nvar, int nks) {
    const int gy = p->slab ? p->gy : p->n1, gz = p->slab ? p->n1 : p->n0;
    return B2_OBS_SCALARS + (long long)nvar * (nks + p->nk + (gy / 2 + 1) + (p->ndim == 3 ? gz / 2 + 1 : 1));
}

extern "C" int b2_observables(b2_plan* p, const double* S, int nvar, int nks, double deltak, double* out_dev,
                              void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (nvar < 1 || nvar > 4) return b2i_set_error("b2_observables: nvar must be in 1..4");
    if (nks < 2 || deltak <= 0.0) return b2i_set_error("b2_observables: bad shell description");
    ObsArgs a;
    a.g = kgrid(p);
    a.visc = visc_of(p, p->nu2, p->nu4, p->nu8, p->num4);
    a.S = (const cplx*)S;
    a.fsize = p->fsize();
    a.nvar = nvar;
    a.nks = nks;
    const int gy = p->slab ? p->gy : p->n1, gz = p->slab ? p->n1 : p->n0;
    const double Ly = p->slab ? p->L0 : p->L1, Lz = p->slab ? p->L1 : p->L0;
    a.is3d = p->ndim == 3;
    a.nkx1 = p->nk;
    a.nky1 = gy / 2 + 1;
    a.nkz1 = a.is3d ? gz / 2 + 1 : 1;
    a.nx_even = p->n2 % 2 == 0;
    a.inv_dk = 1.0 / deltak;
    a.inv_dkx = p->L2 / (2.0 * M_PI);
    a.inv_dky = Ly / (2.0 * M_PI);
    a.inv_dkz = a.is3d ? Lz / (2.0 * M_PI) : 1.0;
    a.nrows = (long long)p->n0 * p->n1;
    a.out = out_dev;
    const long long ntot = b2_observables_size(p, nvar, nks);
    CUDA_TRY(cudaMemsetAsync(out_dev, 0, sizeof(double) * ntot, s));
    const size_t smem = sizeof(double) * (size_t)(ntot - B2_OBS_SCALARS);
    if (smem > 200 * 1024) return b2i_set_error("b2_observables: histograms do not fit shared memory");
    if (smem > 48 * 1024) {
        cudaError_t ce = cudaFuncSetAttribute(observables_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ce != cudaSuccess) return b2i_set_error("observables_kernel: %s", cudaGetErrorString(ce));
    }
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    long long grid = (long long)nsm * (smem > 100 * 1024 ? 1 : 2);
    if (grid > a.nrows) grid = a.nrows;
    observables_kernel<<<(unsigned)grid, B2_ROW_THREADS, smem, s>>>(a);
    B2_LAUNCH_CHECK("observables_kernel");
    return 0;
}

// ------------------------------------------------------------------------------- profiling hooks
// Optional per-kernel-class timing with CUDA events on the launching stream (bench.py's roofline
// leg).  Disabled by default: zero overhead unless b2_profile_enable(1) was called.
enum { PC_FIRST_INV = 0, PC_Y_INV, PC_X_FUSED, PC_Y_FWD, PC_Z_FWD, PC_RK, PC_COUNT };
struct ProfRec { int cat; cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof_recs;
static std::vector<cudaEvent_t> g_prof_pool;
static double g_prof_ms[PC_COUNT];
static long long g_prof_n[PC_COUNT];

static cudaEvent_t prof_event() {
    if (!g_prof_pool.empty()) {
        cudaEvent_t e = g_prof_pool.back();
        g_prof_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
struct ProfScope {
    bool on;
    ProfRec r;
    cudaStream_t s;
    ProfScope(int cat, cudaStream_t s_) : on(g_prof_on), s(s_) {
        if (!on) return;
        r.cat = cat;
        r.a = prof_event();
        r.b = prof_event();
        cudaEventRecord(r.a, s);
    }
    ~ProfScope() {
        if (!on) return;
        cudaEventRecord(r.b, s);
        g_prof_recs.push_back(r);
    }
};
extern "C" int b2_profile_enable(int on) {
    g_prof_on = on != 0;
    return 0;
}
extern "C" int b2_profile_reset(void) {
    for (auto& r : g_prof_recs) { g_prof_pool.push_back(r.a); g_prof_pool.push_back(r.b); }
    g_prof_recs.clear();
    for (int i = 0; i < PC_COUNT; ++i) { g_prof_ms[i] = 0.0; g_prof_n[i] = 0; }
    return 0;
}
// synchronises the device, folds pending records; ms[cat], count[cat] for cat < ncat
extern "C" int b2_profile_get(double* ms, long long* count, int ncat) {
    CUDA_TRY(cudaDeviceSynchronize());
    for (auto& r : g_prof_recs) {
        float t = 0.f;
        cudaEventElapsedTime(&t, r.a, r.b);
        g_prof_ms[r.cat] += t;
        g_prof_n[r.cat] += 1;
        g_prof_pool.push_back(r.a);
        g_prof_pool.push_back(r.b);
    }
    g_prof_recs.clear();
    for (int i = 0; i < ncat && i < PC_COUNT; ++i) { ms[i] = g_prof_ms[i]; count[i] = g_prof_n[i]; }
    return 0;
}

// ------------------------------------------------------------------------------- fused path
// per-axis "some mode is kept" flags of the dealiasing mask (bounding box of the kept modes)
__global__ void mask_bounds_kernel(const uint8_t* mask, int n0, int n1, int nk, int* keep0, int* keep1,
                                   int* keepx) {
    const long long row = blockIdx.x;
    const int i0 = (int)(row / n1), i1 = (int)(row % n1);
    bool any = false;
    for (int ikx = threadIdx.x; ikx < nk; ikx += blockDim.x) {
        if (!mask[row * nk + ikx]) {
            any = true;
            keepx[ikx] = 1;
        }
    }
    if (any) {
        keep0[i0] = 1;
        keep1[i1] = 1;
    }
}

static void kept_range(const std::vector<int>& keep, int n, int* lo, int* hi) {
    // [lo, hi) = the single contiguous run of fully dealiased indices around the Nyquist index n/2
    // (the band every truncation shape removes).  Other fully dealiased indices (ky = 0 with
    // NO_KY0, user masks, ...) stay in the visited set [0, lo) U [hi, n), where the per-mode mask
    // zeroes them -- treating the hull of all dealiased indices as one band would skip kept rows.
    *lo = n; *hi = n;
    int c = n / 2;
    if (c >= n || keep[c]) {
        // no band at the Nyquist index: fall back to the longest run (odd sizes, exotic masks)
        int best = 0, bl = n, i = 0;
        while (i < n) {
            if (keep[i]) { ++i; continue; }
            int j = i;
            while (j < n && !keep[j]) ++j;
            if (j - i > best) { best = j - i; bl = i; }
            i = j;
        }
        if (best == 0) return;
        *lo = bl; *hi = bl + best;
        return;
    }
    int l = c, h = c + 1;
    while (l > 0 && !keep[l - 1]) --l;
    while (h < n && !keep[h]) ++h;
    *lo = l; *hi = h;
}

static int compute_mask_bounds(b2_plan* p) {
    p->keep0_lo = p->n0; p->keep0_hi = p->n0;
    p->keep1_lo = p->n1; p->keep1_hi = p->n1;
    p->keepx = p->nk;
    if (!p->mask) return 0;
    int* d = nullptr;
    const int tot = p->n0 + p->n1 + p->nk;
    CUDA_TRY(cudaMalloc((void**)&d, sizeof(int) * tot));
    CUDA_TRY(cudaMemset(d, 0, sizeof(int) * tot));
    mask_bounds_kernel<<<(unsigned)((long long)p->n0 * p->n1), 128>>>(p->mask, p->n0, p->n1, p->nk, d, d + p->n0,
                                                                     d + p->n0 + p->n1);
    B2_LAUNCH_CHECK("mask_bounds_kernel");
    std::vector<int> h(tot);
    CUDA_TRY(cudaMemcpy(h.data(), d, sizeof(int) * tot, cudaMemcpyDeviceToHost));
    cudaFree(d);
    std::vector<int> k0(h.begin(), h.begin() + p->n0), k1(h.begin() + p->n0, h.begin() + p->n0 + p->n1);
    kept_range(k0, p->n0, &p->keep0_lo, &p->keep0_hi);
    kept_range(k1, p->n1, &p->keep1_lo, &p->keep1_hi);
    int kx = 0;
    for (int i = 0; i < p->nk; ++i)
        if (h[p->n0 + p->n1 + i]) kx = i + 1;
    p->keepx = kx > 0 ? kx : 1;
    if (p->n0 == 1) { p->keep0_lo = 1; p->keep0_hi = 1; }
    return 0;
}

// *flag = 1 if any of the nvar fields is non-zero at a dealiased (masked) mode
__global__ void dealiased_check_kernel(const cplx* f, long long fsize, int nvar, const uint8_t* mask,
                                       int* flag) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= fsize || !mask[i]) return;
    for (int v = 0; v < nvar; ++v) {
        const cplx a = f[v * fsize + i];
        if (a.x != 0.0 || a.y != 0.0) {
            *flag = 1;
            return;
        }
    }
}
// Is `fields` (nvar K arrays) exactly zero wherever the mask dealiases?  Writes 0 / 1 to *flag_dev
// (device int).  Lets the host side use the pruned transforms on a state it did not produce.
extern "C" int b2_check_dealiased(b2_plan* p, const double* fields, int nvar, const uint8_t* mask,
                                  int* flag_dev, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (!mask) return b2i_set_error("b2_check_dealiased: no mask");
    CUDA_TRY(cudaMemsetAsync(flag_dev, 0, sizeof(int), s));
    const long long n = p->fsize();
    dealiased_check_kernel<<<B2_1D_GRID(n), 0, s>>>((const cplx*)fields, n, nvar, mask, flag_dev);
    B2_LAUNCH_CHECK("dealiased_check_kernel");
    return 0;
}

// Dealias-pruned transforms: when on, the fused path visits only the bounding box of the modes the
// dealiasing mask keeps (all other modes of the state, the tendencies and every intermediate are
// exact zeros and are neither loaded, transformed nor stored).  Valid only when the state itself is
// dealiased (true after every step); the host side switches it on accordingly.
extern "C" int b2_set_pruning(b2_plan* p, int on) {
    if (on && !p->mask) return b2i_set_error("b2_set_pruning: no dealiasing mask set");
    if (on && p->slab) return b2i_set_error("b2_set_pruning: use b2_slab_set_pruning on slab plans");
    p->prune = on ? 1 : 0;
    return 0;
}
// Slab plans: the kept ranges must be agreed between the ranks (host side, fluidsim_b200/slab.py):
// kx < keepx and the kz band [kz_lo, kz_hi) are global; [yl_lo, yl_hi) is the dealiased band of the
// LOCAL ky rows; [gy_lo, gy_hi) the dealiased band of the global ky index (y passes on the z-slab
// side).  Only the kept local ky rows / kx columns are exchanged.
extern "C" int b2_slab_set_pruning(b2_plan* p, int on, int keepx, int kz_lo, int kz_hi, int yl_lo, int yl_hi,
                                   int gy_lo, int gy_hi) {
    if (!p->slab) return b2i_set_error("b2_slab_set_pruning: not a slab plan");
    if (!on) {
        p->prune = 0;
        return 0;
    }
    if (!p->mask) return b2i_set_error("b2_slab_set_pruning: no dealiasing mask set");
    if (keepx < 1 || keepx > p->nk || kz_lo > kz_hi || kz_hi > p->n1 || yl_lo > yl_hi || yl_hi > p->n0 ||
        gy_lo > gy_hi || gy_hi > p->gy)
        return b2i_set_error("b2_slab_set_pruning: inconsistent ranges");
    p->keepx = keepx;
    p->keep1_lo = kz_lo; p->keep1_hi = kz_hi;
    p->gyk_lo = gy_lo; p->gyk_hi = gy_hi;
    p->prune = 1;
    // the local band follows from the global one and the ky distribution (yl_lo / yl_hi are only
    // cross-checked so that host and library agree on the layout)
    b2i_slab_local_band(p, p->rank, &p->keep0_lo, &p->keep0_hi);
    if (p->keep0_lo != yl_lo || p->keep0_hi != yl_hi) {
        p->prune = 0;
        return b2i_set_error("b2_slab_set_pruning: local band [%d, %d) disagrees with the library's [%d, %d)",
                             yl_lo, yl_hi, p->keep0_lo, p->keep0_hi);
    }
    return 0;
}
// kept local ky rows of every rank for the current pruning state (all-to-all split sizes)
extern "C" int b2_slab_kept_rows(const b2_plan* p, int* nkl) {
    if (!p->slab) return b2i_set_error("b2_slab_kept_rows: not a slab plan");
    for (int r = 0; r < p->nranks; ++r) {
        int lo, hi;
        b2i_slab_local_band(p, r, &lo, &hi);
        nkl[r] = p->nyl - (hi - lo);
    }
    return 0;
}
/* kept index ranges: out[0..4] = keep0_lo, keep0_hi, keep1_lo, keep1_hi, keepx */
extern "C" int b2_get_pruning_bounds(const b2_plan* p, int* out) {
    out[0] = p->keep0_lo; out[1] = p->keep0_hi; out[2] = p->keep1_lo; out[3] = p->keep1_hi; out[4] = p->keepx;
    return 0;
}

extern "C" int b2_set_physics(b2_plan* p, int solver, double nu2, double nu4, double nu8, double num4,
                              int has_f, double f, double N, double beta, const uint8_t* mask) {
    if (solver < 0 || solver > 2) return b2i_set_error("b2_set_physics: unknown solver %d", solver);
    if ((solver == B2_SOLVER_NS2D) != (p->ndim == 2))
        return b2i_set_error("b2_set_physics: solver %d does not match a %d-D plan", solver, p->ndim);
    if (p->slab) {
        // wavenumber-dependent constants use the GLOBAL grid: L0/L1 were stored as (Ly, Lz)
    }
    p->solver = solver;
    p->nu2 = nu2; p->nu4 = nu4; p->nu8 = nu8; p->num4 = num4;
    p->has_f = has_f; p->f = f; p->N = N; p->beta = beta;
    // the kept ranges are recomputed on every push: the caller may have edited the mask in place,
    // or a new mask may live at the address of a freed one
    p->mask = mask;
    p->prune = 0;
    if (p->slab) {  // slab plans: ranges agreed between the ranks (b2_slab_set_pruning)
        p->keep0_lo = p->keep0_hi = p->n0;
        p->keep1_lo = p->keep1_hi = p->n1;
        p->keepx = p->nk;
        return 0;
    }
    return compute_mask_bounds(p);
}

/* params.no_vz_kz0 (solvers/ns3d/solver.py:135-137, 260-263): vz(kz = 0) = 0 (and b) after every
 * projection of the tendencies and of the state */
extern "C" int b2_set_no_vz_kz0(b2_plan* p, int on) {
    if (on && p->ndim != 3) return b2i_set_error("b2_set_no_vz_kz0: 3-D solvers only");
    p->no_vz_kz0 = on ? 1 : 0;
    return 0;
}

/* params.projection (solvers/ns3d/solver.py:139-174): 0 = None (project_perpk3d), 1 = "toroidal" /
 * "vortical" (operators3d.py:911-958), 2 = "poloidal" (operators3d.py:788-856) */
extern "C" int b2_set_projection(b2_plan* p, int projection) {
    if (projection < 0 || projection > 2) return b2i_set_error("b2_set_projection: unknown projection %d", projection);
    if (projection && p->ndim != 3) return b2i_set_error("b2_set_projection: 3-D solvers only");
    p->projection = projection;
    return 0;
}

extern "C" int b2_work_fields(const b2_plan* p, int solver, int* nwork, int* nvar) {
    switch (solver) {
        case B2_SOLVER_NS3D: *nwork = 6; *nvar = 3; return 0;
        case B2_SOLVER_NS3D_STRAT: *nwork = 7; *nvar = 4; return 0;
        case B2_SOLVER_NS2D: *nwork = 4; *nvar = 1; return 0;
    }
    return b2i_set_error("b2_work_fields: unknown solver %d", solver);
}

extern "C" int b2_set_buffers(b2_plan* p, double* acc, double* stage, double* work) {
    p->acc = (cplx*)acc; p->stage = (cplx*)stage; p->work = (cplx*)work;
    return 0;
}

enum { M_TEND = 0, M_RK4_0, M_RK4_1, M_RK4_2, M_RK4_3, M_RK2_0, M_RK2_1 };

struct RKArgs {
    KGrid g;
    Visc visc;
    cplx* W;          // raw FFT output: nout fields, stride fsize
    cplx* Wo;         // 3 fields: receive the vorticity of the next stage input
    const cplx* Sin;  // stage input (strat coupling terms)
    double fcor;      // Coriolis parameter added to omega_z(k=0) (0 if params.f is None)
    cplx* S;          // state (nvar fields)
    cplx* A;          // accumulator
    cplx* B;          // next stage input
    cplx* Tout;       // M_TEND output
    const uint8_t* mask;
    long long fsize;
    double dt, N2;
    int projection;   // params.projection (0 perpk3d, 1 toroidal, 2 poloidal)
    int no_vz_kz0;    // solvers/ns3d/solver.py:260-263
    const double* dt_ptr;  // device-resident time increment (CFL steps); overrides dt when set
};

// Epilogue of a stage: raw FFT(v x omega) -> (+ buoyancy terms) -> Leray projection -> dealiasing
// -> exact-linear RK update.  Replaces project_state_spect + oper.dealiasing
// (/root/reference/fluidsim/solvers/ns3d/solver.py:251-263), compute_fb_fft
// (strat/solver.py:29-33, 198, 206-209) and step_Euler / rk4_step1-3 / step_like_RK2
// (base/time_stepping/pseudo_spect.py:51-68, 902-984, 503-517).
template <int SOLVER, int MODE>
__global__ void __launch_bounds__(B2_ROW_THREADS) rk_stage_kernel(RKArgs a) {
    constexpr int NV = SOLVER == B2_SOLVER_NS3D ? 3 : (SOLVER == B2_SOLVER_NS3D_STRAT ? 4 : 1);
    const KGrid& g = a.g;
    B2_ROW_SETUP
    for (int ikx = threadIdx.x; ikx < g.nkx; ikx += blockDim.x) {
        const double Kx = g.kx[ikx];
        const long long i = rbase + ikx;
        const bool origin = row_origin && ikx == 0;
        const double K2 = Kx * Kx + Ky * Ky + Kz * Kz;
        const double invK2 = inv_k2_nozero(K2, origin);
        const bool masked = a.mask ? a.mask[i] != 0 : false;
        cplx T[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) T[v] = a.W[v * a.fsize + i];
        if (SOLVER == B2_SOLVER_NS3D_STRAT) {
            const cplx b = a.Sin[3 * a.fsize + i], vz = a.Sin[2 * a.fsize + i];
            T[2].x += b.x; T[2].y += b.y;
            const cplx w3 = a.W[3 * a.fsize + i], w4 = a.W[4 * a.fsize + i], w5 = a.W[5 * a.fsize + i];
            const double dr = Kx * w3.x + Ky * w4.x + Kz * w5.x;
            const double di = Kx * w3.y + Ky * w4.y + Kz * w5.y;
            // div = i (dr + i di) = (-di, dr);  fb = -div - N^2 vz
            T[3] = make_double2(di - a.N2 * vz.x, -dr - a.N2 * vz.y);
        }
        if (SOLVER != B2_SOLVER_NS2D) project_any(a.projection, Kx, Ky, Kz, invK2, T[0], T[1], T[2]);
        const bool kz0 = SOLVER != B2_SOLVER_NS2D && a.no_vz_kz0 && Kz == 0.0;
        if (kz0) {  // dealiasing_variable(vz_fft, where_kz_0) (+ b_fft), solver.py:260-263
            T[2] = make_double2(0.0, 0.0);
            if (SOLVER == B2_SOLVER_NS3D_STRAT) T[NV - 1] = make_double2(0.0, 0.0);
        }
        if (masked) {
#pragma unroll
            for (int v = 0; v < NV; ++v) T[v] = make_double2(0.0, 0.0);
        }
        if (MODE == M_TEND) {
#pragma unroll
            for (int v = 0; v < NV; ++v) a.Tout[v * a.fsize + i] = T[v];
            continue;
        }
        const double dt = a.dt_ptr ? *a.dt_ptr : a.dt;
        double E = 1.0, E2 = 1.0;
        if (MODE != M_RK4_3) {
            const double fd = freq_diss(a.visc, K2, origin);
            if (MODE != M_RK4_1 && MODE != M_RK2_0) E = exp(-dt * fd);
            if (MODE != M_RK4_3) E2 = exp(-dt / 2 * fd);
        }
        cplx Sn[NV];  // new state (last stage) or next stage input B (other stages)
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const long long j = v * a.fsize + i;
            const cplx t = T[v];
            if (MODE == M_RK4_0) {
                const cplx s = a.S[j];
                const double c6 = dt / 6, c2 = dt / 2;
                a.A[j] = make_double2((s.x + c6 * t.x) * E, (s.y + c6 * t.y) * E);
                Sn[v] = make_double2((s.x + c2 * t.x) * E2, (s.y + c2 * t.y) * E2);
            } else if (MODE == M_RK4_1) {
                const cplx s = a.S[j];
                cplx ac = a.A[j];
                const double c = dt / 3 * E2, h = dt / 2;
                ac.x += c * t.x; ac.y += c * t.y;
                a.A[j] = ac;
                Sn[v] = make_double2(s.x * E2 + h * t.x, s.y * E2 + h * t.y);
            } else if (MODE == M_RK4_2) {
                const cplx s = a.S[j];
                cplx ac = a.A[j];
                const double c = dt / 3 * E2, c2 = dt * E2;
                ac.x += c * t.x; ac.y += c * t.y;
                a.A[j] = ac;
                Sn[v] = make_double2(s.x * E + c2 * t.x, s.y * E + c2 * t.y);
            } else if (MODE == M_RK4_3) {
                const cplx ac = a.A[j];
                const double c = dt / 6;
                Sn[v] = make_double2(ac.x + c * t.x, ac.y + c * t.y);
            } else if (MODE == M_RK2_0) {
                const cplx s = a.S[j];
                const double h = dt / 2;
                Sn[v] = make_double2((s.x + h * t.x) * E2, (s.y + h * t.y) * E2);
            } else {  // M_RK2_1
                const cplx s = a.S[j];
                const double c2 = dt * E2;
                Sn[v] = make_double2(s.x * E + c2 * t.x, s.y * E + c2 * t.y);
            }
        }
        if (MODE == M_RK4_0 || MODE == M_RK4_1 || MODE == M_RK4_2 || MODE == M_RK2_0) {
            // next stage input, and (3-D) its vorticity for the next first inverse pass
#pragma unroll
            for (int v = 0; v < NV; ++v) a.B[v * a.fsize + i] = Sn[v];
            if (SOLVER != B2_SOLVER_NS2D) {
                cplx ox, oy, oz;
                curl3(Kx, Ky, Kz, Sn[0], Sn[1], Sn[2], ox, oy, oz);
                if (origin) oz.x += a.fcor;
                a.Wo[i] = ox;
                a.Wo[a.fsize + i] = oy;
                a.Wo[2 * a.fsize + i] = oz;
            }
        }
        if (MODE == M_RK4_3 || MODE == M_RK2_1) {
            // end of step: project_state_spect + dealiasing (solvers/ns3d/time_stepping.py:15-16)
            if (SOLVER != B2_SOLVER_NS2D) project_any(a.projection, Kx, Ky, Kz, invK2, Sn[0], Sn[1], Sn[2]);
            if (kz0) {
                Sn[2] = make_double2(0.0, 0.0);
                if (SOLVER == B2_SOLVER_NS3D_STRAT) Sn[NV - 1] = make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int v = 0; v < NV; ++v)
                a.S[v * a.fsize + i] = masked ? make_double2(0.0, 0.0) : Sn[v];
        }
    }
}

template <int SOLVER>
static int launch_rk_stage_s(int mode, const RKArgs& a, unsigned grid, cudaStream_t s) {
    switch (mode) {
#define B2_CASE(m) case m: rk_stage_kernel<SOLVER, m><<<grid, B2_ROW_THREADS, 0, s>>>(a); break;
        B2_CASE(M_TEND) B2_CASE(M_RK4_0) B2_CASE(M_RK4_1) B2_CASE(M_RK4_2) B2_CASE(M_RK4_3)
        B2_CASE(M_RK2_0) B2_CASE(M_RK2_1)
#undef B2_CASE
        default: return b2i_set_error("bad RK mode");
    }
    B2_LAUNCH_CHECK("rk_stage_kernel");
    return 0;
}

// tendencies_fft += forcing_fft (/root/reference/fluidsim/solvers/ns3d/solver.py:243-244,
// ns2d/solver.py:185-186): the forcing of fluidsim's forcing makers lives on a few low-wavenumber
// modes ((2 nkmax_forcing)^3 at most, base/forcing/specific.py:137-345), so it is kept as a sparse
// list and added to the raw transform output in front of the epilogue -- no field-sized pass.
__global__ void forcing_add_kernel(cplx* W, long long fsize, int nvar, long long nm, const long long* idx,
                                   const cplx* val) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nm * nvar) return;
    const int v = (int)(j / nm);
    const long long m = j - (long long)v * nm;
    const cplx f = val[j];
    cplx* dst = W + (long long)v * fsize + idx[m];
    dst->x += f.x;
    dst->y += f.y;
}

extern "C" int b2_set_forcing_sparse(b2_plan* p, long long nmodes, const long long* idx, const double* val,
                                     int nvar) {
    if (nmodes < 0) return b2i_set_error("b2_set_forcing_sparse: negative mode count");
    if (nmodes > 0 && (!idx || !val)) return b2i_set_error("b2_set_forcing_sparse: NULL arrays");
    const int nv_max = p->solver == B2_SOLVER_NS2D ? 1 : 3;
    if (nmodes > 0 && (nvar < 1 || nvar > nv_max))
        return b2i_set_error("b2_set_forcing_sparse: nvar = %d (the forced variables are the %d leading state "
                             "variables; a forced buoyancy is not supported)", nvar, nv_max);
    p->force_n = nmodes;
    p->force_idx = idx;
    p->force_val = (const cplx*)val;
    p->force_nvar = nvar;
    return 0;
}

static int launch_rk_stage(b2_plan* p, int mode, const RKArgs& a, cudaStream_t s) {
    ProfScope ps(PC_RK, s);
    const unsigned grid = nrows_fused(p);
    if (grid == 0) return 0;
    if (p->force_n > 0) {
        const long long tot = p->force_n * p->force_nvar;
        forcing_add_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, s>>>(a.W, a.fsize, p->force_nvar, p->force_n,
                                                                        p->force_idx, p->force_val);
        B2_LAUNCH_CHECK("forcing_add_kernel");
    }
    switch (p->solver) {
        case B2_SOLVER_NS3D: return launch_rk_stage_s<B2_SOLVER_NS3D>(mode, a, grid, s);
        case B2_SOLVER_NS3D_STRAT: return launch_rk_stage_s<B2_SOLVER_NS3D_STRAT>(mode, a, grid, s);
        case B2_SOLVER_NS2D: return launch_rk_stage_s<B2_SOLVER_NS2D>(mode, a, grid, s);
    }
    return b2i_set_error("physics not set");
}

// number of z chunks of the overlapped y-inverse / x / y-forward section (1 = sequential).
// Tuning knob: B2_NCHUNK (0 or 1 disables the overlap).
static int overlap_chunks(const b2_plan* p) {
    static int req = -1;
    if (req < 0) {
        const char* e = getenv("B2_NCHUNK");
        req = e ? atoi(e) : 1;  // measured slower than the sequential passes (profiles/r1_tuning.md)
    }
    if (req <= 1 || p->n0 < 64 || p->n2 < 512) return 1;
    int n = req > 32 ? 32 : req;
    while (n > 1 && p->n0 / n < 8) n /= 2;
    return n;
}

static int ensure_streams(b2_plan* p) {
    if (p->streams_ready) return 0;
    CUDA_TRY(cudaStreamCreateWithFlags(&p->sy1, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&p->sx, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&p->sy2, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&p->ev_begin, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&p->ev_end, cudaEventDisableTiming));
    for (int i = 0; i < 32; ++i) {
        CUDA_TRY(cudaEventCreateWithFlags(&p->ev_y[i], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&p->ev_x[i], cudaEventDisableTiming));
    }
    p->streams_ready = true;
    return 0;
}

// raw nonlinear term of stage input `Sin` into work[0..nout-1] (scaled FFT, not yet projected)
static int nonlinear_raw(b2_plan* p, const cplx* Sin, bool need_curl, cudaStream_t s, double* vmax = nullptr) {
    int nwork, nvar;
    if (b2_work_fields(p, p->solver, &nwork, &nvar)) return -1;
    const long long fs = p->fsize();
    const cplx* in[8];
    cplx* W[8];
    const cplx* Wc[8];
    for (int v = 0; v < nvar; ++v) in[v] = Sin + v * fs;
    for (int f = 0; f < nwork; ++f) Wc[f] = W[f] = p->wfield(f);
    const int nout = p->solver == B2_SOLVER_NS3D ? 3 : (p->solver == B2_SOLVER_NS3D_STRAT ? 6 : 1);
    const double scale = 1.0 / ((double)p->n0 * p->n1 * p->n2);
    int e;
    const bool pr = p->prune != 0;
    const int nkeep = pr ? p->keepx : p->nk;
    if (need_curl && p->solver != B2_SOLVER_NS2D) {
        ProfScope ps(PC_RK, s);
        rot_kernel<<<nrows_fused(p), B2_ROW_THREADS, 0, s>>>(kgrid_fused(p), in[0], in[1], in[2], W[3], W[4],
                                                            W[5], p->has_f ? p->f : 0.0);
        B2_LAUNCH_CHECK("rot_kernel");
    }
    {
        ProfScope ps(PC_FIRST_INV, s);
        if ((e = b2i_first_inverse_pass(p, in, W, s))) return e;
    }
    const int nchunk = overlap_chunks(p);
    if (nchunk > 1) {
        // The x pass is bound by the L1 / FP64 pipes, the y passes by HBM: run them concurrently on
        // different z chunks (three streams, one x-pass CTA + one y-pass CTA co-resident per SM).
        if ((e = ensure_streams(p))) return e;
        CUDA_TRY(cudaEventRecord(p->ev_begin, s));
        CUDA_TRY(cudaStreamWaitEvent(p->sy1, p->ev_begin, 0));
        CUDA_TRY(cudaStreamWaitEvent(p->sx, p->ev_begin, 0));
        CUDA_TRY(cudaStreamWaitEvent(p->sy2, p->ev_begin, 0));
        b2i_xpass_share_sm(true);
        for (int c = 0; c < nchunk; ++c) {
            const int z0 = (int)((long long)p->n0 * c / nchunk), z1 = (int)((long long)p->n0 * (c + 1) / nchunk);
            {
                ProfScope ps(PC_Y_INV, p->sy1);
                if ((e = b2i_strided_plain(p, 1, +1, Wc, W, nwork, 1.0, p->sy1, pr, z0, z1 - z0))) return e;
            }
            CUDA_TRY(cudaEventRecord(p->ev_y[c], p->sy1));
            CUDA_TRY(cudaStreamWaitEvent(p->sx, p->ev_y[c], 0));
            {
                ProfScope ps(PC_X_FUSED, p->sx);
                if ((e = b2i_xpass_fused(p, W, (long long)(z1 - z0) * p->n1, scale, nkeep, p->nk,
                                         (long long)z0 * p->n1, p->sx, vmax)))
                    return e;
            }
            CUDA_TRY(cudaEventRecord(p->ev_x[c], p->sx));
            CUDA_TRY(cudaStreamWaitEvent(p->sy2, p->ev_x[c], 0));
            {
                ProfScope ps(PC_Y_FWD, p->sy2);
                if ((e = b2i_strided_plain(p, 1, -1, Wc, W, nout, 1.0, p->sy2, pr, z0, z1 - z0))) return e;
            }
        }
        b2i_xpass_share_sm(false);
        CUDA_TRY(cudaEventRecord(p->ev_end, p->sy2));
        CUDA_TRY(cudaStreamWaitEvent(s, p->ev_end, 0));
    } else {
        if (p->n0 > 1) {
            ProfScope ps(PC_Y_INV, s);
            if ((e = b2i_strided_plain(p, 1, +1, Wc, W, nwork, 1.0, s, pr))) return e;
        }
        {
            ProfScope ps(PC_X_FUSED, s);
            if ((e = b2i_xpass_fused(p, W, (long long)p->n0 * p->n1, scale, nkeep, p->nk, 0, s, vmax))) return e;
        }
        {
            ProfScope ps(PC_Y_FWD, s);
            if ((e = b2i_strided_plain(p, 1, -1, Wc, W, nout, 1.0, s, pr))) return e;
        }
    }
    if (p->n0 > 1) {
        ProfScope ps(PC_Z_FWD, s);
        if ((e = b2i_strided_plain(p, 0, -1, Wc, W, nout, 1.0, s, pr))) return e;
    }
    return 0;
}

static int check_fused_ready(b2_plan* p, bool need_rk) {
    if (p->solver < 0) return b2i_set_error("b2_set_physics has not been called");
    if (p->slab) return b2i_set_error("slab plans are stepped with b2_slab_phase_a/b/c (see b200spectral.h)");
    if (!b2_plan_is_fast(p))
        return b2i_set_error("fused path needs power-of-two sizes in [8, 2048] (got %d x %d x %d)", p->n0,
                             p->n1, p->n2);
    if (!p->work) return b2i_set_error("b2_set_buffers has not been called (work)");
    if (need_rk && (!p->acc || !p->stage)) return b2i_set_error("b2_set_buffers: acc/stage missing");
    if (p->alias_tw && !p->stage) return b2i_set_error("aliased buffers need the stage buffer");
    return 0;
}

static RKArgs rk_args(b2_plan* p, const cplx* Sin, cplx* S, double dt) {
    RKArgs a;
    a.g = kgrid_fused(p);
    a.visc = visc_of(p, p->nu2, p->nu4, p->nu8, p->num4);
    a.W = p->wfield(0);
    a.Wo = p->wfield(3);
    a.Sin = Sin;
    a.S = S;
    a.A = p->acc;
    a.B = p->stage;
    a.Tout = nullptr;
    a.mask = p->mask;
    a.fsize = p->fsize();
    a.dt = dt;
    a.dt_ptr = nullptr;
    a.N2 = p->N * p->N;
    a.no_vz_kz0 = p->no_vz_kz0;
    a.projection = p->projection;
    a.fcor = p->has_f ? p->f : 0.0;
    return a;
}

extern "C" int b2_tendencies(b2_plan* p, const double* S_in, double* T_out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int e;
    if ((e = check_fused_ready(p, false))) return e;
    const cplx* Sin = (const cplx*)S_in;
    if ((e = nonlinear_raw(p, Sin, true, s))) return e;
    RKArgs a = rk_args(p, Sin, nullptr, 0.0);
    a.Tout = (cplx*)T_out;
    return launch_rk_stage(p, M_TEND, a, s);
}

extern "C" int b2_time_step(b2_plan* p, int scheme, double dt, double* S_, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int e;
    if ((e = check_fused_ready(p, true))) return e;
    cplx* S = (cplx*)S_;
    if (scheme == B2_SCHEME_RK4) {
        const int modes[4] = {M_RK4_0, M_RK4_1, M_RK4_2, M_RK4_3};
        for (int st = 0; st < 4; ++st) {
            const cplx* Sin = st == 0 ? S : p->stage;
            if ((e = nonlinear_raw(p, Sin, st == 0, s))) return e;
            RKArgs a = rk_args(p, Sin, S, dt);
            if ((e = launch_rk_stage(p, modes[st], a, s))) return e;
        }
        return 0;
    }
    if (scheme == B2_SCHEME_RK2) {
        const int modes[2] = {M_RK2_0, M_RK2_1};
        for (int st = 0; st < 2; ++st) {
            const cplx* Sin = st == 0 ? S : p->stage;
            if ((e = nonlinear_raw(p, Sin, st == 0, s))) return e;
            RKArgs a = rk_args(p, Sin, S, dt);
            if ((e = launch_rk_stage(p, modes[st], a, s))) return e;
        }
        return 0;
    }
    return b2i_set_error("Problem name time_scheme (scheme id %d)", scheme);
}

// ------------------------------------------------------------------------------- CFL time step
// compute_time_increment_CLF + _compute_time_increment_CLF_from_tmp
// (/root/reference/fluidsim/base/time_stepping/base.py:320-354) on the device: vmax[] are the
// max |v_i| of the stage-0 physical velocity (side output of the fused x pass).
__global__ void cfl_dt_kernel(double* vmax, double inv_dx, double inv_dy, double inv_dz, double cfl,
                              double dt_max, double* dt) {
    const double tmp = vmax[0] * inv_dx + vmax[1] * inv_dy + vmax[2] * inv_dz;
    const double dt_cfl = tmp > 0.0 ? cfl / tmp : dt_max;
    const double maybe_new_dt = dt_cfl < dt_max ? dt_cfl : dt_max;
    const double normalize_diff = fabs(*dt - maybe_new_dt) / maybe_new_dt;
    if (normalize_diff > 0.02) *dt = maybe_new_dt;
    vmax[0] = vmax[1] = vmax[2] = 0.0;
}

// One step with the CFL time increment decided on the device: dt_dev (in/out, device double) holds
// the current deltat; vmax_dev = 3 device doubles (zeroed by the caller once).  The max |v| come
// from the stage-0 x pass, the new deltat is used by all RK epilogues of the step.
extern "C" int b2_time_step_cfl(b2_plan* p, int scheme, double cfl, double deltat_max, double* dt_dev,
                                double* vmax_dev, double* S_, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int e;
    if ((e = check_fused_ready(p, true))) return e;
    if (!dt_dev || !vmax_dev) return b2i_set_error("b2_time_step_cfl: dt_dev / vmax_dev missing");
    if (scheme != B2_SCHEME_RK4 && scheme != B2_SCHEME_RK2)
        return b2i_set_error("Problem name time_scheme (scheme id %d)", scheme);
    cplx* S = (cplx*)S_;
    const int nst = scheme == B2_SCHEME_RK4 ? 4 : 2;
    const int base_mode = scheme == B2_SCHEME_RK4 ? M_RK4_0 : M_RK2_0;
    for (int st = 0; st < nst; ++st) {
        const cplx* Sin = st == 0 ? S : p->stage;
        if ((e = nonlinear_raw(p, Sin, st == 0, s, st == 0 ? vmax_dev : nullptr))) return e;
        if (st == 0) {
            const double idx = p->n2 / p->L2, idy = p->n1 / p->L1, idz = p->n0 > 1 ? p->n0 / p->L0 : 0.0;
            cfl_dt_kernel<<<1, 1, 0, s>>>(vmax_dev, idx, idy, idz, cfl, deltat_max, dt_dev);
            B2_LAUNCH_CHECK("cfl_dt_kernel");
        }
        RKArgs a = rk_args(p, Sin, S, 0.0);
        a.dt_ptr = dt_dev;
        if ((e = launch_rk_stage(p, base_mode + st, a, s))) return e;
    }
    return 0;
}

// ------------------------------------------------------------------------------- development hooks
// raw strided pass over nf contiguous K fields (micro-benchmarks / tuning only)
extern "C" int b2_dev_strided_pass(b2_plan* p, int axis, int dir, const double* in, double* out, int nf,
                                   void* stream) {
    if (nf < 1 || nf > 8) return b2i_set_error("b2_dev_strided_pass: nf out of range");
    const cplx* i_[8];
    cplx* o_[8];
    const long long fs = p->fsize();
    for (int f = 0; f < nf; ++f) {
        i_[f] = (const cplx*)in + f * fs;
        o_[f] = (cplx*)out + f * fs;
    }
    return b2i_strided_plain(p, axis, dir, i_, o_, nf, 1.0, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------- slab stepping
// One RK stage on a slab plan = phase A, all-to-all, phase B, all-to-all, phase C; the two
// all-to-alls (per field, equal contiguous blocks per peer) are issued by the host side through
// torch.distributed / NCCL between the phases (fluidsim_b200/slab.py).
extern "C" int b2_slab_set_buffers(b2_plan* p, double* xa, double* xb) {
    if (!p->slab) return b2i_set_error("b2_slab_set_buffers: not a slab plan");
    p->xa = (cplx*)xa;
    p->xb = (cplx*)xb;
    return 0;
}

// elements between consecutive fields of xa / xb (default: a full K field).  With the pruned
// exchange a field needs only ny * nz_loc * keepx elements in xa and (kept ky rows) * nz_loc * keepx
// in xb: b2_slab_buffer_need reports both for the current pruning state.
extern "C" int b2_slab_set_buffer_strides(b2_plan* p, long long xa_field, long long xb_field) {
    if (!p->slab) return b2i_set_error("b2_slab_set_buffer_strides: not a slab plan");
    if (xa_field < 0 || xb_field < 0) return b2i_set_error("b2_slab_set_buffer_strides: negative stride");
    p->xa_fs = xa_field;
    p->xb_fs = xb_field;
    return 0;
}
extern "C" int b2_slab_buffer_need(const b2_plan* p, long long* xa_field, long long* xb_field) {
    if (!p->slab) return b2i_set_error("b2_slab_buffer_need: not a slab plan");
    if (p->prune) {
        const long long nyk = p->gy - (p->gyk_hi - p->gyk_lo);
        *xa_field = (long long)p->gy * p->nzl * p->keepx;
        *xb_field = nyk * p->nzl * p->keepx;
    } else {
        *xa_field = *xb_field = p->fsize();
    }
    return 0;
}
// ns3d only: raw transform outputs are written into `stage` and rewritten in place by the epilogue;
// `work` then holds only the 3 vorticity fields (12 instead of 15 K fields on one GPU)
extern "C" int b2_set_aliasing(b2_plan* p, int on) {
    if (on && p->solver != B2_SOLVER_NS3D) return b2i_set_error("b2_set_aliasing: ns3d only (set the physics first)");
    p->alias_tw = on ? 1 : 0;
    return 0;
}

static int slab_ready(b2_plan* p) {
    if (!p->slab) return b2i_set_error("not a slab plan");
    if (p->solver != B2_SOLVER_NS3D && p->solver != B2_SOLVER_NS3D_STRAT)
        return b2i_set_error("slab stepping supports ns3d and ns3d.strat");
    if (!b2_plan_is_fast(p)) return b2i_set_error("slab stepping needs power-of-two sizes in [8, 2048]");
    if (!p->work || !p->xa || !p->xb) return b2i_set_error("slab buffers not set");
    if (p->alias_tw && !p->stage) return b2i_set_error("aliased buffers need the stage buffer");
    long long na, nb;
    b2_slab_buffer_need(p, &na, &nb);
    if (p->xa_stride() < na || p->xb_stride() < nb)
        return b2i_set_error("slab exchange buffers too small for this pruning state (need %lld / %lld elements "
                             "per field, have %lld / %lld)", na, nb, p->xa_stride(), p->xb_stride());
    return 0;
}

// Fine-grained stage pieces (field ranges [f0, f1)) so that the host side can pipeline the
// per-field all-to-alls with the FFT passes of the other fields.
extern "C" int b2_slab_curl(b2_plan* p, const double* S_in, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int e;
    if ((e = slab_ready(p))) return e;
    const long long fs = p->fsize();
    const cplx* Sin = (const cplx*)S_in;
    if (nrows_fused(p) == 0) return 0;
    ProfScope ps(PC_RK, s);
    rot_kernel<<<nrows_fused(p), B2_ROW_THREADS, 0, s>>>(kgrid_fused(p), Sin, Sin + fs, Sin + 2 * fs,
                                                        p->wfield(3), p->wfield(4), p->wfield(5),
                                                        p->has_f ? p->f : 0.0);
    B2_LAUNCH_CHECK("rot_kernel");
    return 0;
}

static int slab_counts(b2_plan* p, int* nin, int* nout) {
    const int nv = p->solver == B2_SOLVER_NS3D_STRAT ? 4 : 3;
    *nin = nv + 3;
    *nout = p->solver == B2_SOLVER_NS3D ? 3 : 6;
    return nv;
}

// z-inverse of work-field slots [f0, f1) (0..2 = v from S_in, 3..5 = omega from work, 6 = b) -> xa
extern "C" int b2_slab_zinv(b2_plan* p, const double* S_in, int f0, int f1, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int e, nin, nout;
    if ((e = slab_ready(p))) return e;
    slab_counts(p, &nin, &nout);
    if (f0 < 0 || f1 > nin || f0 >= f1) return b2i_set_error("b2_slab_zinv: bad field range");
    const long long fs = p->fsize();
    const cplx* Sin = (const cplx*)S_in;
    const cplx* in[8];
    cplx* out[8];
    for (int f = f0; f < f1; ++f) {
        in[f - f0] = f < 3 ? Sin + f * fs : (f < 6 ? p->wfield(f) : Sin + 3 * fs);
        out[f - f0] = p->xa + f * p->xa_stride();
    }
    ProfScope ps(PC_FIRST_INV, s);
    return b2i_slab_zpass(p, +1, in, out, f1 - f0, s);
}

// number of z chunks of the exchange layout (pipelining granularity); nz_loc must be a multiple
extern "C" int b2_slab_set_chunks(b2_plan* p, int nc) {
    if (!p->slab) return b2i_set_error("b2_slab_set_chunks: not a slab plan");
    if (nc < 1 || p->nzl % nc) return b2i_set_error("b2_slab_set_chunks: nz_loc=%d not a multiple of %d", p->nzl, nc);
    p->slab_nc = nc;
    return 0;
}

// y-inverse of fields [f0, f1), z chunk `chunk` (-1: all): xb (exchanged, rank-grouped kept rows) ->
// xa (natural (zc, ny, pitch) array of the chunk)
extern "C" int b2_slab_yinv(b2_plan* p, int f0, int f1, int chunk, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int e, nin, nout;
    if ((e = slab_ready(p))) return e;
    slab_counts(p, &nin, &nout);
    if (f0 < 0 || f1 > nin || f0 >= f1) return b2i_set_error("b2_slab_yinv: bad field range");
    if (chunk >= p->slab_nc) return b2i_set_error("b2_slab_yinv: bad chunk");
    const long long fs = p->fsize();
    const cplx* in[8];
    cplx* out[8];
    for (int f = f0; f < f1; ++f) {
        in[f - f0] = p->xb + f * p->xb_stride();
        out[f - f0] = p->xa + f * p->xa_stride();
    }
    ProfScope ps(PC_Y_INV, s);
    for (int c = (chunk < 0 ? 0 : chunk); c < (chunk < 0 ? p->slab_nc : chunk + 1); ++c)
        if ((e = b2i_slab_ypass(p, +1, in, out, f1 - f0, c, s))) return e;
    return 0;
}

extern "C" int b2_slab_xpass(b2_plan* p, int chunk, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int e, nin, nout;
    if ((e = slab_ready(p))) return e;
    slab_counts(p, &nin, &nout);
    if (chunk >= p->slab_nc) return b2i_set_error("b2_slab_xpass: bad chunk");
    const long long fs = p->fsize();
    cplx* XW[8];
    for (int f = 0; f < nin; ++f) XW[f] = p->xa + f * p->xa_stride();
    const double scale = 1.0 / ((double)p->gy * p->n1 * p->n2);
    const int pitch = p->prune ? p->keepx : p->nk;
    const long long lines_c = (long long)p->gy * (p->nzl / p->slab_nc);
    ProfScope ps(PC_X_FUSED, s);
    if (chunk < 0) return b2i_xpass_fused(p, XW, lines_c * p->slab_nc, scale, pitch, pitch, 0, s);
    return b2i_xpass_fused(p, XW, lines_c, scale, pitch, pitch, lines_c * chunk, s);
}

// y-forward of output fields [f0, f1), z chunk `chunk` (-1: all): xa (natural) -> xb (exchanged)
extern "C" int b2_slab_yfwd(b2_plan* p, int f0, int f1, int chunk, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int e, nin, nout;
    if ((e = slab_ready(p))) return e;
    slab_counts(p, &nin, &nout);
    if (f0 < 0 || f1 > nout || f0 >= f1) return b2i_set_error("b2_slab_yfwd: bad field range");
    if (chunk >= p->slab_nc) return b2i_set_error("b2_slab_yfwd: bad chunk");
    const long long fs = p->fsize();
    const cplx* in[8];
    cplx* out[8];
    for (int f = f0; f < f1; ++f) {
        in[f - f0] = p->xa + f * p->xa_stride();
        out[f - f0] = p->xb + f * p->xb_stride();
    }
    ProfScope ps(PC_Y_FWD, s);
    for (int c = (chunk < 0 ? 0 : chunk); c < (chunk < 0 ? p->slab_nc : chunk + 1); ++c)
        if ((e = b2i_slab_ypass(p, -1, in, out, f1 - f0, c, s))) return e;
    return 0;
}

// z-forward of output fields [f0, f1): xa (exchange layout) -> work (K layout)
extern "C" int b2_slab_zfwd(b2_plan* p, int f0, int f1, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int e, nin, nout;
    if ((e = slab_ready(p))) return e;
    slab_counts(p, &nin, &nout);
    if (f0 < 0 || f1 > nout || f0 >= f1) return b2i_set_error("b2_slab_zfwd: bad field range");
    const long long fs = p->fsize();
    const cplx* in[8];
    cplx* out[8];
    for (int f = f0; f < f1; ++f) { in[f - f0] = p->xa + f * p->xa_stride(); out[f - f0] = p->wfield(f); }
    ProfScope ps(PC_Z_FWD, s);
    return b2i_slab_zpass(p, -1, in, out, f1 - f0, s);
}

// RK epilogue on work[0..nout-1] (stage < 0: tendencies only, written to T_out)
extern "C" int b2_slab_rk(b2_plan* p, int scheme, int stage, double dt, const double* S_in, double* S_,
                          double* T_out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int e;
    if ((e = slab_ready(p))) return e;
    int mode;
    if (stage < 0) mode = M_TEND;
    else if (scheme == B2_SCHEME_RK4 && stage < 4) mode = M_RK4_0 + stage;
    else if (scheme == B2_SCHEME_RK2 && stage < 2) mode = M_RK2_0 + stage;
    else return b2i_set_error("b2_slab_rk: bad scheme/stage %d/%d", scheme, stage);
    if (mode != M_TEND && (!p->acc || !p->stage)) return b2i_set_error("b2_set_buffers: acc/stage missing");
    RKArgs a = rk_args(p, (const cplx*)S_in, (cplx*)S_, dt);
    a.Tout = (cplx*)T_out;
    return launch_rk_stage(p, mode, a, s);
}

// Coarse phases (A, B, C) = the pieces above without pipelining.
extern "C" int b2_slab_phase_a(b2_plan* p, const double* S_in, int need_curl, void* stream) {
    int e, nin, nout;
    if ((e = slab_ready(p))) return e;
    slab_counts(p, &nin, &nout);
    if (need_curl && (e = b2_slab_curl(p, S_in, stream))) return e;
    return b2_slab_zinv(p, S_in, 0, nin, stream);
}
extern "C" int b2_slab_phase_b(b2_plan* p, void* stream) {
    int e, nin, nout;
    if ((e = slab_ready(p))) return e;
    slab_counts(p, &nin, &nout);
    if ((e = b2_slab_yinv(p, 0, nin, -1, stream))) return e;
    if ((e = b2_slab_xpass(p, -1, stream))) return e;
    return b2_slab_yfwd(p, 0, nout, -1, stream);
}
extern "C" int b2_slab_phase_c(b2_plan* p, int scheme, int stage, double dt, const double* S_in, double* S_,
                               double* T_out, void* stream) {
    int e, nin, nout;
    if ((e = slab_ready(p))) return e;
    slab_counts(p, &nin, &nout);
    if ((e = b2_slab_zfwd(p, 0, nout, stream))) return e;
    return b2_slab_rk(p, scheme, stage, dt, S_in, S_, T_out, stream);
}

// ------------------------------------------------------------------------------- native collectives
// The two global transposes of every 3-D FFT as grouped ncclSend / ncclRecv issued by the library
// itself (the communicator belongs to the plan), so that a host in any
// language drives the multi-GPU path through this C ABI alone: one call per time step.  NCCL is
// bound at run time (dlopen of the libnccl.so.2 already loaded by the host process, e.g. PyTorch's),
// so the library neither links against it nor needs it for single-GPU use.
// Replaces: the MPI all-to-alls inside fluidfft's fft3d.mpi_with_fftwmpi3d transforms and the
// allreduce of _compute_time_increment_CLF_uxuyuz (/root/reference/fluidsim/base/time_stepping/base.py:320-354).
typedef struct { char internal[128]; } b2_nccl_uid;
struct NcclApi {
    void* handle;
    int (*GetUniqueId)(b2_nccl_uid*);
    int (*CommInitRank)(void**, int, b2_nccl_uid, int);
    int (*CommDestroy)(void*);
    int (*GroupStart)();
    int (*GroupEnd)();
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t);
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t);
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
    const char* (*GetErrorString)(int);
};
enum { B2_NCCL_FLOAT64 = 8, B2_NCCL_SUM = 0, B2_NCCL_MAX = 2 };  // ncclDouble, ncclSum, ncclMax (nccl.h)
static NcclApi* nccl_api() {
    static NcclApi api;
    static int state = 0;  // 0 untried, 1 ok, -1 unavailable
    if (state == 0) {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        state = -1;
        if (h) {
            api.handle = h;
#define B2_SYM(field, name) *(void**)(&api.field) = dlsym(h, name)
            B2_SYM(GetUniqueId, "ncclGetUniqueId");
            B2_SYM(CommInitRank, "ncclCommInitRank");
            B2_SYM(CommDestroy, "ncclCommDestroy");
            B2_SYM(GroupStart, "ncclGroupStart");
            B2_SYM(GroupEnd, "ncclGroupEnd");
            B2_SYM(Send, "ncclSend");
            B2_SYM(Recv, "ncclRecv");
            B2_SYM(AllReduce, "ncclAllReduce");
            B2_SYM(GetErrorString, "ncclGetErrorString");
#undef B2_SYM
            if (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GroupStart && api.GroupEnd &&
                api.Send && api.Recv && api.AllReduce && api.GetErrorString)
                state = 1;
        }
    }
    return state == 1 ? &api : nullptr;
}
#define NCCL_TRY(expr)                                                                       \
    do {                                                                                     \
        int _r = (expr);                                                                     \
        if (_r != 0) return b2i_set_error(#expr ": %s", nccl_api()->GetErrorString(_r));     \
    } while (0)

/* 128-byte NCCL unique id (rank 0 creates it, the host broadcasts it to the other ranks) */
extern "C" int b2_nccl_unique_id(void* out128) {
    NcclApi* n = nccl_api();
    if (!n) return b2i_set_error("b2_nccl_unique_id: libnccl.so.2 not found");
    NCCL_TRY(n->GetUniqueId((b2_nccl_uid*)out128));
    return 0;
}
extern "C" int b2_slab_comm_init(b2_plan* p, const void* uid128) {
    if (!p->slab) return b2i_set_error("b2_slab_comm_init: not a slab plan");
    NcclApi* n = nccl_api();
    if (!n) return b2i_set_error("b2_slab_comm_init: libnccl.so.2 not found");
    if (p->comm_ready) return 0;
    b2_nccl_uid uid;
    memcpy(&uid, uid128, sizeof(uid));
    NCCL_TRY(n->CommInitRank(&p->nccl_comm, p->nranks, uid, p->rank));
    CUDA_TRY(cudaStreamCreateWithFlags(&p->comm_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 128; ++i) CUDA_TRY(cudaEventCreateWithFlags(&p->ev_comm[i], cudaEventDisableTiming));
    p->comm_ready = true;
    return 0;
}
extern "C" int b2_slab_comm_destroy(b2_plan* p) {
    if (!p || !p->comm_ready) return 0;
    cudaStreamSynchronize(p->comm_stream);
    nccl_api()->CommDestroy(p->nccl_comm);
    cudaStreamDestroy(p->comm_stream);
    for (int i = 0; i < 128; ++i) cudaEventDestroy(p->ev_comm[i]);
    p->comm_ready = false;
    return 0;
}

struct SlabExchange {
    long long mine;              // complex elements sent to (K side) / received from every peer
    long long theirs[8];         // complex elements of peer r on the z-slab side (<= 8 ranks, one node)
    long long theirs_off[8];
    long long cs_a, cs_b;        // chunk strides of xa / xb
};
static SlabExchange slab_exchange(const b2_plan* p) {
    SlabExchange ex;
    const long long pitch = p->prune ? p->keepx : p->nk;
    const long long row = (long long)(p->nzl / p->slab_nc) * pitch;
    long long tot = 0;
    for (int r = 0; r < p->nranks; ++r) {
        int lo, hi;
        b2i_slab_local_band(p, r, &lo, &hi);
        const long long nkl = p->nyl - (hi - lo);
        ex.theirs[r] = nkl * row;
        ex.theirs_off[r] = tot;
        tot += ex.theirs[r];
        if (r == p->rank) ex.mine = nkl * row;
    }
    ex.cs_a = (long long)p->gy * row;
    ex.cs_b = tot;
    return ex;
}
// inverse direction: xa (K side, equal blocks per peer) -> xb (z-slab side, rows grouped by owner);
// forward: the opposite.  One grouped send/recv per (field, chunk) on the plan's comm stream.
static int slab_a2a(b2_plan* p, const SlabExchange& ex, int f, int c, bool inverse) {
    NcclApi* n = nccl_api();
    cplx* a = p->xa + (long long)f * p->xa_stride() + (long long)c * ex.cs_a;
    cplx* b = p->xb + (long long)f * p->xb_stride() + (long long)c * ex.cs_b;
    NCCL_TRY(n->GroupStart());
    for (int r = 0; r < p->nranks; ++r) {
        cplx* ablk = a + (long long)r * ex.mine;
        cplx* bblk = b + ex.theirs_off[r];
        if (inverse) {
            if (ex.mine > 0) NCCL_TRY(n->Send(ablk, (size_t)ex.mine * 2, B2_NCCL_FLOAT64, r, p->nccl_comm, p->comm_stream));
            if (ex.theirs[r] > 0) NCCL_TRY(n->Recv(bblk, (size_t)ex.theirs[r] * 2, B2_NCCL_FLOAT64, r, p->nccl_comm, p->comm_stream));
        } else {
            if (ex.theirs[r] > 0) NCCL_TRY(n->Send(bblk, (size_t)ex.theirs[r] * 2, B2_NCCL_FLOAT64, r, p->nccl_comm, p->comm_stream));
            if (ex.mine > 0) NCCL_TRY(n->Recv(ablk, (size_t)ex.mine * 2, B2_NCCL_FLOAT64, r, p->nccl_comm, p->comm_stream));
        }
    }
    NCCL_TRY(n->GroupEnd());
    b2i_count_launch();
    return 0;
}

// One evaluation of the nonlinear term + RK epilogue with the pipelined schedule of
// fluidsim_b200/slab.py::_run_stage: the all-to-all of (field f, chunk c) runs on the comm stream
// while the compute stream does the z pass of the next field / the y and x passes of the previous
// chunk.  vmax_dev != NULL: max |v| side output of the x passes (stage 0 of a CFL step).
static int slab_stage_native(b2_plan* p, int scheme, int stage, double dt, const double* dt_dev, const double* S_in,
                             double* S_, double* T_out, bool need_curl, double* vmax_dev, cudaStream_t s) {
    int e, nin, nout;
    if ((e = slab_ready(p))) return e;
    if (!p->comm_ready) return b2i_set_error("b2_slab_comm_init has not been called");
    slab_counts(p, &nin, &nout);
    const int nc = p->slab_nc;
    if (nin * nc * 2 + nout * nc * 2 > 128) return b2i_set_error("too many (field, chunk) pieces");
    const SlabExchange ex = slab_exchange(p);
    cudaStream_t cs = p->comm_stream;
    int order[8], no = 0;
    if (need_curl) {
        for (int f = 0; f < 3; ++f) order[no++] = f;
        for (int f = 6; f < nin; ++f) order[no++] = f;
        for (int f = 3; f < 6; ++f) order[no++] = f;
    } else {
        for (int f = 0; f < nin; ++f) order[no++] = f;
    }
    cudaEvent_t* ev = p->ev_comm;
    auto evA = [&](int f) { return ev[f]; };                              // zinv(f) done (compute)
    auto evI = [&](int f, int c) { return ev[8 + f * nc + c]; };          // a2a_inv(f, c) done (comm)
    auto evY = [&](int f, int c) { return ev[8 + (nin + f) * nc + c]; };  // yfwd(f, c) done (compute)
    auto evF = [&](int f, int c) { return ev[8 + (nin + nout + f) * nc + c]; };  // a2a_fwd done (comm)
    if (8 + (nin + 2 * nout) * nc > 128) return b2i_set_error("too many (field, chunk) pieces");
    bool curl_done = !need_curl;
    for (int k = 0; k < no; ++k) {
        const int f = order[k];
        if (!curl_done && f >= 3 && f < 6) {
            if ((e = b2_slab_curl(p, S_in, s))) return e;
            curl_done = true;
        }
        if ((e = b2_slab_zinv(p, S_in, f, f + 1, s))) return e;
        CUDA_TRY(cudaEventRecord(evA(f), s));
        CUDA_TRY(cudaStreamWaitEvent(cs, evA(f), 0));
        if ((e = slab_a2a(p, ex, f, 0, true))) return e;
        CUDA_TRY(cudaEventRecord(evI(f, 0), cs));
    }
    for (int c = 1; c < nc; ++c)
        for (int k = 0; k < no; ++k) {
            const int f = order[k];
            if ((e = slab_a2a(p, ex, f, c, true))) return e;
            CUDA_TRY(cudaEventRecord(evI(f, c), cs));
        }
    for (int c = 0; c < nc; ++c) {
        for (int k = 0; k < no; ++k) {
            const int f = order[k];
            CUDA_TRY(cudaStreamWaitEvent(s, evI(f, c), 0));
            if ((e = b2_slab_yinv(p, f, f + 1, c, s))) return e;
        }
        {
            const long long fs_unused = 0; (void)fs_unused;
            cplx* XW[8];
            for (int f = 0; f < nin; ++f) XW[f] = p->xa + f * p->xa_stride();
            const double scale = 1.0 / ((double)p->gy * p->n1 * p->n2);
            const int pitch = p->prune ? p->keepx : p->nk;
            const long long lines_c = (long long)p->gy * (p->nzl / nc);
            ProfScope ps(PC_X_FUSED, s);
            if ((e = b2i_xpass_fused(p, XW, lines_c, scale, pitch, pitch, lines_c * c, s, vmax_dev))) return e;
        }
        for (int f = 0; f < nout; ++f) {
            if ((e = b2_slab_yfwd(p, f, f + 1, c, s))) return e;
            CUDA_TRY(cudaEventRecord(evY(f, c), s));
            CUDA_TRY(cudaStreamWaitEvent(cs, evY(f, c), 0));
            if ((e = slab_a2a(p, ex, f, c, false))) return e;
            CUDA_TRY(cudaEventRecord(evF(f, c), cs));
        }
    }
    if (vmax_dev) {
        // CFL: global max |v_i| (allreduce MAX of three doubles on the comm stream, after the x passes)
        NcclApi* n = nccl_api();
        CUDA_TRY(cudaEventRecord(evA(7), s));
        CUDA_TRY(cudaStreamWaitEvent(cs, evA(7), 0));
        NCCL_TRY(n->AllReduce(vmax_dev, vmax_dev, 3, B2_NCCL_FLOAT64, B2_NCCL_MAX, p->nccl_comm, cs));
        CUDA_TRY(cudaEventRecord(evA(7), cs));
    }
    for (int f = 0; f < nout; ++f) {
        for (int c = 0; c < nc; ++c) CUDA_TRY(cudaStreamWaitEvent(s, evF(f, c), 0));
        if ((e = b2_slab_zfwd(p, f, f + 1, s))) return e;
    }
    if (vmax_dev) CUDA_TRY(cudaStreamWaitEvent(s, evA(7), 0));
    (void)dt_dev;
    return 0;  // the caller launches the epilogue (it may first need the CFL kernel)
}

static int slab_rk_native(b2_plan* p, int scheme, int stage, double dt, const double* dt_dev, const double* S_in,
                          double* S_, double* T_out, cudaStream_t s) {
    int mode;
    if (stage < 0) mode = M_TEND;
    else if (scheme == B2_SCHEME_RK4 && stage < 4) mode = M_RK4_0 + stage;
    else if (scheme == B2_SCHEME_RK2 && stage < 2) mode = M_RK2_0 + stage;
    else return b2i_set_error("bad scheme/stage %d/%d", scheme, stage);
    if (mode != M_TEND && (!p->acc || !p->stage)) return b2i_set_error("b2_set_buffers: acc/stage missing");
    RKArgs a = rk_args(p, (const cplx*)S_in, (cplx*)S_, dt);
    a.dt_ptr = dt_dev;
    a.Tout = (cplx*)T_out;
    return launch_rk_stage(p, mode, a, s);
}

/* N(S_in) on a slab plan with the library's own all-to-alls (stage < 0 semantics of b2_slab_rk) */
extern "C" int b2_slab_tendencies(b2_plan* p, const double* S_in, double* T_out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int e;
    if ((e = slab_stage_native(p, B2_SCHEME_RK4, -1, 0.0, nullptr, S_in, nullptr, T_out, true, nullptr, s))) return e;
    return slab_rk_native(p, B2_SCHEME_RK4, -1, 0.0, nullptr, S_in, nullptr, T_out, s);
}

/* one full RK2 / RK4 step on a slab plan: FFT passes, all-to-alls (NCCL, overlapped), epilogues */
extern "C" int b2_slab_time_step(b2_plan* p, int scheme, double dt, double* S_, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (scheme != B2_SCHEME_RK4 && scheme != B2_SCHEME_RK2)
        return b2i_set_error("Problem name time_scheme (scheme id %d)", scheme);
    const int nst = scheme == B2_SCHEME_RK4 ? 4 : 2;
    int e;
    for (int st = 0; st < nst; ++st) {
        const double* Sin = st == 0 ? S_ : (const double*)p->stage;
        if ((e = slab_stage_native(p, scheme, st, dt, nullptr, Sin, S_, nullptr, st == 0, nullptr, s))) return e;
        if ((e = slab_rk_native(p, scheme, st, dt, nullptr, Sin, S_, nullptr, s))) return e;
    }
    return 0;
}

/* the same with the CFL time increment decided on the device from the GLOBAL max |v_i|
 * (base/time_stepping/base.py:320-354 with its MPI allreduce); see b2_time_step_cfl */
extern "C" int b2_slab_time_step_cfl(b2_plan* p, int scheme, double cfl, double deltat_max, double* dt_dev,
                                     double* vmax_dev, double* S_, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (scheme != B2_SCHEME_RK4 && scheme != B2_SCHEME_RK2)
        return b2i_set_error("Problem name time_scheme (scheme id %d)", scheme);
    if (!dt_dev || !vmax_dev) return b2i_set_error("b2_slab_time_step_cfl: dt_dev / vmax_dev missing");
    const int nst = scheme == B2_SCHEME_RK4 ? 4 : 2;
    int e;
    for (int st = 0; st < nst; ++st) {
        const double* Sin = st == 0 ? S_ : (const double*)p->stage;
        if ((e = slab_stage_native(p, scheme, st, 0.0, dt_dev, Sin, S_, nullptr, st == 0, st == 0 ? vmax_dev : nullptr, s)))
            return e;
        if (st == 0) {
            // slab plans store (L0, L1, L2) = (Ly, Lz, Lx) and n1 = nz
            const double idx = p->n2 / p->L2, idy = p->gy / p->L0, idz = p->n1 / p->L1;
            cfl_dt_kernel<<<1, 1, 0, s>>>(vmax_dev, idx, idy, idz, cfl, deltat_max, dt_dev);
            B2_LAUNCH_CHECK("cfl_dt_kernel");
        }
        if ((e = slab_rk_native(p, scheme, st, 0.0, dt_dev, Sin, S_, nullptr, s))) return e;
    }
    return 0;
}
