// Internal declarations shared by the translation units of libb200spectral.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200spectral.h"
#include "fft_core.cuh"

struct b2_plan {
    int ndim;
    int n0, n1, n2;  // X shape (2-D plans are stored with n0 == 1)
    int nk;          // n2/2 + 1
    double L0, L1, L2;
    bool fast0, fast1, fast2;  // register-FFT path available along that axis
    cplx *tw0, *tw1, *tw2;     // exp(-2 pi i k / n) tables (device)
    double *k0, *k1, *kx;      // dimensional wavenumbers along K axes 0, 1, 2 (device)
    // physics (b2_set_physics)
    int solver;
    double nu2, nu4, nu8, num4, f, N, beta;
    int has_f;
    int projection; // params.projection: 0 None (project_perpk3d), 1 toroidal / vortical, 2 poloidal
    int no_vz_kz0;  // params.no_vz_kz0: vz (and b) are zeroed at kz = 0 after every projection
    const uint8_t* mask;
    // caller-owned buffers (b2_set_buffers)
    cplx *acc, *stage, *work;
    // dealias-pruned transforms (b2_set_pruning): kept index ranges per K axis, [0, lo) and [hi, n)
    int prune;
    int keep0_lo, keep0_hi, keep1_lo, keep1_hi, keepx;
    // slab decomposition (b2_plan_create_slab): the plan then describes the LOCAL K-layout array
    // (n0 = ny_loc, n1 = nz, n2 = nx; k0 = local ky, k1 = kz; tw1 = table of length nz)
    bool slab;
    int rank, nranks;
    int gy;        // global ny
    int nyl, nzl;  // ny / nranks, nz / nranks
    bool fasty;
    cplx* twy;     // table of length gy
    // chunked y/x/y overlap (api.cu): side streams and events, created on first use
    cudaStream_t sy1, sx, sy2;
    cudaEvent_t ev_begin, ev_end, ev_y[32], ev_x[32];
    bool streams_ready;
    cplx *xa, *xb; // exchange buffers (b2_slab_set_buffers), nwork fields each
    int gyk_lo, gyk_hi;  // global dealiased ky band (pruned slab y passes)
    int slab_nc;         // z chunks of the exchange layout (b2_slab_set_chunks)
    int ky_cyclic;       // ky rows dealt round-robin to the ranks (balances the pruned K side)
    // sparse forcing (b2_set_forcing_sparse): forcing_fft of the current time step, added to the raw
    // nonlinear term of every stage before projection / dealiasing
    long long force_n;
    const long long* force_idx;
    const cplx* force_val;
    int force_nvar;
    // native collectives (b2_slab_comm_init): NCCL communicator owned by the plan, its stream and the
    // events that order the per-(field, chunk) all-to-alls against the FFT passes
    void* nccl_comm;
    cudaStream_t comm_stream;
    cudaEvent_t ev_comm[128];
    bool comm_ready;
    long long xa_fs, xb_fs;  // elements between consecutive fields of xa / xb (0: fsize())
    // memory-lean buffers (b2_set_aliasing, ns3d): the raw transform outputs live in `stage`
    // (the epilogue rewrites them in place into the next stage input), `work` holds only omega (3 fields)
    int alias_tw;
    cplx* wfield(int f) const {  // work-field slot f: 0..2 = v / raw outputs, 3..5 = omega, 6 = b
        if (!alias_tw) return work + (long long)f * fsize();
        return f < 3 ? stage + (long long)f * fsize() : work + (long long)(f - 3) * fsize();
    }
    long long xa_stride() const { return xa_fs > 0 ? xa_fs : fsize(); }
    long long xb_stride() const { return xb_fs > 0 ? xb_fs : fsize(); }
    long long fsize() const { return (long long)n0 * n1 * nk; }  // complex elements per K field
    long long xsize() const { return (long long)n0 * n1 * n2; }
};

int b2i_set_error(const char* fmt, ...);
int b2i_check_launch(const char* what);
void b2i_count_launch();

#define B2_LAUNCH_CHECK(what)                     \
    do {                                          \
        b2i_count_launch();                       \
        int _e = b2i_check_launch(what);          \
        if (_e) return _e;                        \
    } while (0)

// --- strided passes (strided.cu).  axis: 0 (z, skipped when n0 == 1) or 1 (y).  dir: -1 fwd, +1 inv.
int b2i_strided_plain(b2_plan* p, int axis, int dir, const cplx* const* in, cplx* const* out, int nf,
                      double scale, cudaStream_t s, bool pruned = false, int outer0 = 0, int nouter = -1);
// first inverse pass of a stage with the k-space prologue fused on load:
//   ns3d / strat : in = nvar stage-input fields; out = W[0..2] = v, W[3..5] = curl v (+f), W[6] = b
//   ns2d         : in = rot; out = W[0]=ux, W[1]=uy, W[2]=d_x rot, W[3]=d_y rot
int b2i_first_inverse_pass(b2_plan* p, const cplx* const* in, cplx* const* out, cudaStream_t s);

// --- x passes (xpass.cu)
int b2i_xpass_c2r(b2_plan* p, const cplx* K, double* X, cudaStream_t s);
int b2i_xpass_r2c(b2_plan* p, const double* X, cplx* K, double scale, cudaStream_t s);
// fused c2r -> product -> r2c for p->solver; W fields as produced by b2i_first_inverse_pass
int b2i_xpass_fused(b2_plan* p, cplx* const* W, long long nlines, double scale, int nkeep, int pitch,
                    long long line0, cudaStream_t s, double* vmax = nullptr);
void b2i_xpass_share_sm(bool on);
// dealiased band [lo, hi) of the LOCAL ky rows of rank r (lo == hi: none), from the global band
void b2i_slab_local_band(const b2_plan* p, int r, int* lo, int* hi);

// --- slab (multi-GPU) passes (strided.cu).  Exchange layout of one field: [peer r][ky_loc][z_loc][kx]
// z pass between the local K layout (ny_loc, nz, nk) and the exchange layout:
//   dir = +1 (inverse): in = K-layout fields, out = exchange-layout send buffers
//   dir = -1 (forward): in = exchange-layout receive buffers, out = K-layout fields
int b2i_slab_zpass(b2_plan* p, int dir, const cplx* const* in, cplx* const* out, int nf, cudaStream_t s);
// y pass in place on receive buffers viewed as (ny, nz_loc, nk)
int b2i_slab_ypass(b2_plan* p, int dir, const cplx* const* in, cplx* const* out, int nf, int chunk,
                   cudaStream_t s);
