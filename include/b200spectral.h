/* b200spectral.h -- C ABI of libb200spectral.so (B200 / sm_100a pseudo-spectral hot path).
 *
 * Drop-in boundary for fluidsim's pseudo-spectral time step.  Every entry point takes raw DEVICE
 * pointers (float64 / complex128 = double[2], C-contiguous, the reference's array shapes) plus a
 * cudaStream_t passed as void*; returns 0 on success, <0 on error (message: b2_last_error()).
 * Nothing here allocates field-sized memory: state, accumulators and work buffers are caller-owned
 * (PyTorch tensors on the Python side).  No torch types cross this boundary.
 *
 * Array conventions (sequential layout, the one fluidfft's `fft3d.with_pyfftw` exposes,
 * /root/reference/fluidsim/operators/operators3d.py:120-133):
 *   3-D:  X (n0=nz, n1=ny, n2=nx) float64      K (n0, n1, n2/2+1) complex128, dimX_K = (0,1,2)
 *   2-D:  X (n0=ny, n1=nx)         float64      K (n0, n1/2+1)     complex128, not transposed
 * Forward transform is scaled by 1/(n0*n1*n2); inverse is unscaled (fluidfft convention,
 * /root/reference/fluidsim/base/state.py:385-392 relies on it).
 *
 * Each function names the reference interface it replaces (paths relative to
 * /root/reference/fluidsim/).
 */
#ifndef B200SPECTRAL_H
#define B200SPECTRAL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2_plan b2_plan;

/* solver ids for b2_set_physics */
#define B2_SOLVER_NS3D 0       /* solvers/ns3d/solver.py:180-253        */
#define B2_SOLVER_NS3D_STRAT 1 /* solvers/ns3d/strat/solver.py:138-216  */
#define B2_SOLVER_NS2D 2       /* solvers/ns2d/solver.py:111-194        */

/* time schemes for b2_time_step */
#define B2_SCHEME_RK2 2 /* base/time_stepping/pseudo_spect.py:469-517 */
#define B2_SCHEME_RK4 4 /* base/time_stepping/pseudo_spect.py:798-984 */

const char* b2_last_error(void);
int b2_version(void);

/* ---- plan: replaces fluidfft.create_fft_object(type_fft, n0, n1, n2) + the K-grid part of
 * OperatorsPseudoSpectral3D.__init__ (operators/operators3d.py:207-231).  ndim = 2 or 3; for
 * ndim == 2 pass n2 = 0 and L2 = 0 (n0 = ny, n1 = nx).  L0, L1, L2 are the box lengths along the
 * X axes (Lz, Ly, Lx) resp. (Ly, Lx). */
int b2_plan_create(b2_plan** out, int ndim, int n0, int n1, int n2, double L0, double L1, double L2);
int b2_plan_destroy(b2_plan* p);
/* shapeX[3], shapeK[3] (2-D: leading entry is 1): get_shapeX_seq / get_shapeK_seq */
int b2_plan_shapes(const b2_plan* p, int* shapeX, int* shapeK);
/* 1 if the fused register-FFT path handles these sizes (all axes powers of two >= 8) */
int b2_plan_is_fast(const b2_plan* p);

/* ---- transforms: replace fft_as_arg / ifft_as_arg / ifft_as_arg_destroy
 * (solvers/ns3d/solver.py:210-241, base/state.py:318-332).
 * b2_ifft_c2r: `work` is a K-sized complex scratch; pass NULL to transform in place in K
 * (ifft_as_arg_destroy semantics: K is clobbered). */
int b2_fft_r2c(b2_plan* p, const double* X, double* K, void* stream);
int b2_ifft_c2r(b2_plan* p, const double* K, double* X, double* work, void* stream);

/* ---- k-space / x-space operator kernels (fluidfft operator methods) */
/* rotfft_from_vecfft_outin: solvers/ns3d/solver.py:199 */
int b2_rotfft_from_vecfft(b2_plan* p, const double* vx, const double* vy, const double* vz,
                          double* rx, double* ry, double* rz, void* stream);
/* divfft_from_vecfft: used by div_vb_fft_from_vb, solvers/ns3d/strat/solver.py:206 */
int b2_divfft_from_vecfft(b2_plan* p, const double* vx, const double* vy, const double* vz,
                          double* div, void* stream);
/* project_perpk3d: solvers/ns3d/solver.py:255-259 (in place) */
int b2_project_perpk3d(b2_plan* p, double* vx, double* vy, double* vz, void* stream);
/* project_toroidal / project_poloidal: operators/operators3d.py:911-958 / 788-856 (in place) */
int b2_project_toroidal(b2_plan* p, double* vx, double* vy, double* vz, void* stream);
int b2_project_poloidal(b2_plan* p, double* vx, double* vy, double* vz, void* stream);
/* vector_product(a, b) -> written into b: solvers/ns3d/solver.py:226 */
int b2_vector_product(const double* ax, const double* ay, const double* az, double* bx, double* by,
                      double* bz, long long n, void* stream);
/* out = a * b (real fields): v_i * b products of div_vb_fft_from_vb */
int b2_mul_real(const double* a, const double* b, double* out, long long n, void* stream);
/* dealiasing_setofvar / dealiasing_variable: operators/operators3d.py:38-71, 336-342.
 * fields = nvar contiguous K arrays; mask = uint8 K-shaped (where_dealiased). */
int b2_dealias(b2_plan* p, double* fields, int nvar, const uint8_t* mask, void* stream);
/* 2-D: vecfft_from_rotfft, gradfft_from_fft (solvers/ns2d/solver.py:158,165), compute_Frot (:34-38) */
int b2_vecfft_from_rotfft2d(b2_plan* p, const double* rot, double* ux, double* uy, void* stream);
int b2_gradfft_from_fft2d(b2_plan* p, const double* f, double* px, double* py, void* stream);
int b2_rotfft_from_vecfft2d(b2_plan* p, const double* ux, const double* uy, double* rot, void* stream);
int b2_compute_frot(const double* ux, const double* uy, const double* px, const double* py, double beta,
                    double* out, long long n, void* stream);
/* tendencies_nonlin_ns2dstrat (solvers/ns2d/strat/solver.py:21-27; bouss = 0):
 *   f_rot = -ux px_rot - uy py_rot ; f_b = -ux px_b - uy py_b - N^2 uy
 * tendencies_nonlin_ns2dbouss (solvers/ns2d/bouss/solver.py:21-27; bouss = 1):
 *   f_rot = -ux px_rot - uy py_rot + px_b ; f_b = -ux px_b - uy py_b
 * real physical-space fields of n points; the outputs may alias px_rot / px_b */
int b2_tendencies_ns2d_buoyancy(const double* ux, const double* uy, const double* px_rot, const double* py_rot,
                                const double* px_b, const double* py_b, double N, int bouss, double* f_rot,
                                double* f_b, long long n, void* stream);
/* compute_fb_fft: solvers/ns3d/strat/solver.py:29-33 : fb = -div_vb - N^2 vz (in place in div_vb) */
int b2_compute_fb_fft(double* div_vb, double N, const double* vz, long long nk, void* stream);
/* a += b (complex arrays of nk elements): `fz_fft += b_fft`, strat/solver.py:198 */
int b2_add_inplace(double* a, const double* b, long long nk, void* stream);

/* ---- linear term + elementwise RK kernels (base/time_stepping/pseudo_spect.py) */
/* compute_freq_diss (base/solvers/pseudo_spect.py:134-191) + ExactLinearCoefs.compute (:122-141):
 * exact = exp(-dt*sigma), exact2 = exp(-dt/2*sigma), real K-shaped */
int b2_exact_coefs(b2_plan* p, double nu2, double nu4, double nu8, double num4, double dt, double* exact,
                   double* exact2, void* stream);
/* step_Euler :51-56   out = (S + dt*T) * diss   (nvar fields, diss broadcast over nvar) */
int b2_step_euler(b2_plan* p, const double* S, double dt, const double* T, const double* diss,
                  double* out, int nvar, void* stream);
/* step_like_RK2 :64-68   S = S*diss + dt*diss2*T */
int b2_step_like_rk2(b2_plan* p, double* S, double dt, const double* T, const double* diss,
                     const double* diss2, int nvar, void* stream);
/* rk4_step1 :938-941   acc += dt/3*diss2*T ; S12 = S*diss2 + dt/2*T */
int b2_rk4_step1(b2_plan* p, const double* S, double* acc, double* S12, const double* T,
                 const double* diss2, double dt, int nvar, void* stream);
/* rk4_step2 :968-971   acc += dt/3*diss2*T ; S1 = S*diss + dt*diss2*T */
int b2_rk4_step2(b2_plan* p, const double* S, double* acc, double* S1, const double* T,
                 const double* diss, const double* diss2, double dt, int nvar, void* stream);
/* rk4_step3 :984   S = acc + dt/6*T */
int b2_rk4_step3(b2_plan* p, double* S, const double* acc, const double* T, double dt, int nvar,
                 void* stream);

/* ---- reductions */
/* sum_wavenumbers(|f|^2) over nvar fields (r2c-aware weights), result written to *out_dev (device
 * double).  compute_energy_from_K = 0.5 * this. */
int b2_sum_wavenumbers_abs2(b2_plan* p, const double* fields, int nvar, double* out_dev, void* stream);
/* every reduction the reference's periodic outputs take from state_spect, in ONE pass: component
 * energies, dissipation rates epsK / epsK_hypo / epsK4 / epsK8, enstrophy
 * (solvers/ns3d/output/spatial_means.py:23-73), shell-binned 3-D spectra and 1-D spectra of every
 * variable (solvers/ns3d/output/spectra.py:15-60; fluidfft compute_3dspectrum / compute_1dspectra).
 * nks = oper.nk_spectra, deltak = oper.deltak.  out_dev (device doubles, b2_observables_size of
 * them): [0..3] E per variable, [4] epsK, [5] epsK_hypo, [6] epsK4, [7] epsK8, [8] enstrophy,
 * [9..15] reserved, then spec3d[nvar][nks], s_kx[nvar][nx/2+1], s_ky[nvar][ny/2+1],
 * s_kz[nvar][nz/2+1].  Uses the viscosities of b2_set_physics.  Slab plans: local contribution,
 * the caller all-reduces (SUM). */
long long b2_observables_size(const b2_plan* p, int nvar, int nks);
int b2_observables(b2_plan* p, const double* S, int nvar, int nks, double deltak, double* out_dev,
                   void* stream);
/* max |x| over n doubles -> *out_dev : _compute_time_increment_CLF_uxuyuz, base/time_stepping/base.py:320-339 */
int b2_max_abs(const double* x, long long n, double* out_dev, void* stream);
/* sum of all doubles (NaN check `np.isnan(np.sum(state_spect[0]))`, solvers/ns3d/time_stepping.py:19) */
int b2_sum(const double* x, long long n, double* out_dev, void* stream);

/* ---- fused path: replaces Simul.tendencies_nonlin + TimeSteppingPseudoSpectral._time_step_RK2/4 +
 * one_time_step_computation's project/dealias (solvers/ns3d/time_stepping.py:8-20). */
/* physics parameters: params.nu_2/4/8/m4, params.f (has_f=0 -> None), params.N, params.beta;
 * mask = where_dealiased (uint8, K-shaped, device; may be NULL = no dealiasing) */
int b2_set_physics(b2_plan* p, int solver, double nu2, double nu4, double nu8, double num4, int has_f,
                   double f, double N, double beta, const uint8_t* mask);
/* params.projection (solvers/ns3d/solver.py:139-174) used by the fused path for the tendencies and
 * the end-of-step state: 0 None (project_perpk3d), 1 "toroidal" / "vortical", 2 "poloidal" */
int b2_set_projection(b2_plan* p, int projection);
/* params.no_vz_kz0 (solvers/ns3d/solver.py:135-137, 260-263): vz (and b) are zeroed at kz = 0
 * after every projection of the tendencies and of the end-of-step state */
int b2_set_no_vz_kz0(b2_plan* p, int on);
/* dealias-pruned transforms: with on != 0 the fused path visits only the bounding box of the modes
 * kept by the mask (exact: everything outside is zero).  Requires the state to be dealiased, which
 * holds after every step (solvers/ns3d/time_stepping.py:16); off by default. */
int b2_set_pruning(b2_plan* p, int on);
/* *flag_dev (device int) = 1 if any of the nvar fields is non-zero at a mode the mask dealiases */
int b2_check_dealiased(b2_plan* p, const double* fields, int nvar, const uint8_t* mask, int* flag_dev,
                       void* stream);
/* out[0..4] = kept ranges [0,out[0]) U [out[1],n0), [0,out[2]) U [out[3],n1), kx < out[4] */
int b2_get_pruning_bounds(const b2_plan* p, int* out);
/* number of K-sized complex work fields the fused path needs for this solver (W), nvar of state */
int b2_work_fields(const b2_plan* p, int solver, int* nwork, int* nvar);
/* caller-owned buffers: acc, stage (each nvar K-fields), work (nwork K-fields) */
int b2_set_buffers(b2_plan* p, double* acc, double* stage, double* work);
/* memory-lean ns3d (call after b2_set_physics): the raw transform outputs are written into `stage`
 * and rewritten in place by the RK epilogue, so `work` needs only 3 K fields (the vorticity of the
 * next stage input): 12 instead of 15 K fields per GPU.  The reference keeps the same data in
 * state_spect + 2 tmp sets + fields_tmp[6] + fields_spect_tmp[3] (solvers/ns3d/state.py:46-52). */
int b2_set_aliasing(b2_plan* p, int on);
/* forcing: `tendencies_fft += self.forcing.get_forcing()` (solvers/ns3d/solver.py:243-244,
 * ns2d/solver.py:185-186).  The forcing makers of base/forcing/specific.py:137-345 act on at most
 * (2 nkmax_forcing)^3 low-wavenumber modes, so forcing_fft is passed as a sparse list: idx = nmodes
 * linear indices into a K field (device int64), val = nvar x nmodes complex128 (device), nvar leading
 * state variables.  The plan keeps the pointers (caller-owned, valid until replaced); every stage of
 * the following steps adds the list to the raw nonlinear term before projection and dealiasing.
 * nmodes = 0 switches forcing off. */
int b2_set_forcing_sparse(b2_plan* p, long long nmodes, const long long* idx, const double* val, int nvar);
/* T_out = N(S_in)  (projected + dealiased).  S_in is preserved; T_out may not alias S_in. */
int b2_tendencies(b2_plan* p, const double* S_in, double* T_out, void* stream);
/* one full step of `scheme` with time increment dt, in place on S, including the final
 * project_state_spect + dealiasing of one_time_step_computation. */
int b2_time_step(b2_plan* p, int scheme, double dt, double* S, void* stream);
/* ---- slab decomposition over the GPUs of one node.  One plan per rank; X
 * space split along z, K space along ky: local K layout (ny/P, nz, nx/2+1), dimX_K = (1,0,2), the
 * layout of fluidfft's fft3d.mpi_with_fftwmpi3d (operators/operators3d.py:384-391).  The fused
 * stage is cut at its two global transposes: phase A (z-inverse into send buffers), all-to-all,
 * phase B (y-inverse, fused x pass, y-forward), all-to-all, phase C (z-forward + RK epilogue).
 * The all-to-alls exchange, per field, nranks equal contiguous blocks (xa -> xb after phase A for
 * nvar+3 fields, xb -> xa after phase B for 3 (ns3d) / 6 (strat) fields) and are issued by the host
 * side (torch.distributed / NCCL; fluidsim_b200/slab.py). */
/* ky_cyclic != 0: ky rows are dealt round-robin (global row = yl * nranks + rank) instead of in
 * contiguous blocks, which balances the kept rows of the pruned transforms between the ranks */
int b2_plan_create_slab(b2_plan** out, int nz, int ny, int nx, double Lz, double Ly, double Lx, int rank,
                        int nranks, int ky_cyclic);
int b2_slab_set_buffers(b2_plan* p, double* xa, double* xb);
/* complex elements between consecutive fields of xa / xb (default: one full K field).  With the
 * pruned exchange (b2_slab_set_pruning) only b2_slab_buffer_need elements per field are touched. */
int b2_slab_set_buffer_strides(b2_plan* p, long long xa_field, long long xb_field);
int b2_slab_buffer_need(const b2_plan* p, long long* xa_field, long long* xb_field);
/* pruned exchange (see b2_set_pruning): kept ranges agreed between the ranks by the host side; the
 * all-to-alls then carry only the kept local ky rows x kx < keepx (uneven splits) */
int b2_slab_set_pruning(b2_plan* p, int on, int keepx, int kz_lo, int kz_hi, int yl_lo, int yl_hi,
                        int gy_lo, int gy_hi);
/* kept local ky rows of each rank (nranks ints) for the current pruning state */
int b2_slab_kept_rows(const b2_plan* p, int* nkl);
/* fine-grained pieces over work-field ranges [f0, f1) -- lets the host pipeline the per-field
 * all-to-alls with the FFT passes of the other fields (fluidsim_b200/slab.py) */
int b2_slab_curl(b2_plan* p, const double* S_in, void* stream);
int b2_slab_zinv(b2_plan* p, const double* S_in, int f0, int f1, void* stream);
/* the exchange layout is cut in nc z chunks ([chunk][peer][ky_loc][z in chunk][kx]) so that the
 * all-to-all of chunk c+1 overlaps the y / x passes of chunk c; chunk = -1 means all chunks */
int b2_slab_set_chunks(b2_plan* p, int nc);
int b2_slab_yinv(b2_plan* p, int f0, int f1, int chunk, void* stream);
int b2_slab_xpass(b2_plan* p, int chunk, void* stream);
int b2_slab_yfwd(b2_plan* p, int f0, int f1, int chunk, void* stream);
int b2_slab_zfwd(b2_plan* p, int f0, int f1, void* stream);
int b2_slab_rk(b2_plan* p, int scheme, int stage, double dt, const double* S_in, double* S, double* T_out,
               void* stream);
/* coarse phases = the same pieces, all fields at once */
int b2_slab_phase_a(b2_plan* p, const double* S_in, int need_curl, void* stream);
int b2_slab_phase_b(b2_plan* p, void* stream);
/* stage < 0: tendencies only (written to T_out); else stage of `scheme`, updating acc/stage/S */
int b2_slab_phase_c(b2_plan* p, int scheme, int stage, double dt, const double* S_in, double* S,
                    double* T_out, void* stream);

/* ---- native collectives: the communicator belongs to the plan, so a host
 * in any language drives the multi-GPU path with one call per time step.  NCCL (libnccl.so.2) is
 * bound at run time.  Rank 0 calls b2_nccl_unique_id, the host broadcasts the 128 bytes (MPI_Bcast,
 * torch.distributed, a file ...), every rank calls b2_slab_comm_init.  The all-to-alls are grouped
 * ncclSend / ncclRecv on a stream owned by the plan, overlapped with the FFT passes exactly like the
 * host-orchestrated schedule above; they replace the MPI transposes inside fluidfft's
 * fft3d.mpi_with_fftwmpi3d.fft_as_arg / ifft_as_arg. */
int b2_nccl_unique_id(void* out128);
int b2_slab_comm_init(b2_plan* p, const void* uid128);
int b2_slab_comm_destroy(b2_plan* p);
/* T_out = N(S_in) on a slab plan (Simul.tendencies_nonlin under MPI) */
int b2_slab_tendencies(b2_plan* p, const double* S_in, double* T_out, void* stream);
/* one full RK2 / RK4 step on a slab plan (TimeSteppingPseudoSpectral._time_step_RK2/4 +
 * one_time_step_computation's project / dealias under MPI) */
int b2_slab_time_step(b2_plan* p, int scheme, double dt, double* S, void* stream);
/* the same with the CFL time increment from the GLOBAL max |v_i| (ncclAllReduce MAX replaces the
 * mpi.comm.allreduce of base/time_stepping/base.py:25-26,320-354) */
int b2_slab_time_step_cfl(b2_plan* p, int scheme, double cfl, double deltat_max, double* dt_dev,
                          double* vmax_dev, double* S, void* stream);

/* one step with the CFL time increment (base/time_stepping/base.py:320-354) decided on the device:
 * the max |v_i| are a side output of the stage-0 fused x pass (no extra pass over memory, no
 * state_phys), dt_dev (device double, in/out) carries deltat with the 2 % hysteresis, vmax_dev = 3
 * device doubles zeroed once by the caller. */
int b2_time_step_cfl(b2_plan* p, int scheme, double cfl, double deltat_max, double* dt_dev,
                     double* vmax_dev, double* S, void* stream);

/* kernels launched by this library since load (for bench.py's gpu_launches) */
long long b2_launch_count(void);

/* ---- measurement hooks (bench.py roofline leg): per-kernel-class device time of the fused path,
 * CUDA events recorded on the launching stream.  Classes: 0 first inverse pass (+curl prologue),
 * 1 y inverse, 2 fused x pass, 3 y forward, 4 z forward, 5 RK / projection epilogue. */
int b2_profile_enable(int on);
int b2_profile_reset(void);
int b2_profile_get(double* ms, long long* count, int ncat);
/* development hook: one raw strided FFT pass (axis 0 = z, 1 = y; dir -1 fwd / +1 inv) over nf
 * contiguous K fields; used by the tuning micro-benchmarks only */
int b2_dev_strided_pass(b2_plan* p, int axis, int dir, const double* in, double* out, int nf, void* stream);

#ifdef __cplusplus
}
#endif
#endif
