/* A host in plain C driving the hot path through the C ABI (include/b200spectral.h): ns3d, n^3
 * Taylor-Green vortex (doc/test_cases/Taylor_Green_vortices/run_simul.py:40-54 of the reference),
 * fused RK4 steps on one B200.  No Python, no torch: device memory from the CUDA runtime.
 *
 *   gcc -O2 -I include examples/host_c/step_ns3d.c -o step_ns3d \
 *       -L fluidsim_b200 -lb200spectral -L /usr/local/cuda/lib64 -lcudart -lm \
 *       -Wl,-rpath,$PWD/fluidsim_b200
 *   ./step_ns3d 128 100          # E(0) = 0.125, decaying slowly at nu_2 = 1/1600
 *
 * tests/test_abi.py compiles and links this file on the CPU box; it needs a GPU to run.
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "b200spectral.h"

#define CK(call)                                                                      \
    do {                                                                              \
        if ((call) != 0) {                                                            \
            fprintf(stderr, "%s failed: %s\n", #call, b2_last_error());               \
            return 1;                                                                 \
        }                                                                             \
    } while (0)
#define CU(call)                                                                      \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) {                                                      \
            fprintf(stderr, "%s failed: %s\n", #call, cudaGetErrorString(e_));        \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

static int wavenumber(int i, int n) { return i <= n / 2 ? i : i - n; }

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 64;
    const int nsteps = argc > 2 ? atoi(argv[2]) : 10;
    const double L = 2.0 * M_PI, dt = 1e-2, nu2 = 1.0 / 1600.0, coef_dealiasing = 2.0 / 3.0;
    const int nk = n / 2 + 1;
    const size_t nX = (size_t)n * n * n, nK = (size_t)n * n * nk; /* points per X / K field */

    b2_plan* plan = NULL;
    CK(b2_plan_create(&plan, 3, n, n, n, L, L, L));
    if (!b2_plan_is_fast(plan)) {
        fprintf(stderr, "the fused path needs power-of-two sizes >= 8\n");
        return 1;
    }
    int nwork = 0, nvar = 0;
    CK(b2_work_fields(plan, B2_SOLVER_NS3D, &nwork, &nvar));

    /* caller-owned device memory: state, RK accumulator, stage input, work fields, dealiasing mask */
    double *S, *acc, *stage, *work, *X, *scalar;
    uint8_t* mask;
    CU(cudaMalloc((void**)&S, nvar * nK * 16));
    CU(cudaMalloc((void**)&acc, nvar * nK * 16));
    CU(cudaMalloc((void**)&stage, nvar * nK * 16));
    CU(cudaMalloc((void**)&work, (size_t)nwork * nK * 16));
    CU(cudaMalloc((void**)&X, nX * 8));
    CU(cudaMalloc((void**)&scalar, 8));
    CU(cudaMalloc((void**)&mask, nK));

    /* where_dealiased, cubic truncation (the mask is an INPUT of the library: any shape can be passed) */
    uint8_t* hmask = (uint8_t*)malloc(nK);
    const double kcut = coef_dealiasing * (n / 2 + 1); /* deltak = 1 for L = 2 pi */
    for (int iz = 0; iz < n; ++iz)
        for (int iy = 0; iy < n; ++iy)
            for (int ix = 0; ix < nk; ++ix)
                hmask[((size_t)iz * n + iy) * nk + ix] =
                    abs(wavenumber(iz, n)) >= kcut || abs(wavenumber(iy, n)) >= kcut || ix >= kcut;
    CU(cudaMemcpy(mask, hmask, nK, cudaMemcpyHostToDevice));

    /* Taylor-Green initial condition in X space (nz, ny, nx), transformed field by field */
    double* hX = (double*)malloc(nX * 8);
    for (int c = 0; c < 3; ++c) {
        for (int iz = 0; iz < n; ++iz)
            for (int iy = 0; iy < n; ++iy)
                for (int ix = 0; ix < n; ++ix) {
                    const double x = L * ix / n, y = L * iy / n, z = L * iz / n;
                    const double v = c == 0 ? sin(x) * cos(y) * cos(z) : c == 1 ? -cos(x) * sin(y) * cos(z) : 0.0;
                    hX[((size_t)iz * n + iy) * n + ix] = v;
                }
        CU(cudaMemcpy(X, hX, nX * 8, cudaMemcpyHostToDevice));
        CK(b2_fft_r2c(plan, X, S + (size_t)c * nK * 2, NULL));
    }

    /* physics + buffers (what Simul.__init__ / TimeStepping.__init__ set up in the reference) */
    CK(b2_set_physics(plan, B2_SOLVER_NS3D, nu2, 0.0, 0.0, 0.0, /*has_f=*/0, 0.0, /*N=*/0.0, /*beta=*/0.0, mask));
    CK(b2_set_buffers(plan, acc, stage, work));

    double energy = 0.0;
    for (int it = 0; it <= nsteps; ++it) {
        if (it % 10 == 0 || it == nsteps) {
            CK(b2_sum_wavenumbers_abs2(plan, S, 3, scalar, NULL));
            CU(cudaMemcpy(&energy, scalar, 8, cudaMemcpyDeviceToHost));
            printf("it = %4d  t = %.3f  E = %.12f\n", it, it * dt, 0.5 * energy);
        }
        if (it < nsteps) {
            /* after the first step the state is dealiased: the pruned transforms may be used */
            CK(b2_set_pruning(plan, it > 0));
            CK(b2_time_step(plan, B2_SCHEME_RK4, dt, S, NULL));
        }
    }
    CU(cudaDeviceSynchronize());
    printf("kernels launched: %lld\n", b2_launch_count());
    CK(b2_plan_destroy(plan));
    cudaFree(S); cudaFree(acc); cudaFree(stage); cudaFree(work); cudaFree(X); cudaFree(scalar); cudaFree(mask);
    free(hmask); free(hX);
    return 0;
}
