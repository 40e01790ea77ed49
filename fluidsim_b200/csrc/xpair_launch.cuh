// Launch side of the paired x pass (xpair.cuh); included by one translation unit per solver kind so
// that the (large) kernel instantiations compile in parallel.
#pragma once
#include <stdlib.h>

#include <type_traits>

#include "internal.h"
#include "passes.cuh"
#include "xpair.cuh"

#define B2_XSIZES(X) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048)

// ------------------------------------------------------------------------------- paired x pass
// per-size configuration of the paired kernel (xpair.cuh): E points of the length-N complex FFT
// per thread, T = N/E threads per line group, G groups per CTA
template <int N> struct XPCfg;
template <> struct XPCfg<2048> { static constexpr int E = 16, G = 1; };  // T = 128
template <> struct XPCfg<1024> { static constexpr int E = 16, G = 1; };  // T = 64
template <> struct XPCfg<512>  { static constexpr int E = 16, G = 2; };  // T = 32
template <> struct XPCfg<256>  { static constexpr int E = 16, G = 4; };  // T = 16
template <> struct XPCfg<128>  { static constexpr int E = 8,  G = 4; };  // T = 16
template <> struct XPCfg<64>   { static constexpr int E = 8,  G = 8; };  // T = 8
template <> struct XPCfg<32>   { static constexpr int E = 4,  G = 8; };  // T = 8
template <> struct XPCfg<16>   { static constexpr int E = 4,  G = 16; }; // T = 4
template <> struct XPCfg<8>    { static constexpr int E = 2,  G = 16; }; // T = 4

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

template <int N, int KIND, bool VMAX, bool PARK0, int MAXREG, bool PF>
static int launch_pair_cfg(const PairOp& op, long long nlines, const cplx* tw, double scale, int nkeep, int pitch,
                           long long line0, cudaStream_t s) {
    constexpr int E = XPCfg<N>::E, G = XPCfg<N>::G, T = N / E;
    constexpr int PARK_D = XPTraits<KIND>::PARK_D + (PARK0 ? 2 : 0);
    const size_t smem = (size_t)G * ((PF ? 2 : 0) + PlaneSize<N, 1>::value * 2 + PARK_D * N + (PF ? 4 * nkeep : 0)) *
                        sizeof(double);
    if (smem > 227 * 1024) return b2i_set_error("paired x pass: tile does not fit shared memory");
    auto kern = xpass_pair_kernel<N, E, G, KIND, VMAX, PARK0, MAXREG, PF>;
    if (smem > 48 * 1024) {
        // per device (a process may drive several GPUs): cheap enough to repeat on every launch
        cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ce != cudaSuccess) return b2i_set_error("xpass_pair_kernel: %s", cudaGetErrorString(ce));
    }
    static const int ppg_env = env_int("B2_XPPG", 1);
    const long long npairs = (nlines + 1) / 2;
    int ppg = ppg_env < 1 ? 1 : ppg_env;
    while (ppg > 1 && npairs / ((long long)G * ppg) < 148 * 4) ppg /= 2;  // keep the grid wide
    const unsigned grid = (unsigned)((npairs + (long long)G * ppg - 1) / ((long long)G * ppg));
    kern<<<grid, G * T, smem, s>>>(op, nlines, tw, scale, nkeep, pitch, line0, ppg);
    B2_LAUNCH_CHECK("xpass_pair_kernel");
    return 0;
}

// N >= 512: the input lines are staged with cp.async.bulk one transform ahead (PF).  Development
// knobs: B2_XNOPF=1 direct global loads; B2_XVAR=1 first pair parked in shared memory, <= 200 registers.
template <int N, int KIND>
static int launch_pair_n(const PairOp& op, long long nlines, const cplx* tw, double scale, int nkeep, int pitch,
                         long long line0, cudaStream_t s) {
    static const int var = env_int("B2_XVAR", 0);
    static const int nopf = env_int("B2_XNOPF", 0);
    if constexpr (N >= 512) {
        if (op.vmax) return launch_pair_cfg<N, KIND, true, false, 255, true>(op, nlines, tw, scale, nkeep, pitch, line0, s);
        if (!nopf) {
            if constexpr (KIND != 2) {
                if (var == 1) return launch_pair_cfg<N, KIND, false, true, 200, true>(op, nlines, tw, scale, nkeep, pitch, line0, s);
            }
            return launch_pair_cfg<N, KIND, false, false, 255, true>(op, nlines, tw, scale, nkeep, pitch, line0, s);
        }
        return launch_pair_cfg<N, KIND, false, false, 255, false>(op, nlines, tw, scale, nkeep, pitch, line0, s);
    } else {
        if (op.vmax) return launch_pair_cfg<N, KIND, true, false, 255, false>(op, nlines, tw, scale, nkeep, pitch, line0, s);
        return launch_pair_cfg<N, KIND, false, false, 255, false>(op, nlines, tw, scale, nkeep, pitch, line0, s);
    }
}

template <int KIND>
static int launch_pair(b2_plan* p, const PairOp& op, long long nlines, double scale, int nkeep, int pitch,
                       long long line0, cudaStream_t s) {
    switch (p->n2) {
#define B2_CASE(n) case n: return launch_pair_n<n, KIND>(op, nlines, p->tw2, scale, nkeep, pitch, line0, s);
        B2_XSIZES(B2_CASE)
#undef B2_CASE
    }
    return b2i_set_error("fused x pass: nx=%d not supported (power of two in [8, 2048])", p->n2);
}

