// Register-resident Stockham FFT core for power-of-two line lengths (float64, sm_100a).
//
// One FFT line of N complex128 points is held by T = N/E threads, E points per thread
// (thread t owns positions t + m*T).  Each stage is a radix-r butterfly done in registers
// (r = E except possibly the last stage); between stages the line is re-distributed
// through a padded shared-memory plane pair (re / im planes of doubles, so every access
// is a conflict-free 64-bit access).  First-stage inputs and last-stage outputs sit at
// positions t + m*T, i.e. loads and stores go straight between HBM and registers.
//
// Replaces: the FFTW plans behind fluidfft's fft_as_arg / ifft_as_arg
// (/root/reference/fluidsim/solvers/ns3d/solver.py:210-241).
#pragma once
#include <cuda_runtime.h>

typedef double2 cplx;

#define B2_DEVINL __device__ __forceinline__

B2_DEVINL cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
B2_DEVINL cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
B2_DEVINL cplx cmul(cplx a, cplx b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
B2_DEVINL cplx cscale(cplx a, double s) { return make_double2(a.x * s, a.y * s); }
B2_DEVINL cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
// multiply by DIR*i  (DIR=-1: forward transform convention exp(-i..), DIR=+1: inverse)
template <int DIR>
B2_DEVINL cplx cmul_i(cplx a) {
    return DIR < 0 ? make_double2(a.y, -a.x) : make_double2(-a.y, a.x);
}
// multiply by exp(DIR * i * pi/4 * k) for small constant k (used inside radix-8/16)
template <int DIR, int K8>  // angle = DIR * 2*pi*K8/8
B2_DEVINL cplx cmul_w8(cplx a) {
    constexpr double h = 0.70710678118654752440;
    constexpr int k = ((K8 % 8) + 8) % 8;
    if (k == 0) return a;
    if (k == 2) return cmul_i<DIR>(a);
    if (k == 4) return make_double2(-a.x, -a.y);
    if (k == 6) return cmul_i<-DIR>(a);
    // odd k: (c + i s) with |c|=|s|=h
    const double c = (k == 1 || k == 7) ? h : -h;
    const double s = (DIR > 0 ? 1.0 : -1.0) * ((k == 1 || k == 3) ? h : -h);
    return make_double2(a.x * c - a.y * s, a.x * s + a.y * c);
}

// ---------------------------------------------------------------------------------------
// butterflies: in-place DFT_R on v[0..R-1], natural order in and out.
// out[n] = sum_m v[m] * exp(DIR*2*pi*i*m*n/R)
template <int R, int DIR>
struct Bfly;

template <int DIR>
struct Bfly<1, DIR> {
    static B2_DEVINL void run(cplx*) {}
};

template <int DIR>
struct Bfly<2, DIR> {
    static B2_DEVINL void run(cplx* v) {
        cplx a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};

template <int DIR>
struct Bfly<4, DIR> {
    static B2_DEVINL void run(cplx* v) {
        cplx t0 = cadd(v[0], v[2]);
        cplx t1 = csub(v[0], v[2]);
        cplx t2 = cadd(v[1], v[3]);
        cplx t3 = cmul_i<DIR>(csub(v[1], v[3]));
        v[0] = cadd(t0, t2);
        v[2] = csub(t0, t2);
        v[1] = cadd(t1, t3);
        v[3] = csub(t1, t3);
    }
};

template <int DIR>
struct Bfly<8, DIR> {
    static B2_DEVINL void run(cplx* v) {
        // decimation in frequency: 8 = 2 x 4
        cplx u[4], w[4];
        u[0] = cadd(v[0], v[4]);
        w[0] = csub(v[0], v[4]);
        u[1] = cadd(v[1], v[5]);
        w[1] = cmul_w8<DIR, 1>(csub(v[1], v[5]));
        u[2] = cadd(v[2], v[6]);
        w[2] = cmul_w8<DIR, 2>(csub(v[2], v[6]));
        u[3] = cadd(v[3], v[7]);
        w[3] = cmul_w8<DIR, 3>(csub(v[3], v[7]));
        Bfly<4, DIR>::run(u);
        Bfly<4, DIR>::run(w);
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            v[2 * n] = u[n];
            v[2 * n + 1] = w[n];
        }
    }
};

template <int DIR>
struct Bfly<16, DIR> {
    static B2_DEVINL void run(cplx* v) {
        // 16 = 4 x 4 : columns k (inputs k + 4l), then twiddle W16^(k n), then rows
        constexpr double c1 = 0.92387953251128675613;  // cos(pi/8)
        constexpr double s1 = 0.38268343236508977173;  // sin(pi/8)
        constexpr double sg = DIR > 0 ? 1.0 : -1.0;
        cplx b[4][4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            cplx col[4] = {v[k], v[k + 4], v[k + 8], v[k + 12]};
            Bfly<4, DIR>::run(col);
#pragma unroll
            for (int n = 0; n < 4; ++n) b[k][n] = col[n];
        }
        // twiddles W16^(k*n), k,n in 1..3  (k*n in {1,2,3,4,6,9})
        const cplx w1 = make_double2(c1, sg * s1);
        const cplx w3 = make_double2(s1, sg * c1);
        b[1][1] = cmul(b[1][1], w1);
        b[1][2] = cmul_w8<DIR, 1>(b[1][2]);
        b[1][3] = cmul(b[1][3], w3);
        b[2][1] = cmul_w8<DIR, 1>(b[2][1]);
        b[2][2] = cmul_i<DIR>(b[2][2]);
        b[2][3] = cmul_w8<DIR, 3>(b[2][3]);
        b[3][1] = cmul(b[3][1], w3);
        b[3][2] = cmul_w8<DIR, 3>(b[3][2]);
        // W16^9 = -W16^1
        b[3][3] = cmul(b[3][3], make_double2(-c1, -sg * s1));
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            cplx row[4] = {b[0][n], b[1][n], b[2][n], b[3][n]};
            Bfly<4, DIR>::run(row);
#pragma unroll
            for (int p = 0; p < 4; ++p) v[n + 4 * p] = row[p];
        }
    }
};

// ---------------------------------------------------------------------------------------
// shared-memory exchange plane: complex128 elements (128-bit LDS/STS), TK interleaved columns,
// element index = pad(w) * TK + c.  A quarter-warp (8 lanes x 16 B) must cover all 32 banks:
//   TK >= 8 : the 8 lanes are 8 columns of one row -> never a conflict, no padding
//   TK == 4 : two rows per quarter-warp -> pad so that rows r apart (r = radix) differ by an odd count
//   TK <= 2 : lanes run along the line -> pad so that strides 8 and 16 map to distinct 16-byte groups
// PM = 1 (TK = 4, E = 16 only): no padding, XOR swizzle of the row index instead -- the two rows a
// quarter-warp touches differ in parity for every access of the radix-16 stages, so the plane is
// exactly N x TK elements and can double as a TMA staging / store buffer (strided.cu).
template <int TK, int PM = 0>
B2_DEVINL int b2_pad(int w) {
    if (PM == 1) return w ^ ((w >> 4) & 1);
    if (TK >= 8) return w;
    if (TK == 4) return w + (w >> 4);
    return w + (w >> 3) + (w >> 6);
}
template <int N, int TK>
struct PlaneSize {  // complex elements per column
    static constexpr int value = TK >= 8 ? N : (TK == 4 ? N + (N >> 4) + 1 : N + (N >> 3) + (N >> 6) + 2);
};

struct SyncBlock {
    B2_DEVINL void operator()() const { __syncthreads(); }
};
struct SyncWarp {
    B2_DEVINL void operator()() const { __syncwarp(); }
};
struct SyncNone {
    B2_DEVINL void operator()() const {}
};
// named barrier over a group of NT threads (NT multiple of 32), id in 1..15
template <int NT>
struct SyncNamed {
    int id;
    B2_DEVINL void operator()() const { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(NT) : "memory"); }
};

// powers w[m] = w1^m, m = 1..r-1, by binary products (depth <= 4 multiplications)
template <int r>
B2_DEVINL void twiddle_powers(cplx w1, cplx* w) {
    w[1] = w1;
    if (r > 2) w[2] = cmul(w1, w1);
    if (r > 3) w[3] = cmul(w[2], w1);
    if (r > 4) {
        w[4] = cmul(w[2], w[2]);
        w[5] = cmul(w[4], w1);
        w[6] = cmul(w[4], w[2]);
        w[7] = cmul(w[4], w[3]);
    }
    if (r > 8) {
        w[8] = cmul(w[4], w[4]);
#pragma unroll
        for (int k = 1; k < 8; ++k) w[8 + k] = cmul(w[8], w[k]);
    }
}

// v[m] *= w1^m, m = 1..r-1.  Radix 16 forms the powers in groups of four (w^4, w^8, w^12 times
// w^1..w^3) so that only six twiddles are live at a time (register pressure of the paired x pass).
template <int r>
B2_DEVINL void apply_twiddles(cplx* v, cplx w1) {
    if constexpr (r <= 8) {
        cplx w[r > 1 ? r : 2];
        twiddle_powers<r>(w1, w);
#pragma unroll
        for (int m = 1; m < r; ++m) v[m] = cmul(v[m], w[m]);
    } else {
        static_assert(r == 16, "radix");
        const cplx w2 = cmul(w1, w1), w3 = cmul(w2, w1);
        v[1] = cmul(v[1], w1);
        v[2] = cmul(v[2], w2);
        v[3] = cmul(v[3], w3);
        const cplx w4 = cmul(w2, w2);
        v[4] = cmul(v[4], w4);
        v[5] = cmul(v[5], cmul(w4, w1));
        v[6] = cmul(v[6], cmul(w4, w2));
        v[7] = cmul(v[7], cmul(w4, w3));
        const cplx w8 = cmul(w4, w4);
        v[8] = cmul(v[8], w8);
        v[9] = cmul(v[9], cmul(w8, w1));
        v[10] = cmul(v[10], cmul(w8, w2));
        v[11] = cmul(v[11], cmul(w8, w3));
        const cplx w12 = cmul(w8, w4);
        v[12] = cmul(v[12], w12);
        v[13] = cmul(v[13], cmul(w12, w1));
        v[14] = cmul(v[14], cmul(w12, w2));
        v[15] = cmul(v[15], cmul(w12, w3));
    }
}

// Stockham stages.  TWS = stride in the twiddle table (table holds exp(-2 pi i k / (N*TWS))).
// Twiddles: one table load (w^1) per butterfly, higher powers by multiplication -- the table
// gathers of a per-element lookup saturate the L1/LSU pipe (profiles/README.md).
template <int N, int E, int DIR, int TK, int TWS, int Ns, class Sync, int PM = 0>
struct FftStages {
    static B2_DEVINL void run(cplx (&x)[E], cplx* __restrict__ plane, int t, int c,
                              const cplx* __restrict__ tw, Sync sync) {
        constexpr int T = N / E;
        constexpr int Rem = N / Ns;
        constexpr int r = Rem >= E ? E : Rem;
        constexpr int Q = E / r;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            cplx v[r];
#pragma unroll
            for (int m = 0; m < r; ++m) v[m] = x[q + m * Q];
            if (Ns > 1) {
                const int jm = (t + q * T) & (Ns - 1);
                cplx w1 = __ldg(tw + (size_t)jm * (TWS * (N / (Ns * r))));
                if (DIR > 0) w1.y = -w1.y;
                apply_twiddles<r>(v, w1);
            }
            Bfly<r, DIR>::run(v);
#pragma unroll
            for (int m = 0; m < r; ++m) x[q + m * Q] = v[m];
        }
        if constexpr (Ns * r < N) {
            sync();  // WAR on the plane (previous exchange / previous line)
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const int j = t + q * T;
                const int base = (j / Ns) * (Ns * r) + (j & (Ns - 1));
#pragma unroll
                for (int m = 0; m < r; ++m) plane[b2_pad<TK, PM>(base + m * Ns) * TK + c] = x[q + m * Q];
            }
            sync();
#pragma unroll
            for (int m = 0; m < E; ++m) x[m] = plane[b2_pad<TK, PM>(t + m * T) * TK + c];
            FftStages<N, E, DIR, TK, TWS, Ns * r, Sync, PM>::run(x, plane, t, c, tw, sync);
        }
    }
};

// x[m] holds in[t + m*T] on entry and out[t + m*T] on exit (unnormalised).
template <int N, int E, int DIR, int TK, int TWS, int PM = 0, class Sync>
B2_DEVINL void fft_line(cplx (&x)[E], cplx* plane, int t, int c, const cplx* __restrict__ tw, Sync sync) {
    FftStages<N, E, DIR, TK, TWS, 1, Sync, PM>::run(x, plane, t, c, tw, sync);
}
