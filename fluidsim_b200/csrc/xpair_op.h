// Argument block and entry points of the paired x pass (xpair.cuh), one translation unit per solver.
#pragma once
#include "internal.h"

// KIND 0: ns3d   in = vx vy vz wx wy wz          out = (v x w)_xyz                   (W[0..2])
// KIND 1: strat  in = vx vy vz wx wy wz b        out = (v x w)_xyz, vx b, vy b, vz b (W[0..5])
// KIND 2: ns2d   in = ux uy d_x rot d_y rot      out = -u . grad rot - beta uy       (W[0])
struct PairOp {
    const cplx* in[7];
    cplx* out[6];
    double* vmax;  // optional max |u| side output (CFL) of the velocity components
    double beta;
};

int b2i_xpair_ns3d(b2_plan* p, const PairOp& op, long long nlines, double scale, int nkeep, int pitch,
                   long long line0, cudaStream_t s);
int b2i_xpair_strat(b2_plan* p, const PairOp& op, long long nlines, double scale, int nkeep, int pitch,
                    long long line0, cudaStream_t s);
int b2i_xpair_ns2d(b2_plan* p, const PairOp& op, long long nlines, double scale, int nkeep, int pitch,
                   long long line0, cudaStream_t s);
