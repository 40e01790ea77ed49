// Paired x pass, ns2d instantiations (see xpair.cuh).
#include "xpair_launch.cuh"

int b2i_xpair_ns2d(b2_plan* p, const PairOp& op, long long nlines, double scale, int nkeep, int pitch,
                   long long line0, cudaStream_t s) {
    return launch_pair<2>(p, op, nlines, scale, nkeep, pitch, line0, s);
}
