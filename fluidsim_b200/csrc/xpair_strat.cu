// Paired x pass, strat instantiations (see xpair.cuh).
#include "xpair_launch.cuh"

int b2i_xpair_strat(b2_plan* p, const PairOp& op, long long nlines, double scale, int nkeep, int pitch,
                   long long line0, cudaStream_t s) {
    return launch_pair<1>(p, op, nlines, scale, nkeep, pitch, line0, s);
}
