// qt_npsum.cuh — numpy's add.reduce summation order for a virtual array f(i) (device).
//
// The tuner's per-candidate means are np.mean over the episodes of one candidate (controllers/tuning.py:912-917);
// numpy sums pairwise in blocks of at most 128 with eight accumulators (numpy/_core/src/umath/loops_utils.h.src,
// pairwise_sum).  Restating that order makes a candidate's score bit-identical to the reference's whenever its
// per-episode metrics are.  Shared by reduce_scores_kernel (qt_reduce.cu) and the rollout's fused score epilogue
// (qt_rollout.cu).
#pragma once
#include <stdint.h>

namespace qt {

// pairwise_sum for a virtual array f(i), i in [lo, lo+n)
template <typename F>
__device__ double np_pairwise(const F &f, int64_t lo, int64_t n) {
    if (n < 8) {
        double res = 0.0;
        for (int64_t i = 0; i < n; i++) res += f(lo + i);
        return res;
    } else if (n <= 128) {
        double r[8];
#pragma unroll
        for (int j = 0; j < 8; j++) r[j] = f(lo + j);
        int64_t i;
        for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int j = 0; j < 8; j++) r[j] += f(lo + i + j);
        }
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += f(lo + i);
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise(f, lo, n2) + np_pairwise(f, lo + n2, n - n2);
    }
}

template <typename F>
__device__ double np_sum(const F &f, int64_t n) {
    if (n <= 0) return 0.0;
    return np_pairwise(f, 0, n);
}

}  // namespace qt
