// Per-particle index / weight arithmetic shared by the deposit kernels.
//
// Restates (does not copy) the arithmetic of the reference kernels so that every particle lands in
// the same cells: library/MAS_library/MAS_library.pyx NGP :290-291, CIC :152-157, TSC :392-399,
// PCS :485-492.  What must match exactly is the CELL selection (fp32 product pos*inv without FMA
// contraction, double-precision floor/trunc).  The TSC/PCS weight polynomials are evaluated in double
// and rounded once to float, as the reference's C does (its literals are doubles): an fp32 `*(1/6)`
// biases every PCS weight by +3e-8 and fails the reference's own 8-decimal mass-conservation test
// (Test/test_MAS.py:52-84).
#pragma once
#include "common.cuh"

namespace pylb {

template <int MAS>
struct Support {
    static constexpr int S = (MAS == PYLB_NGP) ? 1 : MAS + 1;  // 1, 2, 3, 4 cells per axis
};

// lowest cell touched along one axis (unwrapped, may be negative) and the S weights
template <int MAS>
__device__ __forceinline__ int axis_stencil(float p, float inv, float (&C)[Support<MAS>::S]) {
    const float dist = __fmul_rn(p, inv);  // fp32 product, never contracted
    if constexpr (MAS == PYLB_NGP) {
        C[0] = 1.0f;
        return __double2int_rz((double)dist + 0.5);  // <int>(dist + 0.5), :290
    } else if constexpr (MAS == PYLB_CIC) {
        const int id = __float2int_rz(dist);         // <int>dist, :153-155
        const float u = __fsub_rn(dist, (float)id);
        C[0] = __fsub_rn(1.0f, u);
        C[1] = u;
        return id;
    } else if constexpr (MAS == PYLB_TSC) {
        const int m = __double2int_rd((double)dist - 1.5);  // <int>floor(dist-1.5), :393
        // `diff` is a float; the literals are doubles, so the polynomial is evaluated in double and
        // rounded once to float (:396-399).  Keeps sum(weights) == 1 to ~1e-8 like the reference.
        if (fabsf(dist) < 4194304.0f) {
            // m + 1 <= dist - 0.5 < m + 2 exactly, fp32 subtraction is monotone and 0.5 / 1.5 are representable, so
            // the three distances fall into [0.5, 1.5], [0, 0.5], [0.5, 1.5] and the reference's branches are known
            // in advance (where two branches meet they give the same value): no compares, no divergence.
            const float f1 = (float)(m + 1);
            const float d0 = fabsf(__fsub_rn(f1, dist)), d1 = fabsf(__fsub_rn(__fadd_rn(f1, 1.0f), dist)),
                        d2 = fabsf(__fsub_rn(__fadd_rn(f1, 2.0f), dist));
            const double e0 = 1.5 - (double)d0, e2 = 1.5 - (double)d2;
            C[0] = (float)(0.5 * e0 * e0);
            C[1] = (float)(0.75 - (double)__fmul_rn(d1, d1));
            C[2] = (float)(0.5 * e2 * e2);
            return m + 1;
        }
#pragma unroll
        for (int j = 0; j < 3; j++) {      // positions far outside the box (or not finite): the reference's branches as written
            const float diff = fabsf(__fsub_rn((float)(m + j + 1), dist));
            const double dd = (double)diff;
            float c;
            if (diff < 0.5f) c = (float)(0.75 - (double)__fmul_rn(diff, diff));
            else if (diff < 1.5f) c = (float)(0.5 * (1.5 - dd) * (1.5 - dd));
            else c = 0.0f;
            C[j] = c;
        }
        return m + 1;
    } else {
        const int m = __double2int_rd((double)dist - 2.0);  // <int>floor(dist-2.0), :486
        // double evaluation, one rounding to float (:489-492).  x / 6.0 is formed as x * (1.0 / 6.0): the two differ by
        // at most one ulp of the double, i.e. in the rounded float for about one weight in 2^29.
        constexpr double SIXTH = 1.0 / 6.0;
        if (fabsf(dist) < 4194304.0f) {
            // m + 2 = floor(dist): the four distances fall into [1, 2], [0, 1), (0, 1], (1, 2] -- outer, inner, inner,
            // outer branch of the reference; where branches meet (1 and 2) they give the same value
            const float f1 = (float)(m + 1);
            const float d0 = fabsf(__fsub_rn(f1, dist)), d1 = fabsf(__fsub_rn(__fadd_rn(f1, 1.0f), dist)),
                        d2 = fabsf(__fsub_rn(__fadd_rn(f1, 2.0f), dist)), d3 = fabsf(__fsub_rn(__fadd_rn(f1, 3.0f), dist));
            const double u0 = 2.0 - (double)d0, u3 = 2.0 - (double)d3, a1 = (double)d1, a2 = (double)d2;
            C[0] = (float)(u0 * u0 * u0 * SIXTH);
            C[1] = (float)(fma(fma(3.0, a1, -6.0), a1 * a1, 4.0) * SIXTH);
            C[2] = (float)(fma(fma(3.0, a2, -6.0), a2 * a2, 4.0) * SIXTH);
            C[3] = (float)(u3 * u3 * u3 * SIXTH);
            return m + 1;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {      // positions far outside the box (or not finite): the reference's branches as written
            const float diff = fabsf(__fsub_rn((float)(m + j + 1), dist));
            const double dd = (double)diff;
            float c;
            if (diff < 1.0f) c = (float)((4.0 - 6.0 * dd * dd + 3.0 * dd * dd * dd) * SIXTH);
            else if (diff < 2.0f) c = (float)((2.0 - dd) * (2.0 - dd) * (2.0 - dd) * SIXTH);
            else c = 0.0f;
            C[j] = c;
        }
        return m + 1;
    }
}

}  // namespace pylb
