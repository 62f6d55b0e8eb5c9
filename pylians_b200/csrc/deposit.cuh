// Per-particle index / weight arithmetic shared by the deposit kernels.
//
// Restates (does not copy) the arithmetic of the reference kernels so that every particle lands in
// the same cells: library/MAS_library/MAS_library.pyx NGP :290-291, CIC :152-157, TSC :392-399,
// PCS :485-492.  What must match exactly is the CELL selection (fp32 product pos*inv without FMA
// contraction, double-precision floor/trunc); the weight polynomials are evaluated in fp32 and
// agree with the reference's double-then-rounded values to ~1 ulp (budget: 1e-5 relative).
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
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const float diff = fabsf(__fsub_rn((float)(m + j + 1), dist));
            float c;
            if (diff < 0.5f) c = 0.75f - diff * diff;
            else if (diff < 1.5f) { const float t = 1.5f - diff; c = 0.5f * t * t; }
            else c = 0.0f;
            C[j] = c;
        }
        return m + 1;
    } else {
        const int m = __double2int_rd((double)dist - 2.0);  // <int>floor(dist-2.0), :486
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float diff = fabsf(__fsub_rn((float)(m + j + 1), dist));
            float c;
            if (diff < 1.0f) c = (4.0f - 6.0f * diff * diff + 3.0f * diff * diff * diff) * (1.0f / 6.0f);
            else if (diff < 2.0f) { const float t = 2.0f - diff; c = t * t * t * (1.0f / 6.0f); }
            else c = 0.0f;
            C[j] = c;
        }
        return m + 1;
    }
}

}  // namespace pylb
