// Callers and consumers either side of the MA -> Pk path:
//   * the elementwise steps the snapshot drivers do between reading a block and MA / Pk
//     (library/readsnap.py:376, Pk_library/Pk_snapshot.py:88,194,248-254),
//   * smoothing_library (FT_filter, field_smoothing: library/smoothing_library/smoothing_library.pyx:19-114),
//   * bispectrum_library.Bk (library/Pk_library/bispectrum_library.pyx:32-196): shell selection fused with the MAS
//     deconvolution, and the real-space product sums.
// All of it is streaming, HBM-bound work: grid-stride loops, 16-byte accesses where the layout allows, double sums
// reduced by warp shuffles before one red.global per warp.
#include "common.cuh"

namespace pylb {

static unsigned blocks_for(int64_t n, int threads, int per_sm) {
    int64_t b = (n + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count() * per_sm;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

__device__ __forceinline__ void warp_sum_to(double *out, double acc) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc != 0.0) atomicAdd(out, acc);
}

// x *= mul, one fp32 multiply per element (numpy `data *= math.sqrt(time)`, readsnap.py:376; pyfftw's 1/N)
__global__ void __launch_bounds__(256) scale_kernel(float *x, int64_t n, float mul) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = __fmul_rn(x[i], mul);
}

// delta /= mean; delta -= 1.0 with a caller-supplied mean (Pk_snapshot.py:84-88,191-194): fp32 divide, fp32 subtract
__global__ void __launch_bounds__(256) overdensity_mean_kernel(float *g, int64_t n, float mean) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        g[i] = __fsub_rn(__fdiv_rn(g[i], mean), 1.0f);
}

// dst += a*src (delta_tot += Omega*delta, Pk_snapshot.py:250): separate multiply and add, like numpy's temporaries
__global__ void __launch_bounds__(256) axpy_kernel(float *dst, const float *__restrict__ src, float a, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = __fadd_rn(dst[i], __fmul_rn(a, __ldg(src + i)));
}

// ---- smoothing_library.FT_filter :19-83: the filter on the grid, and its sum ----------------------------------
// KIND 0 Top-Hat: 1 where d2 <= R2 (int compared with float, :47-50).  KIND 1 Gaussian: (float)exp(-d2/(2.0*R2))
// evaluated in double (:66-67).  Grid-stride over cells; the double sum of the written floats goes to *norm.
template <int KIND>
__global__ void __launch_bounds__(256) filter_fill_kernel(float *__restrict__ field, int dims, float R2, double *norm) {
    const int middle = dims / 2;
    const int64_t total = (int64_t)dims * dims * dims;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const double inv = 1.0 / (2.0 * (double)R2);
    double acc = 0.0;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += stride) {
        const int k = (int)(c % dims);
        const int64_t r = c / dims;
        const int j = (int)(r % dims), i = (int)(r / dims);
        const int i1 = i > middle ? i - dims : i, j1 = j > middle ? j - dims : j, k1 = k > middle ? k - dims : k;
        const int d2 = i1 * i1 + j1 * j1 + k1 * k1;
        float v;
        if (KIND == 0) v = ((float)d2 <= R2) ? 1.0f : 0.0f;
        else v = (float)exp(-(double)d2 * inv);
        field[c] = v;
        acc += (double)v;
    }
    warp_sum_to(norm, acc);
}

// field[i] = field[i]/normalization: float / double -> double quotient rounded to float (:76-79)
__global__ void __launch_bounds__(256) filter_norm_kernel(float *field, int64_t n, const double *norm) {
    const double s = *norm;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        field[i] = (float)((double)field[i] / s);
}

// field_k *= filter_k, complex64 (:104-108); both operands are read once, one 8-byte access each
__global__ void __launch_bounds__(256) cmul_kernel(float2 *a, const float2 *__restrict__ b, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float2 x = a[i], y = __ldg(b + i);
        a[i] = make_float2(x.x * y.x - x.y * y.y, x.x * y.y + x.y * y.x);
    }
}

// ---- the large part of Pk/XPk's bookkeeping (Pk_library.pyx:397-406): the 2-D table -----------------------------------
// Pk2D[i] = sum[i] * (fact / Nmodes2D[i]) for every field and pair, and all mode counts rewritten in place as doubles
// (the reference's Nmodes arrays are float64).  Same IEEE double operations, in the same order, as the numpy lines
// they replace, so the values are bit-identical; 0.4 M bins at 1024^3 cost the host 1-2 ms per call, the GPU ~5 us.
__global__ void __launch_bounds__(256)
pk_finish_tables_kernel(double *sums, unsigned long long *counts, int64_t o_p2d, int64_t o_x2d, int64_t o_n2d, int64_t B2,
                        int F, int X, int64_t n_counts, double fact) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_counts; i += stride) {
        const double n = (double)counts[i];
        const int64_t b = i - o_n2d;
        if (b >= 0 && b < B2) {
            const double inv = fact / n;               // inf for an empty bin: the host raises ZeroDivisionError first
            for (int f = 0; f < F; f++) sums[o_p2d + b * F + f] *= inv;
            for (int x = 0; x < X; x++) sums[o_x2d + b * X + x] *= inv;
        }
        reinterpret_cast<double *>(counts)[i] = n;
    }
}

// ---- void_library.gaussian_smoothing (void_library/void_library.pyx:45-80): despite its name a top-hat of radius R,
// applied in k-space: delta_k *= 3 (sin kR - kR cos kR)/(kR)^3, kR = (float)(prefact*|k|); the DC mode is skipped.
// The trigonometry is double like the reference's libm calls; kR*kR*kR is a float product (:74).
__global__ void __launch_bounds__(256) tophat_k_kernel(float2 *dk, int dims, float prefact) {
    const int middle = dims / 2, nz = middle + 1;
    const int64_t total = (int64_t)dims * dims * nz;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        if (idx == 0) continue;
        const int kz = (int)(idx % nz);
        const int64_t r = idx / nz;
        const int kx = wavenumber((int)(r / dims), dims, middle), ky = wavenumber((int)(r % dims), dims, middle);
        const float kR = (float)((double)prefact * sqrt((double)(kx * kx + ky * ky + kz * kz)));
        float fact = 1.0f;
        if (fabsf(kR) >= 1e-5f) {
            const double x = (double)kR;
            const float kR3 = __fmul_rn(__fmul_rn(kR, kR), kR);
            fact = (float)(3.0 * (sin(x) - cos(x) * x) / (double)kR3);
        }
        const float2 z = dk[idx];
        dk[idx] = make_float2(__fmul_rn(z.x, fact), __fmul_rn(z.y, fact));
    }
}

// ---- bispectrum_library.Bk :88-130 ------------------------------------------------------------------------
// Every stored mode (no skip rule here, it is commented out in the reference :103-108) is multiplied by its fp32 MAS
// factor; the modes with k_min <= |k| < k_max (|k| = sqrt of the integer norm in double, :120-123) go to out_d and
// switch on out_i, all others are zero.  One pass over delta_k per shell instead of the reference's ID lists.
__global__ void __launch_bounds__(256)
bk_shell_kernel(const float2 *__restrict__ dk, float2 *__restrict__ out_d, float2 *__restrict__ out_i, int dims,
                int mas_index, double kmin, double kmax) {
    const int middle = dims / 2, nz = middle + 1;
    const int64_t total = (int64_t)dims * dims * nz;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const double prefact = M_PI / (double)dims;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int kz = (int)(idx % nz);
        const int64_t r = idx / nz;
        const int iy = (int)(r % dims), ix = (int)(r / dims);
        const int kx = wavenumber(ix, dims, middle), ky = wavenumber(iy, dims, middle);
        const double k = sqrt((double)(kx * kx + ky * ky + kz * kz));
        float2 d = make_float2(0.0f, 0.0f), one = make_float2(0.0f, 0.0f);
        if (k >= kmin && k < kmax) {
            double m = 1.0;
            if (mas_index > 0) {
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    const int ka = a == 0 ? kx : (a == 1 ? ky : kz);
                    if (ka != 0) {
                        const double x = prefact * (double)ka;
                        const double q = x / sin(x);
                        double v = q;
                        for (int p = 1; p < mas_index; p++) v *= q;
                        m *= v;
                    }
                }
            }
            const float mf = (float)m;
            const float2 z = __ldg(dk + idx);
            d = make_float2(__fmul_rn(z.x, mf), __fmul_rn(z.y, mf));
            one.x = 1.0f;
        }
        out_d[idx] = d;
        out_i[idx] = one;
    }
}

// out[0] += sum a*b[*c] with the product formed in fp32 left to right and summed in double (:141-146, :187-193)
__global__ void __launch_bounds__(256)
prod_sum_kernel(const float *__restrict__ a, const float *__restrict__ b, const float *__restrict__ c, int64_t n,
                double *out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float p = __fmul_rn(__ldg(a + i), __ldg(b + i));
        if (c != nullptr) p = __fmul_rn(p, __ldg(c + i));
        acc += (double)p;
    }
    warp_sum_to(out, acc);
}

}  // namespace pylb

using namespace pylb;

extern "C" int pylb_scale_f32(float *x, int64_t n, float mul, void *stream) {
    PYLB_REQUIRE(x || n == 0, "pylb_scale_f32: NULL pointer");
    if (n == 0) return 0;
    scale_kernel<<<blocks_for(n, 256, 16), 256, 0, (cudaStream_t)stream>>>(x, n, mul);
    PYLB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pylb_overdensity_mean(float *grid, int64_t n, float mean, void *stream) {
    PYLB_REQUIRE(grid && n > 0, "pylb_overdensity_mean: bad arguments");
    overdensity_mean_kernel<<<blocks_for(n, 256, 16), 256, 0, (cudaStream_t)stream>>>(grid, n, mean);
    PYLB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pylb_axpy_f32(float *dst, const float *src, float a, int64_t n, void *stream) {
    PYLB_REQUIRE(dst && src && n > 0, "pylb_axpy_f32: bad arguments");
    axpy_kernel<<<blocks_for(n, 256, 16), 256, 0, (cudaStream_t)stream>>>(dst, src, a, n);
    PYLB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pylb_filter_real(float *field, int dims, float R2, int kind, double *scratch, void *stream) {
    PYLB_REQUIRE(field && scratch && dims >= 2 && dims <= 2048 && (kind == 0 || kind == 1), "pylb_filter_real: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = (int64_t)dims * dims * dims;
    PYLB_CHECK(cudaMemsetAsync(scratch, 0, sizeof(double), st));
    if (kind == 0) filter_fill_kernel<0><<<blocks_for(n, 256, 16), 256, 0, st>>>(field, dims, R2, scratch);
    else filter_fill_kernel<1><<<blocks_for(n, 256, 16), 256, 0, st>>>(field, dims, R2, scratch);
    PYLB_LAUNCH_CHECK();
    filter_norm_kernel<<<blocks_for(n, 256, 16), 256, 0, st>>>(field, n, scratch);
    PYLB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pylb_cmul_c64(void *a, const void *b, int64_t n, void *stream) {
    PYLB_REQUIRE(a && b && n > 0, "pylb_cmul_c64: bad arguments");
    cmul_kernel<<<blocks_for(n, 256, 16), 256, 0, (cudaStream_t)stream>>>((float2 *)a, (const float2 *)b, n);
    PYLB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pylb_pk_finish_tables(double *sums, uint64_t *counts, int dims, int F, double fact, void *stream) {
    PYLB_REQUIRE(sums && counts, "pylb_pk_finish_tables: NULL pointer");
    pylb_pk_layout L;
    if (pylb_pk_get_layout(dims, F, &L)) return 1;
    pk_finish_tables_kernel<<<blocks_for(L.n_counts, 256, 4), 256, 0, (cudaStream_t)stream>>>(
        sums, reinterpret_cast<unsigned long long *>(counts), L.o_p2d, L.o_x2d, L.o_n2d, L.B2, L.F, L.X, L.n_counts, fact);
    PYLB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pylb_tophat_k(void *dk, int dims, float prefact, void *stream) {
    PYLB_REQUIRE(dk && dims >= 2 && dims <= 16384, "pylb_tophat_k: bad arguments");
    const int64_t n = (int64_t)dims * dims * (dims / 2 + 1);
    tophat_k_kernel<<<blocks_for(n, 256, 16), 256, 0, (cudaStream_t)stream>>>((float2 *)dk, dims, prefact);
    PYLB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pylb_bk_shell(const void *dk, void *out_d, void *out_i, int dims, int mas_index, double kmin, double kmax,
                             void *stream) {
    PYLB_REQUIRE(dk && out_d && out_i && dims >= 2 && dims <= 16384, "pylb_bk_shell: bad arguments");
    PYLB_REQUIRE(mas_index >= 0 && mas_index <= 4, "pylb_bk_shell: MAS index %d out of range", mas_index);
    const int64_t n = (int64_t)dims * dims * (dims / 2 + 1);
    bk_shell_kernel<<<blocks_for(n, 256, 16), 256, 0, (cudaStream_t)stream>>>((const float2 *)dk, (float2 *)out_d,
                                                                             (float2 *)out_i, dims, mas_index, kmin, kmax);
    PYLB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pylb_prod_sum(const float *a, const float *b, const float *c, int64_t n, double *out, void *stream) {
    PYLB_REQUIRE(a && b && out && n > 0, "pylb_prod_sum: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    PYLB_CHECK(cudaMemsetAsync(out, 0, sizeof(double), st));
    prod_sum_kernel<<<blocks_for(n, 256, 8), 256, 0, st>>>(a, b, c, n, out);
    PYLB_LAUNCH_CHECK();
    return 0;
}
