// Particle mass assignment (NGP / CIC / TSC / PCS, optional weights, 2-D and 3-D, fp32 or fp64 grid).
//
// Two algorithms:
//   DIRECT : one thread per particle, red.global.add straight into the grid.  Right when the grid
//            lives in L2 (small dims) or the particle order is already spatially coherent.
//   TILED  : (deposit_tiled.cu) particles are binned by cell tile, each tile is accumulated in
//            shared memory and flushed with vectorised red.global.add.v4.f32.
#include "deposit.cuh"

namespace pylb {

template <int MAS, bool HASW, typename GT, int NDIM>
__global__ void __launch_bounds__(256)
deposit_direct_kernel(const float *__restrict__ pos, int64_t np, int64_t ps0, int64_t ps1,
                      GT *__restrict__ grid, int dims, float inv, const float *__restrict__ W,
                      float zrep, int x0, int xext, int64_t wst) {
    constexpr int S = Support<MAS>::S;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += stride) {
        float C[3][S];
        int base[3];
        const float *p = pos + i * ps0;
#pragma unroll
        for (int a = 0; a < NDIM; a++) base[a] = axis_stencil<MAS>(__ldg(p + a * ps1), inv, C[a]);
        float w = HASW ? __ldg(W + i * wst) : 1.0f;
        if constexpr (NDIM == 2) w = (zrep == 1.0f) ? w : w * zrep;
        int ix[S], iy[S];
#pragma unroll
        for (int j = 0; j < S; j++) {
            ix[j] = wrap(base[0] + j - x0, dims);   // x window: local plane index; planes >= xext are not held here
            iy[j] = wrap(base[1] + j, dims);
        }
        if constexpr (NDIM == 3) {
            int iz[S];
#pragma unroll
            for (int j = 0; j < S; j++) iz[j] = wrap(base[2] + j, dims);
#pragma unroll
            for (int l = 0; l < S; l++)
#pragma unroll
                for (int m = 0; m < S; m++) {
                    if (ix[l] >= xext) continue;
                    const int64_t row = ((int64_t)ix[l] * dims + iy[m]) * dims;
                    const float cxy = C[0][l] * C[1][m];  // left-to-right product, :159-166
#pragma unroll
                    for (int n = 0; n < S; n++) {
                        float v = cxy * C[2][n];
                        if (HASW) v *= w;
                        red_add(grid + row + iz[n], v);
                    }
                }
        } else {
#pragma unroll
            for (int l = 0; l < S; l++)
#pragma unroll
                for (int m = 0; m < S; m++) {
                    float v = C[0][l] * C[1][m];
                    if (HASW || zrep != 1.0f) v *= w;
                    if (ix[l] < xext) red_add(grid + (int64_t)ix[l] * dims + iy[m], v);
                }
        }
    }
}

template <int MAS, bool HASW, typename GT, int NDIM>
static int launch_direct(const float *pos, int64_t np, int64_t ps0, int64_t ps1, void *grid, int dims,
                         float inv, const float *w, float zrep, int x0, int xext, int64_t wst, cudaStream_t st) {
    if (np == 0) return 0;
    const int threads = 256;
    int64_t blocks = (np + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count() * 32;
    if (blocks > cap) blocks = cap;
    timing_begin(PYLB_T_DIRECT, st);
    deposit_direct_kernel<MAS, HASW, GT, NDIM><<<(unsigned)blocks, threads, 0, st>>>(
        pos, np, ps0, ps1, (GT *)grid, dims, inv, w, zrep, x0, xext, wst);
    timing_end(PYLB_T_DIRECT, st);
    PYLB_LAUNCH_CHECK();
    return 0;
}

template <int MAS, bool HASW, typename GT>
static int direct_ndim(int ndim, const float *pos, int64_t np, int64_t ps0, int64_t ps1, void *grid,
                       int dims, float inv, const float *w, float zrep, int x0, int xext, int64_t wst, cudaStream_t st) {
    if (ndim == 3) return launch_direct<MAS, HASW, GT, 3>(pos, np, ps0, ps1, grid, dims, inv, w, zrep, x0, xext, wst, st);
    return launch_direct<MAS, HASW, GT, 2>(pos, np, ps0, ps1, grid, dims, inv, w, zrep, x0, xext, wst, st);
}

template <int MAS>
static int direct_mas(bool hasw, bool f64, int ndim, const float *pos, int64_t np, int64_t ps0,
                      int64_t ps1, void *grid, int dims, float inv, const float *w, float zrep,
                      int x0, int xext, int64_t wst, cudaStream_t st) {
    if (hasw) {
        if (f64) return direct_ndim<MAS, true, double>(ndim, pos, np, ps0, ps1, grid, dims, inv, w, zrep, x0, xext, wst, st);
        return direct_ndim<MAS, true, float>(ndim, pos, np, ps0, ps1, grid, dims, inv, w, zrep, x0, xext, wst, st);
    }
    if (f64) return direct_ndim<MAS, false, double>(ndim, pos, np, ps0, ps1, grid, dims, inv, w, zrep, x0, xext, wst, st);
    return direct_ndim<MAS, false, float>(ndim, pos, np, ps0, ps1, grid, dims, inv, w, zrep, x0, xext, wst, st);
}

int ma_direct(const float *pos, int64_t np, int ndim, int64_t ps0, int64_t ps1, void *grid, int f64,
              int dims, float inv, int mas, const float *w, float zrep, int x0, int xext, int64_t wst, cudaStream_t st) {
    switch (mas) {
        case PYLB_NGP: return direct_mas<PYLB_NGP>(w != nullptr, f64, ndim, pos, np, ps0, ps1, grid, dims, inv, w, zrep, x0, xext, wst, st);
        case PYLB_CIC: return direct_mas<PYLB_CIC>(w != nullptr, f64, ndim, pos, np, ps0, ps1, grid, dims, inv, w, zrep, x0, xext, wst, st);
        case PYLB_TSC: return direct_mas<PYLB_TSC>(w != nullptr, f64, ndim, pos, np, ps0, ps1, grid, dims, inv, w, zrep, x0, xext, wst, st);
        case PYLB_PCS: return direct_mas<PYLB_PCS>(w != nullptr, f64, ndim, pos, np, ps0, ps1, grid, dims, inv, w, zrep, x0, xext, wst, st);
    }
    set_error("pylb_ma: unknown mass-assignment scheme %d", mas);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// small elementwise helpers on the grid
// ------------------------------------------------------------------------------------------------
__global__ void divide_kernel(float *g, int64_t n, float d) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) g[i] = __fdiv_rn(g[i], d);
}

// sum in double: per-thread partial -> warp shuffle -> one red.global.add.f64 per warp
__global__ void __launch_bounds__(256) sum_kernel(const float *__restrict__ g, int64_t n, double *out) {
    double acc = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = n / 4;
    const float4 *g4 = reinterpret_cast<const float4 *>(g);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = __ldg(g4 + i);
        acc += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
    }
    for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc += (double)g[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// delta = g/mean - 1 as the callers write it (`delta /= np.mean(delta, dtype=np.float64); delta -= 1.0`,
// Pk_snapshot.py:88): the quotient is formed in float64 (x * (1/mean), <= 1 ulp of double away from the
// true quotient), rounded to float32, then 1.0f is subtracted in float32.
__global__ void __launch_bounds__(256) overdensity_kernel(float *g, int64_t n, const double *sum, int64_t n_total) {
    const double rinv = (double)n_total / sum[0];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = n / 4;
    float4 *g4 = reinterpret_cast<float4 *>(g);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v = g4[i];
        v.x = (float)((double)v.x * rinv) - 1.0f; v.y = (float)((double)v.y * rinv) - 1.0f;
        v.z = (float)((double)v.z * rinv) - 1.0f; v.w = (float)((double)v.w * rinv) - 1.0f;
        g4[i] = v;
    }
    for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        g[i] = (float)((double)g[i] * rinv) - 1.0f;
}

__global__ void __launch_bounds__(256)
rsd_kernel(float *pos, const float *__restrict__ vel, int64_t np, float box, float factor, int axis) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += stride) {
        // pos + vel*factor: separate fp32 multiply and add like the reference's C (no contraction)
        float p = __fadd_rn(pos[3 * i + axis], __fmul_rn(__ldg(vel + 3 * i + axis), factor));
        if (p > box || p < 0.0f) p = fmodf(__fadd_rn(p, box), box);  // redshift_space_library.pyx:42-43
        pos[3 * i + axis] = p;
    }
}

// Tiled 2-D transposes (32x32 tile through padded shared memory, coalesced on both sides).
//   MODE 1 (axis 1): out[i][k][j] = in[i][j][k]   one (j,k) plane per blockIdx.z = i
//   MODE 0 (axis 0): out[k][j][i] = in[i][j][k]   one (i,k) plane per blockIdx.z = j
template <int MODE>
__global__ void __launch_bounds__(256) swap_axes_kernel(const float *__restrict__ in, float *__restrict__ out, int N, int64_t pitch) {
    __shared__ float tile[32][33];
    const int p = blockIdx.z;
    const int a0 = blockIdx.y * 32, c0 = blockIdx.x * 32;   // tile origin: a = slow input index (j or i), c = k
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5; // 32 x 8
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int a = a0 + r, c = c0 + tx;
        if (a < N && c < N) {
            const int64_t src = (MODE == 1) ? ((int64_t)p * N + a) * N + c : ((int64_t)a * N + p) * N + c;
            tile[r][tx] = in[src];
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, a = a0 + tx;                  // now a is the fast output index
        if (a < N && c < N) {
            const int64_t dst = (MODE == 1) ? ((int64_t)p * N + c) * pitch + a : ((int64_t)c * N + p) * pitch + a;
            out[dst] = tile[tx][r];
        }
    }
}

__global__ void __launch_bounds__(256) copy_pitched_kernel(const float *__restrict__ in, float *__restrict__ out, int N, int64_t pitch, int64_t nrows) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nrows * N; i += stride) {
        const int64_t r = i / N;
        out[r * pitch + (i - r * N)] = in[i];
    }
}

static unsigned grid_for(int64_t n, int threads, int per_sm) {
    int64_t b = (n + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count() * per_sm;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

}  // namespace pylb

using namespace pylb;

extern "C" int pylb_divide(float *grid, int64_t n, float divisor, void *stream) {
    if (n == 0) return 0;
    divide_kernel<<<grid_for(n, 256, 16), 256, 0, (cudaStream_t)stream>>>(grid, n, divisor);
    PYLB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pylb_h2d_padded(const float *host, float *dev, int dims, void *stream) {
    PYLB_REQUIRE(host && dev && dims >= 2, "pylb_h2d_padded: bad arguments");
    const size_t nz2 = 2 * ((size_t)dims / 2 + 1);
    PYLB_CHECK(cudaMemcpy2DAsync(dev, nz2 * sizeof(float), host, (size_t)dims * sizeof(float), (size_t)dims * sizeof(float),
                                 (size_t)dims * dims, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return 0;
}

extern "C" int pylb_h2d_pitched(const float *host, float *dev, int dims, int64_t dev_pitch, void *stream) {
    PYLB_REQUIRE(host && dev && dims >= 2 && dev_pitch >= dims, "pylb_h2d_pitched: bad arguments");
    PYLB_CHECK(cudaMemcpy2DAsync(dev, (size_t)dev_pitch * sizeof(float), host, (size_t)dims * sizeof(float),
                                 (size_t)dims * sizeof(float), (size_t)dims * dims, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return 0;
}

extern "C" int pylb_overdensity(float *grid, int64_t n, double *scratch, void *stream) {
    PYLB_REQUIRE(n > 0 && scratch != nullptr, "pylb_overdensity: empty grid or NULL scratch");
    PYLB_REQUIRE(((uintptr_t)grid & 15) == 0, "pylb_overdensity: grid must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    PYLB_CHECK(cudaMemsetAsync(scratch, 0, 2 * sizeof(double), st));
    sum_kernel<<<grid_for(n / 4 + 1, 256, 8), 256, 0, st>>>(grid, n, scratch);
    PYLB_LAUNCH_CHECK();
    overdensity_kernel<<<grid_for(n / 4 + 1, 256, 8), 256, 0, st>>>(grid, n, scratch, n);
    PYLB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pylb_swap_axes(const float *in, float *out, int dims, int axis, int64_t out_pitch, void *stream) {
    PYLB_REQUIRE(in && out && in != out && dims >= 1 && out_pitch >= dims, "pylb_swap_axes: bad arguments");
    PYLB_REQUIRE(axis >= 0 && axis <= 2, "pylb_swap_axes: axis must be 0, 1 or 2");
    cudaStream_t st = (cudaStream_t)stream;
    if (axis == 2) {
        const int64_t nrows = (int64_t)dims * dims;
        copy_pitched_kernel<<<grid_for(nrows * dims, 256, 16), 256, 0, st>>>(in, out, dims, out_pitch, nrows);
    } else {
        PYLB_REQUIRE(dims <= 65535, "pylb_swap_axes: dims too large for the launch grid");
        dim3 grid((dims + 31) / 32, (dims + 31) / 32, dims);
        if (axis == 1) swap_axes_kernel<1><<<grid, 256, 0, st>>>(in, out, dims, out_pitch);
        else swap_axes_kernel<0><<<grid, 256, 0, st>>>(in, out, dims, out_pitch);
    }
    PYLB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pylb_grid_sum(const float *grid, int64_t n, double *sum, void *stream) {
    PYLB_REQUIRE(grid && sum && n > 0, "pylb_grid_sum: bad arguments");
    PYLB_REQUIRE(((uintptr_t)grid & 15) == 0, "pylb_grid_sum: grid must be 16-byte aligned");
    sum_kernel<<<grid_for(n / 4 + 1, 256, 8), 256, 0, (cudaStream_t)stream>>>(grid, n, sum);
    PYLB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pylb_overdensity_apply(float *grid, int64_t n, const double *sum, int64_t n_total, void *stream) {
    PYLB_REQUIRE(grid && sum && n > 0 && n_total > 0, "pylb_overdensity_apply: bad arguments");
    PYLB_REQUIRE(((uintptr_t)grid & 15) == 0, "pylb_overdensity_apply: grid must be 16-byte aligned");
    overdensity_kernel<<<grid_for(n / 4 + 1, 256, 8), 256, 0, (cudaStream_t)stream>>>(grid, n, sum, n_total);
    PYLB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pylb_pos_redshift_space(float *pos, const float *vel, int64_t np, float box, float hubble,
                                       float redshift, int axis, void *stream) {
    PYLB_REQUIRE(axis >= 0 && axis < 3, "pylb_pos_redshift_space: axis must be 0, 1 or 2");
    if (np == 0) return 0;
    const float factor = (float)((1.0 + (double)redshift) / (double)hubble);
    rsd_kernel<<<grid_for(np, 256, 16), 256, 0, (cudaStream_t)stream>>>(pos, vel, np, box, factor, axis);
    PYLB_LAUNCH_CHECK();
    return 0;
}
