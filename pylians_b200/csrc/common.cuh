// Shared helpers for the pylians_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/pylians_b200.h"

namespace pylb {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
// optional CUDA-event brackets around the dominant kernels (see pylb_timing_enable)
void timing_begin(int which, cudaStream_t st);
void timing_end(int which, cudaStream_t st);

#define PYLB_CHECK(expr)                                                                          \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            pylb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return 1;                                                                             \
        }                                                                                         \
    } while (0)

#define PYLB_LAUNCH_CHECK()                                                                       \
    do {                                                                                          \
        pylb::count_launch();                                                                     \
        cudaError_t _e = cudaGetLastError();                                                      \
        if (_e != cudaSuccess) {                                                                  \
            pylb::set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return 1;                                                                             \
        }                                                                                         \
    } while (0)

#define PYLB_REQUIRE(cond, ...)                                                                   \
    do {                                                                                          \
        if (!(cond)) {                                                                            \
            pylb::set_error(__VA_ARGS__);                                                         \
            return 1;                                                                             \
        }                                                                                         \
    } while (0)

// The library's scratch comes from the stream-ordered allocator.  Its default pool hands unused memory back to
// the driver at every synchronisation (release threshold 0), and Pk() synchronises once per call to read the bins:
// without this the next call's cudaMallocAsync goes back to the driver (sporadic 30-700 ms stalls were measured).
// This raises the release threshold of the DEVICE'S DEFAULT POOL, i.e. for the whole process: memory that this
// library (or the host application) frees with cudaFreeAsync stays cached in the pool instead of going back to the
// driver.  Set PYLB_KEEP_POOL=0 to leave the pool's threshold alone.
inline void keep_pool_memory() {
    static bool done[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
    done[dev] = true;
    const char *e = getenv("PYLB_KEEP_POOL");
    if (e && e[0] == '0') return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done[dev] = true;
}

// Stream-ordered scratch that is released on every way out of a function, error returns included.
struct ScratchGuard {
    cudaStream_t st;
    void *p[8];
    int n = 0;
    explicit ScratchGuard(cudaStream_t s) : st(s) {}
    void add(void *q) { if (q && n < 8) p[n++] = q; }
    ~ScratchGuard() { for (int i = 0; i < n; i++) cudaFreeAsync(p[i], st); }
    ScratchGuard(const ScratchGuard &) = delete;
    ScratchGuard &operator=(const ScratchGuard &) = delete;
};

inline int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// floor(sqrt(n)) exactly, n < 2^31.  fp32 sqrt is only the first guess; the integer fix-up makes
// the result independent of rounding mode / fast-math (Nmodes is a bit-exact contract).
__host__ __device__ __forceinline__ int isqrt_exact(int n) {
#ifdef __CUDA_ARCH__
    int s = (int)__fsqrt_rn((float)n);
#else
    int s = (int)sqrtf((float)n);
#endif
    while ((long long)s * s > (long long)n) --s;
    while ((long long)(s + 1) * (s + 1) <= (long long)n) ++s;
    return s;
}

// signed wavenumber of FFT index i (Pk_library.pyx:315): i > middle ? i - dims : i
__host__ __device__ __forceinline__ int wavenumber(int i, int dims, int middle) {
    return (i > middle) ? i - dims : i;
}

// non-negative remainder; equals the reference's (i+dims)%dims on its valid domain and stays in
// range where the reference would read/write out of bounds.  Inside the valid domain the index is at
// most one period away, so two predicated adds replace the ~25-instruction runtime modulo; the modulo
// only runs for positions far outside the box.
__device__ __forceinline__ int wrap(int i, int dims) {
    if (i < 0) i += dims;
    else if (i >= dims) i -= dims;
    if ((unsigned)i >= (unsigned)dims) {
        i %= dims;
        if (i < 0) i += dims;
    }
    return i;
}

__device__ __forceinline__ void red_add(float *p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void red_add(double *p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void red_add(double *p, float v) { atomicAdd(p, (double)v); }
__device__ __forceinline__ void red_add_u64(uint64_t *p, uint64_t v) {
    atomicAdd(reinterpret_cast<unsigned long long *>(p), (unsigned long long)v);
}

// Warp-level reduce-by-key for binning kernels whose neighbouring lanes mostly share a bin: for each distinct key
// in the warp the NV values are summed by shuffles and `flush(key, sums)` runs on one lane.  Every lane of the
// warp must call it (full-mask shuffles); lanes without a contribution pass valid = false.
template <int NV, class FLUSH>
__device__ __forceinline__ void warp_reduce_by_key(int key, bool valid, const double (&v)[NV], FLUSH flush) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    unsigned todo = __ballot_sync(full, valid);
    while (todo) {
        const int leader = __ffs(todo) - 1;
        const int k = __shfl_sync(full, key, leader);
        const bool mine = valid && key == k;
        double s[NV];
#pragma unroll
        for (int q = 0; q < NV; q++) {
            s[q] = mine ? v[q] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s[q] += __shfl_xor_sync(full, s[q], o);
        }
        if (lane == leader) flush(k, s);
        todo &= ~__ballot_sync(full, mine);
    }
}

}  // namespace pylb
