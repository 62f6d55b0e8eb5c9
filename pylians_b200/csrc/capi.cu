// C-ABI glue: error channel, launch accounting, pylb_ma dispatch, the reference-compatible host
// entry points (MAS_c.h:3-10) and the slab-transpose pack kernels.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <utility>
#include <vector>

#include "common.cuh"

namespace pylb {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches += n; }

// ---- per-kernel timing ---------------------------------------------------------------------------
struct TimingSlot {
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending, pool;
    cudaEvent_t cur_begin = nullptr;
};
static bool g_timing = false;
static TimingSlot g_slots[PYLB_T_COUNT];

void timing_begin(int which, cudaStream_t st) {
    if (!g_timing) return;
    TimingSlot &s = g_slots[which];
    std::pair<cudaEvent_t, cudaEvent_t> ev;
    if (!s.pool.empty()) { ev = s.pool.back(); s.pool.pop_back(); }
    else { cudaEventCreate(&ev.first); cudaEventCreate(&ev.second); }
    cudaEventRecord(ev.first, st);
    s.pending.push_back(ev);
}
void timing_end(int which, cudaStream_t st) {
    if (!g_timing) return;
    TimingSlot &s = g_slots[which];
    if (!s.pending.empty()) cudaEventRecord(s.pending.back().second, st);
}

int ma_direct(const float *pos, int64_t np, int ndim, int64_t ps0, int64_t ps1, void *grid, int f64,
              int dims, float inv, int mas, const float *w, float zrep, int x0, int xext, int64_t wst, cudaStream_t st);
int ma_tiled(const float *pos, int64_t np, int64_t ps0, int64_t ps1, float *grid, int dims, float inv,
             int mas, const float *w, int64_t wst, int x0, int xext, void *workspace, size_t workspace_bytes, cudaStream_t st);
size_t ma_tiled_workspace(int64_t np, int dims, int xext, int mas, int has_w);
bool ma_tiled_supported(int ndim, int dims, int grid_f64, int xext);
void ma_tiled_force_path(int p);
int ma_partition(const float *pos, int64_t np, int64_t ps0, int64_t ps1, const float *w, int64_t wst, int dims, float inv,
                 int mas, int G, float4 *out, int *offsets, cudaStream_t st);

__global__ void add_f32_kernel(float *__restrict__ dst, const float *__restrict__ src, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] += src[i];
}

// ------------------------------------------------------------------------------------------------
// slab transpose: src [nx][dims][nz] -> dst [G][nx][dims/G][nz]; 16-byte vector copies when the
// row length allows, 8-byte otherwise.  One CTA per (ix, g) block of ny_loc*nz contiguous elements.
// ------------------------------------------------------------------------------------------------
template <typename V>
__global__ void __launch_bounds__(256)
slab_pack_kernel(const V *__restrict__ src, V *__restrict__ dst, long long blk, int nx, int G) {
    // blk = ny_loc*nz in units of V
    const int ix = blockIdx.y, gq = blockIdx.z;
    const V *s = src + ((long long)ix * G + gq) * blk;
    V *d = dst + ((long long)gq * nx + ix) * blk;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < blk; i += (long long)gridDim.x * blockDim.x)
        d[i] = s[i];
}

static int slab_pack_impl(const void *src, void *dst, int dims, int nx, int G, long long pitch, cudaStream_t st) {
    PYLB_REQUIRE(G >= 1 && dims % G == 0 && nx >= 1 && pitch >= dims / 2 + 1, "pylb_slab_pack: dims must be divisible by G");
    const long long blk = (long long)(dims / G) * pitch;   // complex elements per block (rows keep their pitch)
    int bx = (int)((blk + 256 * 8 - 1) / (256 * 8));
    if (bx < 1) bx = 1;
    if (bx > 64) bx = 64;
    dim3 grid(bx, nx, G);
    const bool v16 = (blk % 2 == 0) && (((uintptr_t)src & 15) == 0) && (((uintptr_t)dst & 15) == 0);
    if (v16)
        slab_pack_kernel<float4><<<grid, 256, 0, st>>>((const float4 *)src, (float4 *)dst, blk / 2, nx, G);
    else
        slab_pack_kernel<float2><<<grid, 256, 0, st>>>((const float2 *)src, (float2 *)dst, blk, nx, G);
    PYLB_LAUNCH_CHECK();
    return 0;
}

}  // namespace pylb

using namespace pylb;

extern "C" int pylb_version(void) { return PYLB_VERSION; }
extern "C" const char *pylb_last_error(void) { return g_err; }
extern "C" int64_t pylb_launch_count(void) { return (int64_t)g_launches.load(); }
extern "C" void pylb_timing_enable(int on) { g_timing = on != 0; }
extern "C" int pylb_timing_collect(int which, double *total_ms, int *launches) {
    PYLB_REQUIRE(which >= 0 && which < PYLB_T_COUNT && total_ms && launches, "pylb_timing_collect: bad arguments");
    TimingSlot &s = g_slots[which];
    double tot = 0;
    int n = 0;
    for (auto &ev : s.pending) {
        PYLB_CHECK(cudaEventSynchronize(ev.second));
        float ms = 0;
        PYLB_CHECK(cudaEventElapsedTime(&ms, ev.first, ev.second));
        tot += ms; n++;
        s.pool.push_back(ev);
    }
    s.pending.clear();
    *total_ms = tot; *launches = n;
    return 0;
}

extern "C" void pylb_ma_debug_path(int path) { ma_tiled_force_path(path); }

extern "C" size_t pylb_ma_workspace_bytes(int64_t np, int ndim, int dims, int mas, int has_w, int grid_f64, int algo) {
    if (algo == PYLB_MA_DIRECT) return 0;
    if (!ma_tiled_supported(ndim, dims, grid_f64, dims)) return 0;
    return ma_tiled_workspace(np, dims, dims, mas, has_w);
}

extern "C" int pylb_ma(const float *pos, int64_t np, int ndim, int64_t ps0, int64_t ps1, void *grid, int grid_f64,
                       int dims, float box, int mas, const float *w, int z_repeat, int algo, void *workspace,
                       size_t workspace_bytes, void *stream) {
    PYLB_REQUIRE(ndim == 2 || ndim == 3, "pylb_ma: pos must have 2 or 3 coordinates, got %d", ndim);
    PYLB_REQUIRE(mas >= PYLB_NGP && mas <= PYLB_PCS, "pylb_ma: unknown mass-assignment scheme %d", mas);
    PYLB_REQUIRE(dims >= 1 && np >= 0, "pylb_ma: bad sizes");
    PYLB_REQUIRE(np == 0 || (pos != nullptr && grid != nullptr), "pylb_ma: NULL pointer");
    PYLB_REQUIRE(box > 0.0f, "pylb_ma: BoxSize must be positive");
    const float inv = (float)dims / box;  // `cdef float inv_cell_size = dims/BoxSize`, MAS_library.pyx:135
    const float zrep = (ndim == 2 && z_repeat > 1) ? (float)z_repeat : 1.0f;
    cudaStream_t st = (cudaStream_t)stream;
    bool tiled = false;
    if (algo != PYLB_MA_DIRECT && ma_tiled_supported(ndim, dims, grid_f64, dims)) {
        // AUTO: the tiled path pays a binning pass; it wins once the grid no longer lives in L2
        const bool big = (size_t)dims * dims * dims * sizeof(float) > (size_t)48 << 20;
        tiled = (algo == PYLB_MA_TILED) || (big && np >= (int64_t)1 << 20);
        if (tiled && workspace_bytes < ma_tiled_workspace(np, dims, dims, mas, w != nullptr)) {
            PYLB_REQUIRE(algo != PYLB_MA_TILED, "pylb_ma: workspace too small for the tiled path (%zu < %zu)",
                         workspace_bytes, ma_tiled_workspace(np, dims, dims, mas, w != nullptr));
            tiled = false;
        }
    } else {
        PYLB_REQUIRE(algo != PYLB_MA_TILED, "pylb_ma: tiled path needs a 3-D float32 grid with dims >= 32");
    }
    if (tiled) return ma_tiled(pos, np, ps0, ps1, (float *)grid, dims, inv, mas, w, 1, 0, dims, workspace, workspace_bytes, st);
    return ma_direct(pos, np, ndim, ps0, ps1, grid, grid_f64, dims, inv, mas, w, zrep, 0, dims, 1, st);
}

// ---- x-window deposit: the grid holds planes x0 .. x0+xext-1 (mod dims) of a dims^3 cube -------------
extern "C" size_t pylb_ma_window_workspace_bytes(int64_t np, int dims, int xext, int mas, int algo) {
    if (algo == PYLB_MA_DIRECT || !ma_tiled_supported(3, dims, 0, xext)) return 0;
    return ma_tiled_workspace(np, dims, xext, mas, 1);
}

extern "C" int pylb_ma_window(const float *pos, int64_t np, int64_t ps0, int64_t ps1, float *grid, int dims, int x0,
                              int xext, float box, int mas, const float *w, int64_t w_stride, int algo,
                              void *workspace, size_t workspace_bytes, void *stream) {
    PYLB_REQUIRE(mas >= PYLB_NGP && mas <= PYLB_PCS, "pylb_ma_window: unknown mass-assignment scheme %d", mas);
    PYLB_REQUIRE(dims >= 1 && np >= 0 && x0 >= 0 && x0 < dims && xext >= 1 && xext <= dims, "pylb_ma_window: bad window");
    PYLB_REQUIRE(np == 0 || (pos != nullptr && grid != nullptr), "pylb_ma_window: NULL pointer");
    PYLB_REQUIRE(box > 0.0f, "pylb_ma_window: BoxSize must be positive");
    const float inv = (float)dims / box;
    cudaStream_t st = (cudaStream_t)stream;
    bool tiled = algo != PYLB_MA_DIRECT && ma_tiled_supported(3, dims, 0, xext) && np >= ((int64_t)1 << 16) &&
                 workspace_bytes >= ma_tiled_workspace(np, dims, xext, mas, 1) && workspace != nullptr;
    PYLB_REQUIRE(tiled || algo != PYLB_MA_TILED, "pylb_ma_window: tiled path unavailable (dims < 32, tiny input or workspace too small)");
    if (tiled) return ma_tiled(pos, np, ps0, ps1, grid, dims, inv, mas, w, w_stride, x0, xext, workspace, workspace_bytes, st);
    return ma_direct(pos, np, 3, ps0, ps1, grid, 0, dims, inv, mas, w, 1.0f, x0, xext, w_stride, st);
}

// ---- reference-compatible host entry points (MAS_c.h:3-10) -------------------------------------
// One stream per host thread, created on first use and kept; all device buffers of a call come from the stream-ordered
// pool (no cudaMalloc / cudaFree, no stream creation per call: a caller looping over many small deposits used to pay four
// synchronising allocations each time).
static cudaStream_t masc_stream() {
    static thread_local cudaStream_t st = nullptr;
    if (!st && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) st = nullptr;
    return st;
}

static void masc_host(int mas, float *pos, float *number, float *W, long particles, int dims, int axes, float box) {
    g_err[0] = 0;
    if (axes != 2 && axes != 3) { set_error("MAS_c entry: axes must be 2 or 3"); return; }
    keep_pool_memory();
    const size_t cells = (size_t)dims * dims * (axes == 3 ? dims : 1);
    float *dpos = nullptr, *dgrid = nullptr, *dw = nullptr;
    void *ws = nullptr;
    cudaStream_t st = masc_stream();
    bool ok = st != nullptr;
    ScratchGuard guard(st);
    ok = ok && cudaMallocAsync(&dpos, sizeof(float) * (size_t)particles * axes + 16, st) == cudaSuccess;
    guard.add(dpos);
    ok = ok && cudaMallocAsync(&dgrid, sizeof(float) * cells, st) == cudaSuccess;
    guard.add(dgrid);
    if (ok && W) { ok = cudaMallocAsync(&dw, sizeof(float) * (size_t)particles + 16, st) == cudaSuccess; guard.add(dw); }
    const size_t wsb = pylb_ma_workspace_bytes(particles, axes, dims, mas, W != nullptr, 0, PYLB_MA_AUTO);
    if (ok && wsb) { ok = cudaMallocAsync(&ws, wsb, st) == cudaSuccess; guard.add(ws); }
    if (!ok) { set_error("MAS_c entry: CUDA allocation failed: %s", cudaGetErrorString(cudaGetLastError())); return; }
    cudaMemcpyAsync(dpos, pos, sizeof(float) * (size_t)particles * axes, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(dgrid, number, sizeof(float) * cells, cudaMemcpyHostToDevice, st);
    if (W) cudaMemcpyAsync(dw, W, sizeof(float) * (size_t)particles, cudaMemcpyHostToDevice, st);
    if (pylb_ma(dpos, particles, axes, axes, 1, dgrid, 0, dims, box, mas, dw, 1, PYLB_MA_AUTO, ws, wsb, st) == 0) {
        cudaMemcpyAsync(number, dgrid, sizeof(float) * cells, cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess)
            set_error("MAS_c entry: execution failed: %s", cudaGetErrorString(cudaGetLastError()));
    } else {
        cudaStreamSynchronize(st);               // the host arrays must not be in use by a pending copy when we return
    }
}

extern "C" void NGP(float *pos, float *number, float *W, long particles, int dims, int axes, float BoxSize, int threads) {
    (void)threads; masc_host(PYLB_NGP, pos, number, W, particles, dims, axes, BoxSize);
}
extern "C" void CIC(float *pos, float *number, float *W, long particles, int dims, int axes, float BoxSize, int threads) {
    (void)threads; masc_host(PYLB_CIC, pos, number, W, particles, dims, axes, BoxSize);
}
extern "C" void TSC(float *pos, float *number, float *W, long particles, int dims, int axes, float BoxSize, int threads) {
    (void)threads; masc_host(PYLB_TSC, pos, number, W, particles, dims, axes, BoxSize);
}
extern "C" void PCS(float *pos, float *number, float *W, long particles, int dims, int axes, float BoxSize, int threads) {
    (void)threads; masc_host(PYLB_PCS, pos, number, W, particles, dims, axes, BoxSize);
}

extern "C" int pylb_partition_xslab(const float *pos, int64_t np, int64_t ps0, int64_t ps1, const float *w,
                                    int64_t w_stride, int dims, float box, int mas, int G, void *out_xyzw,
                                    int *offsets, void *stream) {
    PYLB_REQUIRE(G >= 1 && G <= 1024 && dims % G == 0, "pylb_partition_xslab: dims must be divisible by G (1..1024)");
    PYLB_REQUIRE(np >= 0 && np < ((int64_t)1 << 31), "pylb_partition_xslab: particle count out of range");
    PYLB_REQUIRE(offsets != nullptr && (np == 0 || (pos && out_xyzw)), "pylb_partition_xslab: NULL pointer");
    PYLB_REQUIRE(box > 0.0f, "pylb_partition_xslab: BoxSize must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    if (np == 0) { PYLB_CHECK(cudaMemsetAsync(offsets, 0, sizeof(int) * (G + 1), st)); return 0; }
    return ma_partition(pos, np, ps0, ps1, w, w_stride, dims, (float)dims / box, mas, G, (float4 *)out_xyzw, offsets, st);
}

extern "C" int pylb_add_f32(float *dst, const float *src, int64_t n, void *stream) {
    if (n <= 0) return 0;
    PYLB_REQUIRE(dst && src, "pylb_add_f32: NULL pointer");
    int64_t b = (n + 255) / 256;
    if (b > (int64_t)sm_count() * 16) b = (int64_t)sm_count() * 16;
    add_f32_kernel<<<(unsigned)b, 256, 0, (cudaStream_t)stream>>>(dst, src, n);
    PYLB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pylb_slab_pack(const void *src, void *dst, int dims, int nx_local, int G, int64_t pitch, void *stream) {
    PYLB_REQUIRE(src && dst, "pylb_slab_pack: NULL pointer");
    return slab_pack_impl(src, dst, dims, nx_local, G, pitch, (cudaStream_t)stream);
}

