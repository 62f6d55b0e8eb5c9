// Tiled particle deposit for 3-D float32 grids.
//
//   K1  tile_key_kernel      particle -> key of the 16x16x32-cell tile holding its lowest touched cell
//       cub radix sort       (key, particle index) pairs, only the key bits that are in use
//       tile_begin_kernel    first sorted position of every tile (binary search)
//       tile_chunks_kernel   + cub exclusive scan: work items = (tile, chunk of <= CHUNK particles),
//                            so heavy (clustered) tiles are split over several CTAs
//   K2  deposit_tile_kernel  one CTA per work item: zero a (16+S-1)x(16+S-1)x(32+S-1) fp32 tile in
//                            shared memory, accumulate the item's particles into it, flush the tile
//                            with red.global.add.v4.f32 (halo cells overlap neighbouring tiles, so
//                            the flush must add, and `number` is accumulate-in-place anyway).
//
// Shared-memory fp32 atomicAdd is an ATOMS.CAST.SPIN loop on sm_100a; the accumulation is bounded by
// the shared-memory pipe (3 LSU ops per update), not by HBM.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "deposit.cuh"

namespace pylb {

constexpr int TX = 16, TY = 16, TZ = 32;
constexpr int CHUNK = 8192;             // particles per work item
constexpr int64_t BATCH = 1ll << 28;    // particles sorted per pass (bounds the workspace)
constexpr int TILE_THREADS = 256;

struct TileGeom {
    int dims, ntx, nty, ntz, ntiles;
};

static TileGeom tile_geom(int dims) {
    TileGeom t;
    t.dims = dims;
    t.ntx = (dims + TX - 1) / TX;
    t.nty = (dims + TY - 1) / TY;
    t.ntz = (dims + TZ - 1) / TZ;
    t.ntiles = t.ntx * t.nty * t.ntz;
    return t;
}

template <int MAS>
__global__ void __launch_bounds__(256)
tile_key_kernel(const float *__restrict__ pos, int64_t first, int n, int64_t ps0, int64_t ps1, float inv,
                TileGeom tg, unsigned *keys, unsigned *vals) {
    constexpr int S = Support<MAS>::S;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *p = pos + (first + i) * ps0;
    float C[S];
    const int bx = wrap(axis_stencil<MAS>(__ldg(p), inv, C), tg.dims);
    const int by = wrap(axis_stencil<MAS>(__ldg(p + ps1), inv, C), tg.dims);
    const int bz = wrap(axis_stencil<MAS>(__ldg(p + 2 * ps1), inv, C), tg.dims);
    keys[i] = (unsigned)(((bx / TX) * tg.nty + (by / TY)) * tg.ntz + (bz / TZ));
    vals[i] = (unsigned)i;
}

// tile_begin[t] = first sorted position whose key >= t  (t = 0..ntiles)
__global__ void tile_begin_kernel(const unsigned *__restrict__ skeys, int n, int ntiles, int *tile_begin, int *nchunks) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (skeys[mid] < (unsigned)t) lo = mid + 1;
        else hi = mid;
    }
    tile_begin[t] = lo;
}

__global__ void tile_chunks_kernel(const int *__restrict__ tile_begin, int ntiles, int *nchunks) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    nchunks[t] = (t < ntiles) ? (tile_begin[t + 1] - tile_begin[t] + CHUNK - 1) / CHUNK : 0;
}

template <int MAS>
struct TileShape {
    static constexpr int S = Support<MAS>::S;
    static constexpr int SX = TX + S - 1, SY = TY + S - 1;
    static constexpr int SZ = ((TZ + S - 1) + 3) & ~3;  // padded to a multiple of 4 for the v4 flush
    static constexpr int CELLS = SX * SY * SZ;
};

__device__ __forceinline__ void red_add_v4(float *p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

template <int MAS, bool HASW>
__global__ void __launch_bounds__(TILE_THREADS)
deposit_tile_kernel(const float *__restrict__ pos, int64_t first, int64_t ps0, int64_t ps1,
                    const float *__restrict__ W, float inv, TileGeom tg, const unsigned *__restrict__ svals,
                    const int *__restrict__ tile_begin, const int *__restrict__ chunk_off,
                    float *__restrict__ grid) {
    using TS = TileShape<MAS>;
    constexpr int S = TS::S;
    extern __shared__ __align__(16) float tile[];
    __shared__ int s_tile, s_lo, s_hi;

    if (threadIdx.x == 0) {
        // find the tile whose chunk range holds blockIdx.x: chunk_off[t] <= b < chunk_off[t+1]
        const int b = blockIdx.x;
        const int total = chunk_off[tg.ntiles];
        if (b >= total) {
            s_tile = -1;
        } else {
            int lo = 0, hi = tg.ntiles;  // invariant: chunk_off[lo] <= b < chunk_off[hi]
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (chunk_off[mid] <= b) lo = mid;
                else hi = mid;
            }
            s_tile = lo;
            const int begin = tile_begin[lo] + (b - chunk_off[lo]) * CHUNK;
            s_lo = begin;
            s_hi = min(begin + CHUNK, tile_begin[lo + 1]);
        }
    }
    __syncthreads();
    const int t = s_tile;
    if (t < 0) return;  // CTA-uniform: beyond the last work item
    for (int i = threadIdx.x; i < TS::CELLS / 4; i += TILE_THREADS)
        reinterpret_cast<float4 *>(tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int tz = t % tg.ntz, ty = (t / tg.ntz) % tg.nty, tx = t / (tg.ntz * tg.nty);
    const int ox = tx * TX, oy = ty * TY, oz = tz * TZ;

    for (int i = s_lo + threadIdx.x; i < s_hi; i += TILE_THREADS) {
        const int64_t pi = first + (int64_t)svals[i];
        const float *p = pos + pi * ps0;
        float C[3][S];
        const int lx = wrap(axis_stencil<MAS>(__ldg(p), inv, C[0]), tg.dims) - ox;
        const int ly = wrap(axis_stencil<MAS>(__ldg(p + ps1), inv, C[1]), tg.dims) - oy;
        const int lz = wrap(axis_stencil<MAS>(__ldg(p + 2 * ps1), inv, C[2]), tg.dims) - oz;
        const float w = HASW ? __ldg(W + pi) : 1.0f;
        float *base = tile + (lx * TS::SY + ly) * TS::SZ + lz;
#pragma unroll
        for (int l = 0; l < S; l++)
#pragma unroll
            for (int m = 0; m < S; m++) {
                const float cxy = C[0][l] * C[1][m];
#pragma unroll
                for (int n = 0; n < S; n++) {
                    float v = cxy * C[2][n];
                    if (HASW) v *= w;
                    atomicAdd(base + (l * TS::SY + m) * TS::SZ + n, v);
                }
            }
    }
    __syncthreads();

    // flush: local (x,y,z) -> global ((ox+x)%dims, (oy+y)%dims, (oz+z)%dims)
    const int dims = tg.dims;
    if ((dims & 3) == 0) {
        constexpr int ZV = TS::SZ / 4;
        for (int i = threadIdx.x; i < TS::SX * TS::SY * ZV; i += TILE_THREADS) {
            const int zv = i % ZV, y = (i / ZV) % TS::SY, x = i / (ZV * TS::SY);
            const float4 v = reinterpret_cast<const float4 *>(tile)[i];
            if (v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) continue;
            int gx = ox + x, gy = oy + y, gz = oz + zv * 4;
            if (gx >= dims) gx -= dims;
            if (gy >= dims) gy -= dims;
            if (gz >= dims) gz -= dims;  // dims%4==0 and gz%4==0: the 4 cells never straddle the wrap
            if (gx >= dims || gy >= dims || gz >= dims) {  // only when dims < tile extent: scalar, full modulo
                const float a[4] = {v.x, v.y, v.z, v.w};
                for (int q = 0; q < 4; q++)
                    if (a[q] != 0.f)
                        atomicAdd(grid + ((int64_t)(gx % dims) * dims + gy % dims) * dims + (gz + q) % dims, a[q]);
                continue;
            }
            red_add_v4(grid + ((int64_t)gx * dims + gy) * dims + gz, v);
        }
    } else {
        for (int i = threadIdx.x; i < TS::CELLS; i += TILE_THREADS) {
            const float v = tile[i];
            if (v == 0.f) continue;
            const int z = i % TS::SZ, y = (i / TS::SZ) % TS::SY, x = i / (TS::SZ * TS::SY);
            atomicAdd(grid + ((int64_t)((ox + x) % dims) * dims + (oy + y) % dims) * dims + (oz + z) % dims, v);
        }
    }
}

static int bits_for(unsigned v) {
    int b = 1;
    while (b < 32 && (v >> b)) b++;
    return b;
}

struct TiledWs {
    unsigned *k0, *k1, *v0, *v1;
    int *tile_begin, *nchunks, *chunk_off;
    void *tmp;
    size_t tmp_bytes, total;
};

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static int plan_ws(int64_t np, int dims, TiledWs *ws, char *base) {
    const TileGeom tg = tile_geom(dims);
    const int64_t nb = np < BATCH ? np : BATCH;
    size_t sort_tmp = 0, scan_tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (unsigned *)nullptr, (unsigned *)nullptr, (unsigned *)nullptr,
                                    (unsigned *)nullptr, (int)nb, 0, 32);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (int *)nullptr, (int *)nullptr, tg.ntiles + 1);
    size_t o = 0;
    auto take = [&](size_t bytes) { char *p = base ? base + o : nullptr; o += align_up(bytes); return p; };
    ws->k0 = (unsigned *)take(sizeof(unsigned) * nb);
    ws->k1 = (unsigned *)take(sizeof(unsigned) * nb);
    ws->v0 = (unsigned *)take(sizeof(unsigned) * nb);
    ws->v1 = (unsigned *)take(sizeof(unsigned) * nb);
    ws->tile_begin = (int *)take(sizeof(int) * (tg.ntiles + 2));
    ws->nchunks = (int *)take(sizeof(int) * (tg.ntiles + 2));
    ws->chunk_off = (int *)take(sizeof(int) * (tg.ntiles + 2));
    ws->tmp_bytes = sort_tmp > scan_tmp ? sort_tmp : scan_tmp;
    ws->tmp = take(ws->tmp_bytes ? ws->tmp_bytes : 16);
    ws->total = o;
    return 0;
}

size_t ma_tiled_workspace(int64_t np, int dims, int mas, int has_w) {
    (void)mas; (void)has_w;
    if (np <= 0) return 0;
    TiledWs ws;
    plan_ws(np, dims, &ws, nullptr);
    return ws.total;
}

bool ma_tiled_supported(int ndim, int dims, int grid_f64) { return ndim == 3 && !grid_f64 && dims >= 32; }

template <int MAS, bool HASW>
static int tiled_run(const float *pos, int64_t np, int64_t ps0, int64_t ps1, float *grid, int dims, float inv,
                     const float *w, TiledWs &ws, cudaStream_t st) {
    using TS = TileShape<MAS>;
    const TileGeom tg = tile_geom(dims);
    const size_t smem = sizeof(float) * TS::CELLS;
    static bool attr_set = false;
    if (!attr_set) {
        PYLB_CHECK(cudaFuncSetAttribute(deposit_tile_kernel<MAS, HASW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int end_bit = bits_for((unsigned)(tg.ntiles - 1));
    for (int64_t first = 0; first < np; first += BATCH) {
        const int n = (int)((np - first) < BATCH ? (np - first) : BATCH);
        tile_key_kernel<MAS><<<(n + 255) / 256, 256, 0, st>>>(pos, first, n, ps0, ps1, inv, tg, ws.k0, ws.v0);
        PYLB_LAUNCH_CHECK();
        size_t tb = ws.tmp_bytes;
        PYLB_CHECK(cub::DeviceRadixSort::SortPairs(ws.tmp, tb, ws.k0, ws.k1, ws.v0, ws.v1, n, 0, end_bit, st));
        count_launch(3);
        const int nt1 = tg.ntiles + 1;
        tile_begin_kernel<<<(nt1 + 255) / 256, 256, 0, st>>>(ws.k1, n, tg.ntiles, ws.tile_begin, ws.nchunks);
        PYLB_LAUNCH_CHECK();
        tile_chunks_kernel<<<(nt1 + 255) / 256, 256, 0, st>>>(ws.tile_begin, tg.ntiles, ws.nchunks);
        PYLB_LAUNCH_CHECK();
        tb = ws.tmp_bytes;
        PYLB_CHECK(cub::DeviceScan::ExclusiveSum(ws.tmp, tb, ws.nchunks, ws.chunk_off, nt1, st));
        count_launch(2);
        // upper bound on work items: every non-empty tile has at most count/CHUNK + 1 chunks
        const int64_t max_items = (int64_t)n / CHUNK + tg.ntiles;
        timing_begin(PYLB_T_TILE, st);
        deposit_tile_kernel<MAS, HASW><<<(unsigned)max_items, TILE_THREADS, smem, st>>>(
            pos, first, ps0, ps1, w, inv, tg, ws.v1, ws.tile_begin, ws.chunk_off, grid);
        timing_end(PYLB_T_TILE, st);
        PYLB_LAUNCH_CHECK();
    }
    return 0;
}

int ma_tiled(const float *pos, int64_t np, int64_t ps0, int64_t ps1, float *grid, int dims, float inv, int mas,
             const float *w, void *workspace, size_t workspace_bytes, cudaStream_t st) {
    if (np == 0) return 0;
    TiledWs ws;
    plan_ws(np, dims, &ws, (char *)workspace);
    PYLB_REQUIRE(workspace != nullptr && workspace_bytes >= ws.total, "pylb_ma: tiled workspace too small (%zu < %zu)",
                 workspace_bytes, ws.total);
    PYLB_REQUIRE(((uintptr_t)grid & 15) == 0, "pylb_ma: grid must be 16-byte aligned");
    const bool hw = w != nullptr;
    switch (mas) {
        case PYLB_NGP: return hw ? tiled_run<PYLB_NGP, true>(pos, np, ps0, ps1, grid, dims, inv, w, ws, st)
                                 : tiled_run<PYLB_NGP, false>(pos, np, ps0, ps1, grid, dims, inv, w, ws, st);
        case PYLB_CIC: return hw ? tiled_run<PYLB_CIC, true>(pos, np, ps0, ps1, grid, dims, inv, w, ws, st)
                                 : tiled_run<PYLB_CIC, false>(pos, np, ps0, ps1, grid, dims, inv, w, ws, st);
        case PYLB_TSC: return hw ? tiled_run<PYLB_TSC, true>(pos, np, ps0, ps1, grid, dims, inv, w, ws, st)
                                 : tiled_run<PYLB_TSC, false>(pos, np, ps0, ps1, grid, dims, inv, w, ws, st);
        case PYLB_PCS: return hw ? tiled_run<PYLB_PCS, true>(pos, np, ps0, ps1, grid, dims, inv, w, ws, st)
                                 : tiled_run<PYLB_PCS, false>(pos, np, ps0, ps1, grid, dims, inv, w, ws, st);
    }
    set_error("pylb_ma: unknown mass-assignment scheme %d", mas);
    return 1;
}

}  // namespace pylb
