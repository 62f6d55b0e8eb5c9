// Tiled particle deposit for 3-D float32 grids.
//
// Particles are first brought into cell-tile order, then every tile is accumulated in shared memory
// and flushed with red.global.add.v4.f32.  Two ways to get tile order:
//
//  BINSORT (ntiles <= 53248; tiles are 16x16x32 cells, or 32x32x32 when that is needed to stay under
//           the limit) -- a counting sort that moves the (x,y,z,w) payload itself:
//     bin_hist_kernel     one CTA per SM builds a histogram over ALL tiles in shared memory (native int
//                         ATOMS.ADD, 2.6 T/s) and merges it into the global per-tile counts
//     cub ExclusiveSum    counts -> first output slot of every tile
//     bin_pass_kernel x2  two-pass block-local counting sort (hi digit, then lo digit of the tile id); streaming
//                         reads, one contiguous output run per (chunk, digit), no random gather (a random 12-byte
//                         gather costs 3.9 ms per 2^27 particles on B200, a streaming read 0.25 ms: profiles/microbench)
//  RADIX (any ntiles) -- cub radix sort of (tile key, particle index); the tile kernel then gathers
//     particles through the sorted index.
//
//  deposit_lane_kernel   CIC/TSC/PCS: one CTA per work item (tile, chunk of <= CHUNK particles): zero the tile
//                         (+ halo) in shared memory, accumulate with the lanes of a warp mapped to the STENCIL POINTS
//                         of one particle (bank-conflict-free by construction of the tile pitches), flush.  Halo cells
//                         overlap neighbouring tiles, so the flush must add -- and `number` is accumulate-in-place anyway.
//  deposit_tile_kernel   NGP (and the A/B baseline): lane per particle.
//
// Shared-memory fp32 atomicAdd is an ATOMS.CAST.SPIN loop on sm_100a (2.9 updates/clk/SM measured for randomly
// placed cells vs 9.2 for native int atomics); the lane-per-particle kernel is bound by that pipe, which is what the
// stencil-lane layout removes.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "deposit.cuh"

namespace pylb {

// TX x TY x TZ cells per tile; PB = particles per warp batch of the stencil-lane kernel (bounds its weight staging);
// CHUNK = particles per work item
template <int TX_, int TY_, int TZ_, int THREADS_, int PB_, int CHUNK_>
struct TileCfg {
    static constexpr int TX = TX_, TY = TY_, TZ = TZ_, THREADS = THREADS_, PB = PB_, CHUNK = CHUNK_;
};
typedef TileCfg<16, 16, 32, 256, 32, 8192> TileS;      // 39-75 KB of shared memory per CTA, 3-5 CTAs/SM
typedef TileCfg<32, 32, 32, 1024, 16, 32768> TileL;    // 144-223 KB, 1 CTA/SM of 32 warps, 4x fewer tiles

constexpr int64_t BATCH = 1ll << 28;     // particles binned per pass (bounds the workspace)
constexpr int BIN_THREADS = 1024;        // binsort CTAs: one per SM, 32 warps
constexpr int BIN_MAX_TILES = 53248;     // per-CTA histogram must fit shared memory (208 KB of 227 KB); keys are 16-bit

// x0 / xext: x window held by the grid (planes x0 .. x0+xext-1 modulo dims; the whole cube when xext == dims)
struct TileGeom {
    int dims, ntx, nty, ntz, ntiles, x0, xext;
    int slab_w;   // > 0: partition mode -- the key is the x-slab (of slab_w planes) owning the particle's lowest touched cell
};

template <class TC>
static TileGeom tile_geom(int dims, int x0 = 0, int xext = -1) {
    TileGeom t;
    t.dims = dims;
    t.x0 = x0;
    t.slab_w = 0;
    t.xext = xext < 0 ? dims : xext;
    t.ntx = (t.xext + TC::TX - 1) / TC::TX;
    t.nty = (dims + TC::TY - 1) / TC::TY;
    t.ntz = (dims + TC::TZ - 1) / TC::TZ;
    t.ntiles = t.ntx * t.nty * t.ntz;
    return t;
}

// key of the tile holding the particle's lowest touched cell
template <int MAS, class TC>
__device__ __forceinline__ unsigned tile_key(float x, float y, float z, float inv, const TileGeom &tg) {
    float C[Support<MAS>::S];
    int bx = wrap(axis_stencil<MAS>(x, inv, C) - tg.x0, tg.dims);
    if (tg.slab_w > 0) return (unsigned)(bx / tg.slab_w);
    if (bx >= tg.xext) bx = tg.xext - 1;   // particle routed to the wrong slab: keep the key in range (its updates are dropped)
    const int by = wrap(axis_stencil<MAS>(y, inv, C), tg.dims);
    const int bz = wrap(axis_stencil<MAS>(z, inv, C), tg.dims);
    return (unsigned)(((bx / TC::TX) * tg.nty + (by / TC::TY)) * tg.ntz + (bz / TC::TZ));
}

// ------------------------------------------------------------------------------------------------
// BINSORT
// ------------------------------------------------------------------------------------------------
// per-tile particle counts: per-CTA shared histogram, merged with one red.global per (CTA, tile)
template <int MAS, class TC>
__global__ void __launch_bounds__(BIN_THREADS, 1)
bin_hist_kernel(const float *__restrict__ pos, int64_t first, int n, int64_t ps0, int64_t ps1, float inv,
                TileGeom tg, int *__restrict__ counts) {
    extern __shared__ int hist[];
    for (int t = threadIdx.x; t < tg.ntiles; t += BIN_THREADS) hist[t] = 0;
    __syncthreads();
    const float *base = pos + first * ps0;
    if (ps0 == 3 && ps1 == 1 && ((uintptr_t)base & 15) == 0) {
        // dense (np,3) array: 4 particles = 3 aligned float4, 8 particles (6 x 16 B) in flight per thread
        const float4 *p4 = reinterpret_cast<const float4 *>(base);
        const int n4 = n >> 2;
        const int stride = gridDim.x * BIN_THREADS;
        for (int q0 = blockIdx.x * BIN_THREADS + threadIdx.x; q0 < n4; q0 += 2 * stride) {
            float4 a[2], b[2], c[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int q = q0 + u * stride;
                if (q < n4) { a[u] = __ldg(p4 + 3 * (int64_t)q); b[u] = __ldg(p4 + 3 * (int64_t)q + 1); c[u] = __ldg(p4 + 3 * (int64_t)q + 2); }
            }
#pragma unroll
            for (int u = 0; u < 2; u++)
                if (q0 + u * stride < n4) {
                    atomicAdd(&hist[tile_key<MAS, TC>(a[u].x, a[u].y, a[u].z, inv, tg)], 1);
                    atomicAdd(&hist[tile_key<MAS, TC>(a[u].w, b[u].x, b[u].y, inv, tg)], 1);
                    atomicAdd(&hist[tile_key<MAS, TC>(b[u].z, b[u].w, c[u].x, inv, tg)], 1);
                    atomicAdd(&hist[tile_key<MAS, TC>(c[u].y, c[u].z, c[u].w, inv, tg)], 1);
                }
        }
        if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
            const float *p = base + 3 * (int64_t)(4 * n4 + threadIdx.x);
            atomicAdd(&hist[tile_key<MAS, TC>(p[0], p[1], p[2], inv, tg)], 1);
        }
    } else if (ps0 == 4 && ps1 == 1 && ((uintptr_t)base & 15) == 0) {
        // packed (x,y,z,w) records (the particle-exchange payload): one 16-byte load per particle, 4 in flight
        const float4 *p4 = reinterpret_cast<const float4 *>(base);
        const int64_t stride = (int64_t)gridDim.x * BIN_THREADS;
        for (int64_t i0 = (int64_t)blockIdx.x * BIN_THREADS + threadIdx.x; i0 < n; i0 += 4 * stride) {
            float4 q[4];
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (i0 + u * stride < n) q[u] = __ldg(p4 + i0 + u * stride);
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (i0 + u * stride < n) atomicAdd(&hist[tile_key<MAS, TC>(q[u].x, q[u].y, q[u].z, inv, tg)], 1);
        }
    } else {
    // 4 particles per iteration: all 12 loads are issued before the first key is computed
    const int64_t stride = (int64_t)gridDim.x * BIN_THREADS;
    for (int64_t i0 = (int64_t)blockIdx.x * BIN_THREADS + threadIdx.x; i0 < n; i0 += 4 * stride) {
        float x[4], y[4], z[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int64_t i = i0 + u * stride;
            if (i < n) {
                const float *p = pos + (first + i) * ps0;
                x[u] = __ldg(p); y[u] = __ldg(p + ps1); z[u] = __ldg(p + 2 * ps1);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (i0 + u * stride < n) atomicAdd(&hist[tile_key<MAS, TC>(x[u], y[u], z[u], inv, tg)], 1);
    }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < tg.ntiles; t += BIN_THREADS) {
        const int c = hist[t];
        if (c) atomicAdd(&counts[t], c);
    }
}

// ------------------------------------------------------------------------------------------------
// Two-pass block-local counting sort of the payload (default for ntiles <= 65536).
//   tile id = (hi digit << lo_bits) | lo digit, both digits <= 256 values.
//   pass 0: raw particles  -> buckets of equal hi digit          (cursor = per-bucket write position)
//   pass 1: bucket by bucket -> tiles (lo digit inside a bucket) (cursor = per-tile write position)
// One CTA sorts a chunk of PART_CHUNK particles in shared memory (rank by a shared atomic per digit, exclusive
// scan of the <= 256 counters), reserves ONE contiguous output range per digit present with a single
// atom.global (<= 256 per 4096 particles), and copies the staged chunk out so that consecutive threads write
// consecutive addresses inside each run.  ~40 B (pass 0) + 32 B (pass 1) of streaming traffic per particle.
// ------------------------------------------------------------------------------------------------
constexpr int PART_PER_THREAD = 8;
constexpr int PART_CHUNK_MAX = 512 * PART_PER_THREAD;        // 4096 particles, 64 KB of float4 staging (512 threads)
constexpr int PART_MAXBINS = 256;

template <int PART_THREADS>
struct PartSmem {
    static constexpr int PART_CHUNK = PART_THREADS * PART_PER_THREAD;
    float4 stage[PART_CHUNK];
    unsigned char dig[PART_CHUNK];
    int cnt[PART_MAXBINS], start[PART_MAXBINS], gbase[PART_MAXBINS];
    int lo, hi, bucket;
};

// chunk list of pass 1: bucket b (tiles [b << lo_bits, (b+1) << lo_bits)) owns ceil(size_b / PART_CHUNK) chunks
__global__ void __launch_bounds__(PART_MAXBINS)
part_buckets_kernel(const int *__restrict__ tile_begin, int ntiles, int lo_bits, int nb0, int *bcursor, int *bchunk_off,
                    int PART_CHUNK) {
    // one thread per bucket (nb0 <= PART_MAXBINS = blockDim.x) and a shared-memory scan of the chunk counts; the
    // single-thread loop this replaces took 38 us of dependent global loads per deposit
    __shared__ int s[PART_MAXBINS];
    const int b = threadIdx.x;
    int c = 0;
    if (b < nb0) {
        const int t0 = b << lo_bits, t1 = min(ntiles, (b + 1) << lo_bits);
        const int begin = tile_begin[t0];
        bcursor[b] = begin;
        c = (tile_begin[t1] - begin + PART_CHUNK - 1) / PART_CHUNK;
    }
    s[b] = c;
    __syncthreads();
    for (int o = 1; o < PART_MAXBINS; o <<= 1) {
        const int v = b >= o ? s[b - o] : 0;
        __syncthreads();
        s[b] += v;
        __syncthreads();
    }
    if (b < nb0) bchunk_off[b] = s[b] - c;
    if (b == nb0 - 1) bchunk_off[nb0] = s[b];
}

template <int MAS, class TC, bool HASW, bool FIRST, int PART_THREADS>
__global__ void __launch_bounds__(PART_THREADS, 1024 / PART_THREADS)
bin_pass_kernel(const float *__restrict__ pos, const float *__restrict__ W, int64_t wst, int64_t first, int n, int64_t ps0,
                int64_t ps1, float inv, TileGeom tg, const float4 *__restrict__ in, float4 *__restrict__ out,
                int *__restrict__ cursor, const int *__restrict__ tile_begin, const int *__restrict__ bchunk_off,
                int lo_bits, int nb0) {
    extern __shared__ __align__(16) unsigned char part_raw[];
    constexpr int PART_CHUNK = PART_THREADS * PART_PER_THREAD;
    PartSmem<PART_THREADS> &sm = *reinterpret_cast<PartSmem<PART_THREADS> *>(part_raw);
    const int tid = threadIdx.x;
    const int nbins = FIRST ? nb0 : (1 << lo_bits);
    if (tid == 0) {
        if (FIRST) {
            sm.lo = blockIdx.x * PART_CHUNK;
            sm.hi = min(n, sm.lo + PART_CHUNK);
            sm.bucket = 0;
        } else {
            const int blk = blockIdx.x, total = bchunk_off[nb0];
            if (blk >= total) { sm.lo = sm.hi = 0; sm.bucket = 0; }
            else {
                int lo = 0, hi = nb0;          // bchunk_off[lo] <= blk < bchunk_off[hi]
                while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (bchunk_off[mid] <= blk) lo = mid; else hi = mid; }
                const int t0 = lo << lo_bits, t1 = min(tg.ntiles, (lo + 1) << lo_bits);
                sm.bucket = lo;
                sm.lo = tile_begin[t0] + (blk - bchunk_off[lo]) * PART_CHUNK;
                sm.hi = min(tile_begin[t1], sm.lo + PART_CHUNK);
            }
        }
    }
    for (int b = tid; b < PART_MAXBINS; b += PART_THREADS) sm.cnt[b] = 0;
    __syncthreads();
    const int lo = sm.lo, hi = sm.hi;
    if (lo >= hi) return;                       // CTA-uniform
    const int cbase = FIRST ? 0 : (sm.bucket << lo_bits);

    float4 v[PART_PER_THREAD];
    int d[PART_PER_THREAD], r[PART_PER_THREAD];
    // dense (np,3) input: a thread takes 2 x 4 consecutive particles as 3 aligned float4 each (lo is a multiple of 4096)
    const float *rawbase = FIRST ? pos + first * ps0 : nullptr;
    const bool vec = FIRST && ps0 == 3 && ps1 == 1 && (((uintptr_t)rawbase) & 15) == 0 && hi - lo == PART_CHUNK;
    auto index_of = [&](int k) { return vec ? lo + ((k >> 2) * PART_THREADS + tid) * 4 + (k & 3) : lo + k * PART_THREADS + tid; };
    if (vec) {
        const float4 *p4 = reinterpret_cast<const float4 *>(rawbase);
        float4 a[2], b[2], c[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int64_t q = ((int64_t)lo >> 2) + u * PART_THREADS + tid;
            a[u] = __ldg(p4 + 3 * q); b[u] = __ldg(p4 + 3 * q + 1); c[u] = __ldg(p4 + 3 * q + 2);
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
            float wv[4] = {1.0f, 1.0f, 1.0f, 1.0f};
            if (HASW) {
#pragma unroll
                for (int j = 0; j < 4; j++) wv[j] = __ldg(W + (first + index_of(4 * u + j)) * wst);
            }
            v[4 * u + 0] = make_float4(a[u].x, a[u].y, a[u].z, wv[0]);
            v[4 * u + 1] = make_float4(a[u].w, b[u].x, b[u].y, wv[1]);
            v[4 * u + 2] = make_float4(b[u].z, b[u].w, c[u].x, wv[2]);
            v[4 * u + 3] = make_float4(c[u].y, c[u].z, c[u].w, wv[3]);
        }
    } else {
    // packed (x,y,z,w) records (the particle-exchange payload): one 16-byte load; the weight rides along when W
    // points at the record's 4th float
    const bool rec4 = FIRST && ps0 == 4 && ps1 == 1 && (((uintptr_t)rawbase) & 15) == 0;
    const bool w_in_rec = HASW && rec4 && wst == 4 && W + first * wst == rawbase + 3;
#pragma unroll
    for (int k = 0; k < PART_PER_THREAD; k++) {
        const int i = lo + k * PART_THREADS + tid;
        if (i < hi) {
            if (FIRST) {
                if (rec4) {
                    const float4 q = __ldg(reinterpret_cast<const float4 *>(rawbase) + i);
                    v[k] = make_float4(q.x, q.y, q.z, HASW ? (w_in_rec ? q.w : __ldg(W + (first + i) * wst)) : 1.0f);
                    continue;
                }
                const float *p = pos + (first + i) * ps0;
                v[k] = make_float4(__ldg(p), __ldg(p + ps1), __ldg(p + 2 * ps1), HASW ? __ldg(W + (first + i) * wst) : 1.0f);
            } else {
                v[k] = __ldg(in + i);
            }
        }
    }
    }
#pragma unroll
    for (int k = 0; k < PART_PER_THREAD; k++) {
        const int i = index_of(k);
        d[k] = -1;
        if (i < hi) {
            const unsigned t = tile_key<MAS, TC>(v[k].x, v[k].y, v[k].z, inv, tg);
            d[k] = FIRST ? (int)(t >> lo_bits) : (int)(t & ((1u << lo_bits) - 1u));
            r[k] = atomicAdd(&sm.cnt[d[k]], 1);
        }
    }
    __syncthreads();
    // exclusive scan of the <= 256 counters by warp 0 (8 per lane) + one global claim per digit present
    if (tid < 32) {
        int loc[8], sum = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) { loc[q] = sum; sum += sm.cnt[tid * 8 + q]; }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += y; }
        const int excl = incl - sum;
#pragma unroll
        for (int q = 0; q < 8; q++) sm.start[tid * 8 + q] = excl + loc[q];
    }
    for (int b = PART_THREADS - 1 - tid; b < PART_MAXBINS; b += PART_THREADS) {   // last warps first: warp 0 is scanning
        const int c = (b < nbins) ? sm.cnt[b] : 0;
        if (c) sm.gbase[b] = atomicAdd(&cursor[cbase + b], c);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PART_PER_THREAD; k++) {
        if (d[k] >= 0) {
            const int p = sm.start[d[k]] + r[k];
            sm.stage[p] = v[k];
            sm.dig[p] = (unsigned char)d[k];
        }
    }
    __syncthreads();
    for (int i = tid; i < hi - lo; i += PART_THREADS) {
        const int dd = sm.dig[i];
        out[sm.gbase[dd] + (i - sm.start[dd])] = sm.stage[i];
    }
}

// ------------------------------------------------------------------------------------------------
// RADIX fallback
// ------------------------------------------------------------------------------------------------
template <int MAS, class TC>
__global__ void __launch_bounds__(256)
tile_key_kernel(const float *__restrict__ pos, int64_t first, int n, int64_t ps0, int64_t ps1, float inv,
                TileGeom tg, unsigned *keys, unsigned *vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *p = pos + (first + i) * ps0;
    keys[i] = tile_key<MAS, TC>(__ldg(p), __ldg(p + ps1), __ldg(p + 2 * ps1), inv, tg);
    vals[i] = (unsigned)i;
}

// tile_begin[t] = first sorted position whose key >= t  (t = 0..ntiles)
__global__ void tile_begin_kernel(const unsigned *__restrict__ skeys, int n, int ntiles, int *tile_begin) {
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

__global__ void tile_chunks_kernel(const int *__restrict__ tile_begin, int ntiles, int chunk, int *nchunks) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    nchunks[t] = (t < ntiles) ? (tile_begin[t + 1] - tile_begin[t] + chunk - 1) / chunk : 0;
}

// ------------------------------------------------------------------------------------------------
// tile accumulation
// ------------------------------------------------------------------------------------------------
constexpr int pad_to(int v, int r) { return v + ((r - v % 32) + 32) % 32; }   // smallest x >= v with x % 32 == r

// Shared-memory tile: SX x SY x SZV cells (tile + S-1 halo cells per axis), z fastest, row pitch PI, plane pitch PL
// (floats).  The pitches are chosen so that the cells ONE warp instruction of deposit_lane_kernel updates fall into 32
// distinct banks (lanes = stencil points, see there):
//   PCS  lane = (a&1, b, c), planes a and a+2   bank = 16 a + 4 b + c          PI = 4, PL = 16 (mod 32)
//   TSC  lane = (a, b, c), 27 lanes             bank = 9 a + 3 b + c           PI = 3, PL = 9
//   CIC  lane = (particle q of 4, a, b, c)      bank = base_q + 4 a + 2 b + c  PI = 2, PL = 4
template <int MAS, class TC>
struct TileShape {
    static constexpr int S = Support<MAS>::S;
    static constexpr int SX = TC::TX + S - 1, SY = TC::TY + S - 1, SZV = TC::TZ + S - 1;
    static constexpr int RI = MAS == PYLB_PCS ? 4 : MAS == PYLB_TSC ? 3 : 2;
    static constexpr int RL = MAS == PYLB_PCS ? 16 : MAS == PYLB_TSC ? 9 : 4;
    static constexpr int PI = MAS == PYLB_NGP ? SZV : pad_to(SZV, RI);
    static constexpr int PL = MAS == PYLB_NGP ? SY * PI : pad_to(SY * PI, RL);
    static constexpr int CELLS = (SX * PL + 3) & ~3;
    static constexpr int ZV = (SZV + 3) / 4;                       // float4 groups per z-row in the flush
    // deposit_lane_kernel: per-warp staging of the factored weights, word-major with pitch PB+1 (bank = word + particle)
    static constexpr int SW = S * S + S + 1;                       // wxy[S][S], wz[S], W
    static constexpr int STAGE_WORDS = (MAS == PYLB_PCS || MAS == PYLB_TSC) ? SW * (TC::PB + 1) : 0;
    static constexpr size_t LANE_SMEM = sizeof(float) * ((size_t)CELLS + (size_t)(TC::THREADS / 32) * STAGE_WORDS);
    static constexpr size_t PLAIN_SMEM = sizeof(float) * (size_t)CELLS;
};

__device__ __forceinline__ void red_add_v4(float *p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// work item b -> (tile, particle range); returns false beyond the last item.  Called by thread 0.
struct WorkItem { int tile, lo, hi; };
__device__ __forceinline__ bool find_work(int b, const int *__restrict__ chunk_off, const int *__restrict__ tile_begin,
                                          int ntiles, int chunk, WorkItem &w) {
    if (b >= chunk_off[ntiles]) return false;
    int lo = 0, hi = ntiles;                    // invariant: chunk_off[lo] <= b < chunk_off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (chunk_off[mid] <= b) lo = mid; else hi = mid;
    }
    w.tile = lo;
    w.lo = tile_begin[lo] + (b - chunk_off[lo]) * chunk;
    w.hi = min(w.lo + chunk, tile_begin[lo + 1]);
    return true;
}

// flush: tile-local (x,y,z) -> global ((ox+x)%dims, (oy+y)%dims, (oz+z)%dims), added with red.global (halo cells overlap
// the neighbouring tiles, and `number` is accumulate-in-place anyway); 16-byte vector reds when dims % 4 == 0
template <class TS, int THREADS>
__device__ __forceinline__ void flush_tile(const float *tile, float *__restrict__ grid, const TileGeom &tg, int ox, int oy, int oz) {
    const int dims = tg.dims;
    if ((dims & 3) == 0) {
        for (int i = threadIdx.x; i < TS::SX * TS::SY * TS::ZV; i += THREADS) {
            const int zv = i % TS::ZV, y = (i / TS::ZV) % TS::SY, x = i / (TS::ZV * TS::SY);
            const float *row = tile + x * TS::PL + y * TS::PI + zv * 4;
            float4 v;
            if constexpr (TS::PI % 4 == 0 && TS::PL % 4 == 0) {
                v = *reinterpret_cast<const float4 *>(row);        // padding words of a row are never written: zero
            } else {
                v.x = row[0];
                v.y = zv * 4 + 1 < TS::SZV ? row[1] : 0.f;
                v.z = zv * 4 + 2 < TS::SZV ? row[2] : 0.f;
                v.w = zv * 4 + 3 < TS::SZV ? row[3] : 0.f;
            }
            if (v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) continue;
            int gx = ox + x, gy = oy + y, gz = oz + zv * 4;
            if (tg.xext == dims) { if (gx >= dims) gx -= dims; }
            else if (gx >= tg.xext) continue;   // beyond the x window: nothing was deposited there
            if (gy >= dims) gy -= dims;
            if (gz >= dims) gz -= dims;         // dims%4==0 and gz%4==0: the 4 cells never straddle the wrap
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
        for (int i = threadIdx.x; i < TS::SX * TS::SY * TS::SZV; i += THREADS) {
            const int z = i % TS::SZV, y = (i / TS::SZV) % TS::SY, x = i / (TS::SZV * TS::SY);
            const float v = tile[x * TS::PL + y * TS::PI + z];
            if (v == 0.f) continue;
            int gx = ox + x;
            if (tg.xext == dims) gx %= dims;
            else if (gx >= tg.xext) continue;
            atomicAdd(grid + ((int64_t)gx * dims + (oy + y) % dims) * dims + (oz + z) % dims, v);
        }
    }
}

// Particle -> (x,y,z,w).  SORTED: float4 records already in tile order; otherwise through the sorted index (radix path).
template <bool SORTED, bool HASW>
struct ParticleSource {
    const float *pos; int64_t first, ps0, ps1; const float *W; int64_t wst; const unsigned *svals; const float4 *sorted;
    __device__ __forceinline__ float4 operator()(int i) const {
        if (SORTED) return __ldg(sorted + i);
        const int64_t pi = first + (int64_t)svals[i];
        const float *p = pos + pi * ps0;
        return make_float4(__ldg(p), __ldg(p + ps1), __ldg(p + 2 * ps1), HASW ? __ldg(W + pi * wst) : 1.0f);
    }
};

// tile-local cell of the particle's lowest touched grid point, or -1 (particle routed to the wrong x window: dropped)
template <int MAS, class TC>
__device__ __forceinline__ int base_cell(const float4 q, float inv, const TileGeom &tg, int ox, int oy, int oz,
                                         float (&C)[3][Support<MAS>::S]) {
    using TS = TileShape<MAS, TC>;
    const int lx = wrap(axis_stencil<MAS>(q.x, inv, C[0]) - tg.x0, tg.dims) - ox;
    if (lx < 0 || lx >= TC::TX) return -1;
    const int ly = wrap(axis_stencil<MAS>(q.y, inv, C[1]), tg.dims) - oy;
    const int lz = wrap(axis_stencil<MAS>(q.z, inv, C[2]), tg.dims) - oz;
    return lx * TS::PL + ly * TS::PI + lz;
}

// K independent shared-memory float adds as ONE compare-and-swap loop: the K loads, adds and CAS are issued back to
// back, so their latencies overlap instead of adding up as in K consecutive atomicAdd(float) loops (ATOMS.CAST.SPIN
// on sm_100a; there is no native shared fp32 add).  Addresses may repeat: the later CAS fails once and retries.
template <int K>
__device__ __forceinline__ void smem_add_joint(float *(&p)[K], const float (&v)[K]) {
    unsigned o[K];
    bool done[K];
#pragma unroll
    for (int k = 0; k < K; k++) { o[k] = *reinterpret_cast<volatile unsigned *>(p[k]); done[k] = false; }
    bool all;
    do {
        all = true;
#pragma unroll
        for (int k = 0; k < K; k++)
            if (!done[k]) {
                const unsigned nv = __float_as_uint(__uint_as_float(o[k]) + v[k]);
                const unsigned r = atomicCAS(reinterpret_cast<unsigned *>(p[k]), o[k], nv);
                done[k] = r == o[k];
                o[k] = r;
                all = all && done[k];
            }
    } while (!all);
}

// ---- lane-per-particle kernel: NGP (one update per particle) and the A/B baseline for the others ----------------
// Every lane streams its own particle and issues S^3 shared atomicAdd(float); the 32 cells of one warp instruction
// fall into random banks (and, for clustered input, onto equal addresses), ~13 shared-memory wavefronts per update.
template <int MAS, bool HASW, class TC, bool SORTED>
__global__ void __launch_bounds__(TC::THREADS)
deposit_tile_kernel(ParticleSource<SORTED, HASW> src, float inv, TileGeom tg, const int *__restrict__ tile_begin,
                    const int *__restrict__ chunk_off, float *__restrict__ grid) {
    using TS = TileShape<MAS, TC>;
    constexpr int S = TS::S;
    extern __shared__ __align__(16) float tile[];
    __shared__ WorkItem s_w;
    __shared__ int s_ok;
    if (threadIdx.x == 0) s_ok = find_work(blockIdx.x, chunk_off, tile_begin, tg.ntiles, TC::CHUNK, s_w);
    __syncthreads();
    if (!s_ok) return;                           // CTA-uniform: beyond the last work item
    const int t = s_w.tile, lo = s_w.lo, hi = s_w.hi;
    for (int i = threadIdx.x; i < TS::CELLS / 4; i += TC::THREADS) reinterpret_cast<float4 *>(tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int tz = t % tg.ntz, ty = (t / tg.ntz) % tg.nty, tx = t / (tg.ntz * tg.nty);
    const int ox = tx * TC::TX, oy = ty * TC::TY, oz = tz * TC::TZ;
    for (int i = lo + threadIdx.x; i < hi; i += TC::THREADS) {
        const float4 q = src(i);
        float C[3][S];
        const int cell0 = base_cell<MAS, TC>(q, inv, tg, ox, oy, oz, C);
        if (cell0 < 0) continue;
#pragma unroll
        for (int l = 0; l < S; l++)
#pragma unroll
            for (int m = 0; m < S; m++) {
                const float cxy = C[0][l] * C[1][m];
#pragma unroll
                for (int n = 0; n < S; n++) {
                    float v = cxy * C[2][n];
                    if (HASW) v *= q.w;
                    atomicAdd(tile + cell0 + l * TS::PL + m * TS::PI + n, v);
                }
            }
    }
    __syncthreads();
    flush_tile<TS, TC::THREADS>(tile, grid, tg, ox, oy, oz);
}

// ---- stencil-lane kernel: CIC, TSC, PCS ------------------------------------------------------------------------
// The lanes of a warp are the STENCIL POINTS of one particle (TSC: 27 lanes; PCS: 32 lanes x 2 planes; CIC: 8 lanes x
// 4 particles), not 32 different particles.  One warp instruction therefore updates cells that are distinct and, thanks
// to the tile pitches above, lie in distinct banks: a shared float add is a single conflict-free CAS round (~3 wavefronts)
// instead of ~13, and a particle costs 1-2 warp instructions' worth of atomics instead of 27/64.  Per batch of PB
// particles, lane p first evaluates particle p's base cell and its S weights per axis exactly like the reference
// (deposit.cuh) and stages wxy[l][m] = C0[l]*C1[m], wz[n] and W in a per-warp scratch (word-major, pitch PB+1: both the
// staging stores and the per-particle reads are conflict-free); then the warp walks the batch, every lane forming
// (wxy * wz) * W for its own stencil point -- the reference's left-to-right fp32 product (MAS_library.pyx:160-166,
// 400-404, 493-497).  Consecutive particles with the same base cell are summed in registers first (JOINT), and the
// adds of two particles share one CAS loop so that their latencies overlap.
template <int MAS, bool HASW, class TC, bool SORTED, bool JOINT>
__global__ void __launch_bounds__(TC::THREADS)
deposit_lane_kernel(ParticleSource<SORTED, HASW> src, float inv, TileGeom tg, const int *__restrict__ tile_begin,
                    const int *__restrict__ chunk_off, float *__restrict__ grid) {
    using TS = TileShape<MAS, TC>;
    constexpr int S = TS::S, PB = TC::PB, NW = TC::THREADS / 32, SP = PB + 1;
    static_assert(MAS != PYLB_NGP, "NGP has one update per particle: lane-per-particle kernel");
    extern __shared__ __align__(16) float tile[];
    __shared__ WorkItem s_w;
    __shared__ int s_ok;
    if (threadIdx.x == 0) s_ok = find_work(blockIdx.x, chunk_off, tile_begin, tg.ntiles, TC::CHUNK, s_w);
    __syncthreads();
    if (!s_ok) return;
    const int t = s_w.tile, lo = s_w.lo, hi = s_w.hi;
    for (int i = threadIdx.x; i < TS::CELLS / 4; i += TC::THREADS) reinterpret_cast<float4 *>(tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int tz = t % tg.ntz, ty = (t / tg.ntz) % tg.nty, tx = t / (tg.ntz * tg.nty);
    const int ox = tx * TC::TX, oy = ty * TC::TY, oz = tz * TC::TZ;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    if constexpr (MAS == PYLB_CIC) {
        // 4 particles per warp step; lane = (q, a, b, c).  The weights of a CIC stencil point are u or 1-u per axis, so
        // the three fractions travel by shuffle and every lane rebuilds its own weight with the reference's fp32 ops.
        const int q = lane >> 3, la = (lane >> 2) & 1, lb = (lane >> 1) & 1, lc = lane & 1;
        const int loff = la * TS::PL + lb * TS::PI + lc;
        for (int i0 = lo + warp * 32; i0 < hi; i0 += NW * 32) {      // warp-uniform trip count
            int cell0 = -1;
            float ux = 0.f, uy = 0.f, uz = 0.f, w = 1.f;
            if (i0 + lane < hi) {
                const float4 p = src(i0 + lane);
                float C[3][S];
                cell0 = base_cell<MAS, TC>(p, inv, tg, ox, oy, oz, C);
                ux = C[0][1]; uy = C[1][1]; uz = C[2][1]; w = p.w;
            }
            auto fetch = [&](int s, float *&addr, float &v) {
                const int from = 4 * s + q;
                const int b = __shfl_sync(full, cell0, from);
                const float fx = __shfl_sync(full, ux, from), fy = __shfl_sync(full, uy, from), fz = __shfl_sync(full, uz, from);
                v = ((la ? fx : __fsub_rn(1.0f, fx)) * (lb ? fy : __fsub_rn(1.0f, fy))) * (lc ? fz : __fsub_rn(1.0f, fz));
                if (HASW) v *= __shfl_sync(full, w, from);
                addr = tile + (b < 0 ? 0 : b) + loff;
                return b >= 0;
            };
            if constexpr (JOINT) {
#pragma unroll
                for (int s = 0; s < 8; s += 2) {
                    float *a0, *a1; float v0, v1;
                    const bool ok0 = fetch(s, a0, v0), ok1 = fetch(s + 1, a1, v1);
                    if (ok0 && ok1) { float *pp[2] = {a0, a1}; const float vv[2] = {v0, v1}; smem_add_joint<2>(pp, vv); }
                    else if (ok0) atomicAdd(a0, v0);
                    else if (ok1) atomicAdd(a1, v1);
                }
            } else {
#pragma unroll
                for (int s = 0; s < 8; s++) {
                    float *a0; float v0;
                    if (fetch(s, a0, v0)) atomicAdd(a0, v0);
                }
            }
        }
    } else {
        constexpr int NP = MAS == PYLB_PCS ? 2 : 1;                 // x-planes per lane
        int la, lb, lc;
        bool active = true;
        if (MAS == PYLB_PCS) { la = lane >> 4; lb = (lane >> 2) & 3; lc = lane & 3; }
        else { active = lane < 27; la = active ? lane / 9 : 0; lb = active ? (lane / 3) % 3 : 0; lc = active ? lane % 3 : 0; }
        const int loff = la * TS::PL + lb * TS::PI + lc;
        float *stage = tile + TS::CELLS + warp * TS::STAGE_WORDS;
        const float *sxy0 = stage + (la * S + lb) * SP, *sxy1 = stage + ((MAS == PYLB_PCS ? la + 2 : la) * S + lb) * SP;   // second x-plane: PCS only
        const float *sz = stage + (S * S + lc) * SP, *sw = stage + (S * S + S) * SP;
        for (int i0 = lo + warp * PB; i0 < hi; i0 += NW * PB) {      // warp-uniform trip count
            int cell0 = -1;
            if (lane < PB && i0 + lane < hi) {
                const float4 p = src(i0 + lane);
                float C[3][S];
                cell0 = base_cell<MAS, TC>(p, inv, tg, ox, oy, oz, C);
                if (cell0 >= 0) {
#pragma unroll
                    for (int l = 0; l < S; l++)
#pragma unroll
                        for (int m = 0; m < S; m++) stage[(l * S + m) * SP + lane] = C[0][l] * C[1][m];
#pragma unroll
                    for (int n = 0; n < S; n++) stage[(S * S + n) * SP + lane] = C[2][n];
                    if (HASW) stage[(S * S + S) * SP + lane] = p.w;
                }
            }
            __syncwarp();
            // value(s) of this lane's stencil point(s) for particle p of the batch
            auto value = [&](int p, float (&v)[NP]) {
                const float wz = sz[p];
                v[0] = sxy0[p] * wz;
                if (NP == 2) v[NP - 1] = sxy1[p] * wz;
                if (HASW) { const float w = sw[p]; v[0] *= w; if (NP == 2) v[NP - 1] *= w; }
            };
            unsigned todo = __ballot_sync(full, cell0 >= 0);
            if constexpr (JOINT) {
                // walk the batch two particles at a time; a run of equal base cells is first summed in registers
                while (todo) {
                    int p = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int b0 = __shfl_sync(full, cell0, p);
                    float v0[NP];
                    if (active) value(p, v0);
                    int b1 = -1;
                    float v1[NP];
                    while (todo) {                                   // warp-uniform
                        p = __ffs(todo) - 1;
                        todo &= todo - 1;
                        b1 = __shfl_sync(full, cell0, p);
                        if (active) value(p, v1);
                        if (b1 != b0) break;
#pragma unroll
                        for (int k = 0; k < NP; k++) v0[k] += v1[k];
                        b1 = -1;
                    }
                    if (active) {
                        if (b1 >= 0) {
                            float *pp[2 * NP]; float vv[2 * NP];
#pragma unroll
                            for (int k = 0; k < NP; k++) {
                                pp[k] = tile + b0 + loff + 2 * k * TS::PL; vv[k] = v0[k];
                                pp[NP + k] = tile + b1 + loff + 2 * k * TS::PL; vv[NP + k] = v1[k];
                            }
                            smem_add_joint<2 * NP>(pp, vv);
                        } else {
                            float *pp[NP]; float vv[NP];
#pragma unroll
                            for (int k = 0; k < NP; k++) { pp[k] = tile + b0 + loff + 2 * k * TS::PL; vv[k] = v0[k]; }
                            smem_add_joint<NP>(pp, vv);
                        }
                    }
                }
            } else {
                while (todo) {
                    const int p = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int b = __shfl_sync(full, cell0, p);
                    if (active) {
                        float v[NP];
                        value(p, v);
#pragma unroll
                        for (int k = 0; k < NP; k++) atomicAdd(tile + b + loff + 2 * k * TS::PL, v[k]);
                    }
                }
            }
            __syncwarp();                                            // the next batch overwrites the staging
        }
    }
    __syncthreads();
    flush_tile<TS, TC::THREADS>(tile, grid, tg, ox, oy, oz);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int bits_for(unsigned v) {
    int b = 1;
    while (b < 32 && (v >> b)) b++;
    return b;
}

enum { PATH_BIN_S = 0, PATH_BIN_L = 1, PATH_RADIX_S = 2 };

// ------------------------------------------------------------------------------------------------
// partition particles by owning x-slab (multi-GPU particle exchange): the binsort kernels with a slab key.
// out[offsets[g] .. offsets[g+1]) = (x,y,z,w) of the particles whose lowest touched x-plane lies in slab g.
// ------------------------------------------------------------------------------------------------
template <class K>
static int set_smem(K kernel, size_t bytes);

template <int MAS, bool HASW>
static int partition_run(const float *pos, int64_t np, int64_t ps0, int64_t ps1, const float *w, int64_t wst, int dims,
                         float inv, int G, float4 *out, int *offsets, cudaStream_t st) {
    keep_pool_memory();
    TileGeom tg = tile_geom<TileS>(dims);
    tg.slab_w = dims / G;
    tg.ntiles = G;
    const int n = (int)np;
    const size_t hist_smem = sizeof(int) * (size_t)(G < 32 ? 32 : G);
    int *counts = nullptr, *cursor = nullptr;
    void *tmp = nullptr;
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, (int *)nullptr, (int *)nullptr, G + 1);
    PYLB_CHECK(cudaMallocAsync(&counts, sizeof(int) * (G + 2), st));
    PYLB_CHECK(cudaMallocAsync(&cursor, sizeof(int) * (G + 2), st));
    PYLB_CHECK(cudaMallocAsync(&tmp, tb ? tb : 16, st));
    PYLB_CHECK(cudaMemsetAsync(counts, 0, sizeof(int) * (G + 2), st));
    const int P = sm_count();
    bin_hist_kernel<MAS, TileS><<<P, BIN_THREADS, hist_smem, st>>>(pos, 0, n, ps0, ps1, inv, tg, counts);
    PYLB_LAUNCH_CHECK();
    PYLB_CHECK(cub::DeviceScan::ExclusiveSum(tmp, tb, counts, offsets, G + 1, st));
    count_launch(2);
    PYLB_CHECK(cudaMemcpyAsync(cursor, offsets, sizeof(int) * (G + 1), cudaMemcpyDeviceToDevice, st));
    PYLB_REQUIRE(G <= PART_MAXBINS, "pylb_partition_xslab: at most %d slabs", PART_MAXBINS);
    {
        // the block-local counting sort of the tiled deposit's pass 0 with the slab as its digit: one coalesced read
        // of the particles, one contiguous run per (chunk, slab) written out
        constexpr int PT = 256, CH = PT * PART_PER_THREAD;
        const size_t psm = sizeof(PartSmem<PT>);
        if (set_smem(bin_pass_kernel<MAS, TileS, HASW, true, PT>, psm)) return 1;
        bin_pass_kernel<MAS, TileS, HASW, true, PT><<<(unsigned)((n + CH - 1) / CH), PT, psm, st>>>(
            pos, w, wst, 0, n, ps0, ps1, inv, tg, nullptr, out, cursor, nullptr, nullptr, 0, G);
    }
    PYLB_LAUNCH_CHECK();
    cudaFreeAsync(counts, st); cudaFreeAsync(cursor, st); cudaFreeAsync(tmp, st);
    return 0;
}

int ma_partition(const float *pos, int64_t np, int64_t ps0, int64_t ps1, const float *w, int64_t wst, int dims, float inv,
                 int mas, int G, float4 *out, int *offsets, cudaStream_t st) {
    const bool hw = w != nullptr;
    switch (mas) {
        case PYLB_NGP: return hw ? partition_run<PYLB_NGP, true>(pos, np, ps0, ps1, w, wst, dims, inv, G, out, offsets, st)
                                 : partition_run<PYLB_NGP, false>(pos, np, ps0, ps1, w, wst, dims, inv, G, out, offsets, st);
        case PYLB_CIC: return hw ? partition_run<PYLB_CIC, true>(pos, np, ps0, ps1, w, wst, dims, inv, G, out, offsets, st)
                                 : partition_run<PYLB_CIC, false>(pos, np, ps0, ps1, w, wst, dims, inv, G, out, offsets, st);
        case PYLB_TSC: return hw ? partition_run<PYLB_TSC, true>(pos, np, ps0, ps1, w, wst, dims, inv, G, out, offsets, st)
                                 : partition_run<PYLB_TSC, false>(pos, np, ps0, ps1, w, wst, dims, inv, G, out, offsets, st);
        case PYLB_PCS: return hw ? partition_run<PYLB_PCS, true>(pos, np, ps0, ps1, w, wst, dims, inv, G, out, offsets, st)
                                 : partition_run<PYLB_PCS, false>(pos, np, ps0, ps1, w, wst, dims, inv, G, out, offsets, st);
    }
    set_error("pylb_partition_xslab: unknown mass-assignment scheme %d", mas);
    return 1;
}

static int g_force_path = -1;   // tests: exercise every path on small grids (pylb_ma_debug_path)
static int g_force_kernel = 0;  // tile kernel: 0 default, 1 lane per particle, 2 stencil lanes (plain atomicAdd), 3 stencil lanes (joint CAS)
void ma_tiled_force_path(int p) {
    g_force_kernel = 0;
    if (p >= 100) { g_force_kernel = p / 100; p %= 100; if (p == 99) p = -1; }   // k99: automatic sort path, kernel k
    g_force_path = p;
}
// PYLB_TILE_KERNEL = 1 | 2 | 3 overrides the default tile kernel (A/B runs)
static int tile_kernel_choice() {
    static int env = -2;
    if (env == -2) { const char *e = getenv("PYLB_TILE_KERNEL"); env = e ? atoi(e) : 0; }
    const int k = g_force_kernel > 0 ? g_force_kernel : env;
    return (k >= 1 && k <= 3) ? k : 3;
}

static int choose_path(int dims, int xext) {
    if (g_force_path >= PATH_BIN_S && g_force_path <= PATH_RADIX_S) return g_force_path;
    if (tile_geom<TileS>(dims, 0, xext).ntiles <= BIN_MAX_TILES) return PATH_BIN_S;
    if (tile_geom<TileL>(dims, 0, xext).ntiles <= BIN_MAX_TILES) return PATH_BIN_L;
    return PATH_RADIX_S;
}

struct TiledWs {
    // binsort
    int *H, *S, *bcursor, *bchunk_off;
    float4 *sorted, *sorted_tmp;
    // radix
    unsigned *k0, *k1, *v0, *v1;
    // common
    int *tile_begin, *nchunks, *chunk_off;
    void *tmp;
    size_t tmp_bytes, total;
};

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static void plan_ws(int64_t np, int dims, int xext, TiledWs *ws, char *base) {
    const int path = choose_path(dims, xext);
    const int ntiles = path == PATH_BIN_L ? tile_geom<TileL>(dims, 0, xext).ntiles : tile_geom<TileS>(dims, 0, xext).ntiles;
    const int64_t nb = np < BATCH ? np : BATCH;
    size_t o = 0, t1 = 0, t2 = 0, t3 = 0;
    auto take = [&](size_t bytes) { char *p = base ? base + o : nullptr; o += align_up(bytes); return p; };
    memset(ws, 0, sizeof(*ws));
    if (path == PATH_RADIX_S) {
        cub::DeviceRadixSort::SortPairs(nullptr, t1, (unsigned *)nullptr, (unsigned *)nullptr, (unsigned *)nullptr,
                                        (unsigned *)nullptr, (int)nb, 0, 32);
        ws->k0 = (unsigned *)take(sizeof(unsigned) * nb);
        ws->k1 = (unsigned *)take(sizeof(unsigned) * nb);
        ws->v0 = (unsigned *)take(sizeof(unsigned) * nb);
        ws->v1 = (unsigned *)take(sizeof(unsigned) * nb);
    } else {
        ws->H = (int *)take(sizeof(int) * (size_t)(ntiles + 2));   // per-tile counts
        ws->S = (int *)take(sizeof(int) * (size_t)(ntiles + 2));   // per-tile write cursors
        ws->sorted = (float4 *)take(sizeof(float4) * nb);
        ws->sorted_tmp = (float4 *)take(sizeof(float4) * nb);
        ws->bcursor = (int *)take(sizeof(int) * (PART_MAXBINS + 2));
        ws->bchunk_off = (int *)take(sizeof(int) * (PART_MAXBINS + 2));
    }
    cub::DeviceScan::ExclusiveSum(nullptr, t2, (int *)nullptr, (int *)nullptr, ntiles + 1);
    ws->tile_begin = (int *)take(sizeof(int) * (ntiles + 2));
    ws->nchunks = (int *)take(sizeof(int) * (ntiles + 2));
    ws->chunk_off = (int *)take(sizeof(int) * (ntiles + 2));
    ws->tmp_bytes = t1 > t2 ? t1 : t2;
    if (t3 > ws->tmp_bytes) ws->tmp_bytes = t3;
    ws->tmp = take(ws->tmp_bytes ? ws->tmp_bytes : 16);
    ws->total = o;
}

size_t ma_tiled_workspace(int64_t np, int dims, int xext, int mas, int has_w) {
    (void)mas; (void)has_w;
    if (np <= 0) return 0;
    TiledWs ws;
    plan_ws(np, dims, xext, &ws, nullptr);
    return ws.total;
}

bool ma_tiled_supported(int ndim, int dims, int grid_f64) { return ndim == 3 && !grid_f64 && dims >= 32; }

template <class K>
static int set_smem(K kernel, size_t bytes) {
    PYLB_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

static int part_threads() {
    static int t = 0;
    // 256-thread CTAs (2048-particle chunks, 4 CTAs/SM) overlap the load / rank / write-out phases of neighbouring CTAs
    // better than 512-thread CTAs (4096, 2 CTAs/SM): 3.73 ms against 3.99 ms for the whole 512^3 CIC deposit
    if (t == 0) { const char *e = getenv("PYLB_PART_THREADS"); const int v = e ? atoi(e) : 256; t = (v == 512 || v == 128) ? v : 256; }
    return t;
}

template <int MAS, class TC, bool HASW, int PT>
static int run_passes(const float *pos, const float *w, int64_t wst, int64_t first, int n, int64_t ps0, int64_t ps1, float inv,
                      const TileGeom &tg, TiledWs &ws, int lo_bits, int nb0, cudaStream_t st) {
    constexpr int CH = PT * PART_PER_THREAD;
    part_buckets_kernel<<<1, PART_MAXBINS, 0, st>>>(ws.tile_begin, tg.ntiles, lo_bits, nb0, ws.bcursor, ws.bchunk_off, CH);
    PYLB_LAUNCH_CHECK();
    const size_t psm = sizeof(PartSmem<PT>);
    if (set_smem(bin_pass_kernel<MAS, TC, HASW, true, PT>, psm) || set_smem(bin_pass_kernel<MAS, TC, HASW, false, PT>, psm)) return 1;
    const unsigned g0 = (unsigned)((n + CH - 1) / CH);
    bin_pass_kernel<MAS, TC, HASW, true, PT><<<g0, PT, psm, st>>>(
        pos, w, wst, first, n, ps0, ps1, inv, tg, nullptr, ws.sorted_tmp, ws.bcursor, ws.tile_begin, ws.bchunk_off, lo_bits, nb0);
    PYLB_LAUNCH_CHECK();
    bin_pass_kernel<MAS, TC, HASW, false, PT><<<g0 + (unsigned)nb0, PT, psm, st>>>(
        pos, w, wst, first, n, ps0, ps1, inv, tg, ws.sorted_tmp, ws.sorted, ws.S, ws.tile_begin, ws.bchunk_off, lo_bits, nb0);
    PYLB_LAUNCH_CHECK();
    return 0;
}

template <int MAS, bool HASW, class TC, bool BINSORT>
static int tiled_run(const float *pos, int64_t np, int64_t ps0, int64_t ps1, float *grid, int dims, float inv,
                     const float *w, int64_t wst, int x0, int xext, TiledWs &ws, cudaStream_t st) {
    using TS = TileShape<MAS, TC>;
    const TileGeom tg = tile_geom<TC>(dims, x0, xext);
    const size_t hist_smem = sizeof(int) * (size_t)tg.ntiles;
    const int P = sm_count();
    const int kern = MAS == PYLB_NGP ? 1 : tile_kernel_choice();
    // always the maximum these kernels may ever need: the attribute is a limit, and a smaller value set here would make
    // a later, larger launch of the same instantiation fail
    if (set_smem(deposit_tile_kernel<MAS, HASW, TC, BINSORT>, TS::PLAIN_SMEM)) return 1;
    if constexpr (MAS != PYLB_NGP) {
        if (set_smem(deposit_lane_kernel<MAS, HASW, TC, BINSORT, false>, TS::LANE_SMEM) ||
            set_smem(deposit_lane_kernel<MAS, HASW, TC, BINSORT, true>, TS::LANE_SMEM)) return 1;
    }
    if (BINSORT) {
        if (set_smem(bin_hist_kernel<MAS, TC>, sizeof(int) * (size_t)BIN_MAX_TILES)) return 1;
    }
    const int nt1 = tg.ntiles + 1;
    for (int64_t first = 0; first < np; first += BATCH) {
        const int n = (int)((np - first) < BATCH ? (np - first) : BATCH);
        size_t tb = ws.tmp_bytes;
        timing_begin(PYLB_T_SORT, st);
        if (BINSORT) {
            PYLB_CHECK(cudaMemsetAsync(ws.H, 0, sizeof(int) * (size_t)nt1, st));
            bin_hist_kernel<MAS, TC><<<P, BIN_THREADS, hist_smem, st>>>(pos, first, n, ps0, ps1, inv, tg, ws.H);
            PYLB_LAUNCH_CHECK();
            // tile_begin[0..ntiles] = exclusive scan of the counts (counts[ntiles] = 0)
            PYLB_CHECK(cub::DeviceScan::ExclusiveSum(ws.tmp, tb, ws.H, ws.tile_begin, nt1, st));
            count_launch(2);
            PYLB_CHECK(cudaMemcpyAsync(ws.S, ws.tile_begin, sizeof(int) * (size_t)nt1, cudaMemcpyDeviceToDevice, st));
            int lo_bits = (bits_for((unsigned)(tg.ntiles - 1)) + 1) / 2;
            if (lo_bits > 8) lo_bits = 8;
            const int nb0 = (tg.ntiles + (1 << lo_bits) - 1) >> lo_bits;      // <= 256 because ntiles <= 65536
            if (part_threads() == 256) {
                if (run_passes<MAS, TC, HASW, 256>(pos, w, wst, first, n, ps0, ps1, inv, tg, ws, lo_bits, nb0, st)) return 1;
            } else if (part_threads() == 128) {
                if (run_passes<MAS, TC, HASW, 128>(pos, w, wst, first, n, ps0, ps1, inv, tg, ws, lo_bits, nb0, st)) return 1;
            } else {
                if (run_passes<MAS, TC, HASW, 512>(pos, w, wst, first, n, ps0, ps1, inv, tg, ws, lo_bits, nb0, st)) return 1;
            }
        } else {
            tile_key_kernel<MAS, TC><<<(n + 255) / 256, 256, 0, st>>>(pos, first, n, ps0, ps1, inv, tg, ws.k0, ws.v0);
            PYLB_LAUNCH_CHECK();
            const int end_bit = bits_for((unsigned)(tg.ntiles - 1));
            PYLB_CHECK(cub::DeviceRadixSort::SortPairs(ws.tmp, tb, ws.k0, ws.k1, ws.v0, ws.v1, n, 0, end_bit, st));
            count_launch(3);
            tile_begin_kernel<<<(nt1 + 255) / 256, 256, 0, st>>>(ws.k1, n, tg.ntiles, ws.tile_begin);
            PYLB_LAUNCH_CHECK();
        }
        tile_chunks_kernel<<<(nt1 + 255) / 256, 256, 0, st>>>(ws.tile_begin, tg.ntiles, TC::CHUNK, ws.nchunks);
        PYLB_LAUNCH_CHECK();
        tb = ws.tmp_bytes;
        PYLB_CHECK(cub::DeviceScan::ExclusiveSum(ws.tmp, tb, ws.nchunks, ws.chunk_off, nt1, st));
        count_launch(2);
        // upper bound on work items: every non-empty tile has at most count/CHUNK + 1 chunks
        const unsigned max_items = (unsigned)((int64_t)n / TC::CHUNK + tg.ntiles);
        const ParticleSource<BINSORT, HASW> src{pos, first, ps0, ps1, w, wst, ws.v1, ws.sorted};
        timing_end(PYLB_T_SORT, st);
        timing_begin(PYLB_T_TILE, st);
        if constexpr (MAS != PYLB_NGP) {
            if (kern == 3)
                deposit_lane_kernel<MAS, HASW, TC, BINSORT, true><<<max_items, TC::THREADS, TS::LANE_SMEM, st>>>(
                    src, inv, tg, ws.tile_begin, ws.chunk_off, grid);
            else if (kern == 2)
                deposit_lane_kernel<MAS, HASW, TC, BINSORT, false><<<max_items, TC::THREADS, TS::LANE_SMEM, st>>>(
                    src, inv, tg, ws.tile_begin, ws.chunk_off, grid);
        }
        if (kern == 1)
            deposit_tile_kernel<MAS, HASW, TC, BINSORT><<<max_items, TC::THREADS, TS::PLAIN_SMEM, st>>>(
                src, inv, tg, ws.tile_begin, ws.chunk_off, grid);
        timing_end(PYLB_T_TILE, st);
        PYLB_LAUNCH_CHECK();
    }
    return 0;
}

template <int MAS, bool HASW>
static int tiled_path(int path, const float *pos, int64_t np, int64_t ps0, int64_t ps1, float *grid, int dims,
                      float inv, const float *w, int64_t wst, int x0, int xext, TiledWs &ws, cudaStream_t st) {
    if (path == PATH_BIN_S) return tiled_run<MAS, HASW, TileS, true>(pos, np, ps0, ps1, grid, dims, inv, w, wst, x0, xext, ws, st);
    if (path == PATH_BIN_L) return tiled_run<MAS, HASW, TileL, true>(pos, np, ps0, ps1, grid, dims, inv, w, wst, x0, xext, ws, st);
    return tiled_run<MAS, HASW, TileS, false>(pos, np, ps0, ps1, grid, dims, inv, w, wst, x0, xext, ws, st);
}

int ma_tiled(const float *pos, int64_t np, int64_t ps0, int64_t ps1, float *grid, int dims, float inv, int mas,
             const float *w, int64_t wst, int x0, int xext, void *workspace, size_t workspace_bytes, cudaStream_t st) {
    if (np == 0) return 0;
    TiledWs ws;
    plan_ws(np, dims, xext, &ws, (char *)workspace);
    PYLB_REQUIRE(workspace != nullptr && workspace_bytes >= ws.total, "pylb_ma: tiled workspace too small (%zu < %zu)",
                 workspace_bytes, ws.total);
    PYLB_REQUIRE(((uintptr_t)grid & 15) == 0, "pylb_ma: grid must be 16-byte aligned");
    const int path = choose_path(dims, xext);
    const bool hw = w != nullptr;
    switch (mas) {
        case PYLB_NGP: return hw ? tiled_path<PYLB_NGP, true>(path, pos, np, ps0, ps1, grid, dims, inv, w, wst, x0, xext, ws, st)
                                 : tiled_path<PYLB_NGP, false>(path, pos, np, ps0, ps1, grid, dims, inv, w, wst, x0, xext, ws, st);
        case PYLB_CIC: return hw ? tiled_path<PYLB_CIC, true>(path, pos, np, ps0, ps1, grid, dims, inv, w, wst, x0, xext, ws, st)
                                 : tiled_path<PYLB_CIC, false>(path, pos, np, ps0, ps1, grid, dims, inv, w, wst, x0, xext, ws, st);
        case PYLB_TSC: return hw ? tiled_path<PYLB_TSC, true>(path, pos, np, ps0, ps1, grid, dims, inv, w, wst, x0, xext, ws, st)
                                 : tiled_path<PYLB_TSC, false>(path, pos, np, ps0, ps1, grid, dims, inv, w, wst, x0, xext, ws, st);
        case PYLB_PCS: return hw ? tiled_path<PYLB_PCS, true>(path, pos, np, ps0, ps1, grid, dims, inv, w, wst, x0, xext, ws, st)
                                 : tiled_path<PYLB_PCS, false>(path, pos, np, ps0, ps1, grid, dims, inv, w, wst, x0, xext, ws, st);
    }
    set_error("pylb_ma: unknown mass-assignment scheme %d", mas);
    return 1;
}

}  // namespace pylb
