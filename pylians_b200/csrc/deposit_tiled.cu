// Tiled particle deposit for 3-D float32 grids.
//
// Particles are first brought into cell-tile order, then every tile is accumulated in shared memory
// and flushed with red.global.add.v4.f32.  Two ways to get tile order:
//
//  BINSORT (ntiles <= 32768; tiles are 16x16x32 cells, or 32x32x32 when that is needed to stay under
//           the limit) -- a single-pass counting sort that moves the (x,y,z,w) payload itself:
//     bin_hist_kernel     one CTA per SM builds a histogram over ALL tiles in shared memory (native int
//                         ATOMS.ADD, 2.6 T/s) and merges it into the global per-tile counts
//     cub ExclusiveSum    counts -> first output slot of every tile
//     bin_scatter_kernel  per 32k-particle sub-chunk: count per tile in shared memory, reserve one
//                         contiguous range per (sub-chunk, tile) with a single atom.global, rank inside
//                         it with shared atomics, write the payload as float4.  Streaming reads, ~0.4
//                         global atomics per particle, no random gather (a random 12-byte gather costs
//                         3.9 ms per 2^27 particles on B200, a streaming read 0.25 ms: profiles/microbench).
//  RADIX (any ntiles) -- cub radix sort of (tile key, particle index); the tile kernel then gathers
//     particles through the sorted index.
//
//  deposit_tile_kernel   one CTA per work item (tile, chunk of <= CHUNK particles): zero the tile
//                         (+ halo) in shared memory, accumulate, flush.  Halo cells overlap neighbouring
//                         tiles, so the flush must add -- and `number` is accumulate-in-place anyway.
//
// Shared-memory fp32 atomicAdd is an ATOMS.CAST.SPIN loop on sm_100a (2.9 updates/clk/SM measured vs
// 9.2 for native int atomics); that pipe, not HBM, bounds the accumulation.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "deposit.cuh"

namespace pylb {

template <int TX_, int TY_, int TZ_, int THREADS_>
struct TileCfg {
    static constexpr int TX = TX_, TY = TY_, TZ = TZ_, THREADS = THREADS_;
};
typedef TileCfg<16, 16, 32, 256> TileS;    // 38-51 KB of shared memory per CTA, 4-5 CTAs/SM
typedef TileCfg<32, 32, 32, 1024> TileL;   // 144-176 KB, 1 CTA/SM of 32 warps, 4x fewer tiles

constexpr int CHUNK = 8192;              // particles per work item
constexpr int REGROUP_BATCH = 1024;      // tile kernel: particles regrouped by shared-memory bank at a time
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
constexpr int BIN_PER_THREAD = 16;                            // particles per thread per sub-chunk
constexpr int BIN_SUB = BIN_THREADS * BIN_PER_THREAD;         // 16384 particles per sub-chunk (a round of 148 CTAs must fit L2)
constexpr int BIN_MLP = 8;                                    // particles whose loads are in flight per thread

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

// Scatter the (x,y,z,w) payload into tile order.  cursor[t] starts at tile_begin[t].  Per sub-chunk:
//   sweep A  key every particle, count per tile in shared memory (native ATOMS.ADD);
//   claim    one atom.global.add per tile present in the sub-chunk reserves a contiguous output range
//            (ranges of one tile handed to different CTAs are adjacent, so L2 write-combines them);
//   sweep B  re-read the particles (L1/L2 hits), rank inside the range with a shared atomic, write float4.
template <int MAS, class TC, bool HASW>
__global__ void __launch_bounds__(BIN_THREADS, 1)
bin_scatter_kernel(const float *__restrict__ pos, const float *__restrict__ W, int64_t wst, int64_t first, int n, int64_t ps0,
                   int64_t ps1, float inv, TileGeom tg, int *__restrict__ cursor, float4 *__restrict__ out) {
    extern __shared__ int slot[];
    const int nsub = (n + BIN_SUB - 1) / BIN_SUB;
    for (int sc = blockIdx.x; sc < nsub; sc += gridDim.x) {
        const int lo = sc * BIN_SUB;
        for (int t = threadIdx.x; t < tg.ntiles; t += BIN_THREADS) slot[t] = 0;
        __syncthreads();
        // sweep A: BIN_MLP particles' coordinates are loaded before any is used (memory-level parallelism)
        unsigned short keys[BIN_PER_THREAD];   // ntiles <= 32768, 0xffff = no particle
#pragma unroll
        for (int k0 = 0; k0 < BIN_PER_THREAD; k0 += BIN_MLP) {
            float px[BIN_MLP], py[BIN_MLP], pz[BIN_MLP];
#pragma unroll
            for (int u = 0; u < BIN_MLP; u++) {
                const int i = lo + (k0 + u) * BIN_THREADS + threadIdx.x;
                if (i < n) {
                    const float *p = pos + (first + i) * ps0;
                    px[u] = __ldg(p); py[u] = __ldg(p + ps1); pz[u] = __ldg(p + 2 * ps1);
                }
            }
#pragma unroll
            for (int u = 0; u < BIN_MLP; u++) {
                const int i = lo + (k0 + u) * BIN_THREADS + threadIdx.x;
                keys[k0 + u] = 0xffffu;
                if (i < n) {
                    keys[k0 + u] = (unsigned short)tile_key<MAS, TC>(px[u], py[u], pz[u], inv, tg);
                    atomicAdd(&slot[keys[k0 + u]], 1);
                }
            }
        }
        __syncthreads();
        // claim: atom.global in groups of 8 issued back to back, results stored afterwards
        for (int t0 = 0; t0 < tg.ntiles; t0 += 8 * BIN_THREADS) {
            int base[8];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int t = t0 + q * BIN_THREADS + threadIdx.x;
                base[q] = -1;
                if (t < tg.ntiles) {
                    const int c = slot[t];
                    if (c) base[q] = atomicAdd(&cursor[t], c);
                }
            }
#pragma unroll
            for (int q = 0; q < 8; q++)
                if (base[q] >= 0) slot[t0 + q * BIN_THREADS + threadIdx.x] = base[q];
        }
        __syncthreads();
        // sweep B: re-read (L1/L2), rank inside the claimed range, write the payload
#pragma unroll
        for (int k0 = 0; k0 < BIN_PER_THREAD; k0 += BIN_MLP) {
            float4 v[BIN_MLP];
#pragma unroll
            for (int u = 0; u < BIN_MLP; u++) {
                if (keys[k0 + u] != 0xffffu) {
                    const int i = lo + (k0 + u) * BIN_THREADS + threadIdx.x;
                    const float *p = pos + (first + i) * ps0;
                    v[u] = make_float4(__ldg(p), __ldg(p + ps1), __ldg(p + 2 * ps1), HASW ? __ldg(W + (first + i) * wst) : 1.0f);
                }
            }
#pragma unroll
            for (int u = 0; u < BIN_MLP; u++)
                if (keys[k0 + u] != 0xffffu) out[atomicAdd(&slot[keys[k0 + u]], 1)] = v[u];
        }
        __syncthreads();
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

__global__ void tile_chunks_kernel(const int *__restrict__ tile_begin, int ntiles, int *nchunks) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    nchunks[t] = (t < ntiles) ? (tile_begin[t + 1] - tile_begin[t] + CHUNK - 1) / CHUNK : 0;
}

// ------------------------------------------------------------------------------------------------
// tile accumulation
// ------------------------------------------------------------------------------------------------
template <int MAS, class TC>
struct TileShape {
    static constexpr int S = Support<MAS>::S;
    static constexpr int SX = TC::TX + S - 1, SY = TC::TY + S - 1;
    static constexpr int SZ = ((TC::TZ + S - 1) + 3) & ~3;  // padded to a multiple of 4 for the v4 flush
    static constexpr int CELLS = SX * SY * SZ;
    static constexpr int HI_WORDS = ((CELLS / 2) + 3) & ~3;   // fixed-point tiles: 16-bit carry counters, two per word
};

__device__ __forceinline__ void red_add_v4(float *p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// SORTED: particles come as float4 (x,y,z,w) already in tile order.  Otherwise through the sorted index.
// FIXED (unweighted only): the tile is accumulated as 48-bit fixed point, unit 2^-31, with NATIVE 32-bit shared
//   atomics: `lo` takes the update (ATOMS.ADD with return), a wrap-around of `lo` adds one to a 16-bit carry counter
//   packed two per word in `hi`.  A weight is in [0,1], so an update is at most 2^31 and a CTA's <= 8192 particles
//   can never overflow the 16-bit carry.  Rounding: 2.3e-10 absolute per update (fp32 accumulation rounds every
//   partial sum to 6e-8 relative), the sum itself is exact and order independent -- the tile result is
//   deterministic.  Why: atomicAdd(float) on shared memory is an ATOMS.CAST.SPIN loop on sm_100a (2.9
//   updates/clk/SM measured against 9.2 for native integer atomics).
template <int MAS, bool HASW, class TC, bool SORTED, bool FIXED, bool REGROUP>
__global__ void __launch_bounds__(TC::THREADS)
deposit_tile_kernel(const float *__restrict__ pos, int64_t first, int64_t ps0, int64_t ps1,
                    const float *__restrict__ W, int64_t wst, float inv, TileGeom tg, const unsigned *__restrict__ svals,
                    const float4 *__restrict__ sorted, const int *__restrict__ tile_begin,
                    const int *__restrict__ chunk_off, float *__restrict__ grid, int agg) {
    using TS = TileShape<MAS, TC>;
    constexpr int S = TS::S;
    constexpr int TILE_THREADS = TC::THREADS;
    extern __shared__ __align__(16) float tile[];
    __shared__ int s_tile, s_lo, s_hi, s_agg;

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
            // a tile holding clearly more particles than the average tile is where a halo sits: only its CTAs
            // pay for looking for lanes that share a cell (agg = that particle count, 0 = never)
            s_agg = agg > 0 && tile_begin[lo + 1] - tile_begin[lo] >= agg;
        }
    }
    __syncthreads();
    const int t = s_tile;
    if (t < 0) return;  // CTA-uniform: beyond the last work item
    static_assert(!(FIXED && HASW), "fixed-point accumulation needs weights in [0,1]");
    unsigned *lo = reinterpret_cast<unsigned *>(tile);
    unsigned *hi = lo + TS::CELLS;             // HI_WORDS words
    for (int i = threadIdx.x; i < (FIXED ? (TS::CELLS + TS::HI_WORDS) / 4 : TS::CELLS / 4); i += TILE_THREADS)
        reinterpret_cast<float4 *>(tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int tz = t % tg.ntz, ty = (t / tg.ntz) % tg.nty, tx = t / (tg.ntz * tg.nty);
    const int ox = tx * TC::TX, oy = ty * TC::TY, oz = tz * TC::TZ;

    auto load = [&](int i) -> float4 {
        if (SORTED) return __ldg(sorted + i);
        const int64_t pi = first + (int64_t)svals[i];
        const float *p = pos + pi * ps0;
        return make_float4(__ldg(p), __ldg(p + ps1), __ldg(p + 2 * ps1), HASW ? __ldg(W + pi * wst) : 1.0f);
    };
    // tile-local cell of the particle's lowest touched grid point, or -1 (mis-routed particle: its updates are dropped)
    auto base_cell = [&](const float4 q, float (&C)[3][S]) -> int {
        const int lx = wrap(axis_stencil<MAS>(q.x, inv, C[0]) - tg.x0, tg.dims) - ox;
        if (lx < 0 || lx >= TC::TX) return -1;
        const int ly = wrap(axis_stencil<MAS>(q.y, inv, C[1]), tg.dims) - oy;
        const int lz = wrap(axis_stencil<MAS>(q.z, inv, C[2]), tg.dims) - oz;
        return (lx * TS::SY + ly) * TS::SZ + lz;
    };
    auto put = [&](const int cell0, const float (&C)[3][S], const float w) {
        float *base = tile + cell0;
#pragma unroll
        for (int l = 0; l < S; l++)
#pragma unroll
            for (int m = 0; m < S; m++) {
                const float cxy = C[0][l] * C[1][m];
#pragma unroll
                for (int n = 0; n < S; n++) {
                    float v = cxy * C[2][n];
                    if (HASW) v *= w;
                    if (FIXED) {
                        const int c = cell0 + (l * TS::SY + m) * TS::SZ + n;
                        const unsigned u = __float2uint_rn(v * 2147483648.0f);
                        const unsigned old = atomicAdd(lo + c, u);
                        if (old + u < old) atomicAdd(hi + (c >> 1), 1u << ((c & 1) * 16));
                    } else {
                        atomicAdd(base + (l * TS::SY + m) * TS::SZ + n, v);
                    }
                }
            }
    };
    if (!REGROUP) {
        // (hoisting 4 particle loads ahead of the updates changes nothing for CIC and costs PCS 15 % in registers:
        //  the kernel waits on the shared-memory pipe, not on these loads)
        if (FIXED || S > 2 || !s_agg) {
            for (int i = s_lo + threadIdx.x; i < s_hi; i += TILE_THREADS) {
                const float4 q = load(i);
                float C[3][S];
                const int cell0 = base_cell(q, C);
                if (cell0 >= 0) put(cell0, C, q.w);
            }
        } else {
            // Warp-aggregated updates for clustered inputs.  Lanes whose particles share a base cell update the same
            // S^3 addresses; a shared atomicAdd(float) is a CAS loop, so n lanes on one address cost ~n rounds each.
            // When a group of >= AGG_MIN lanes shares its base cell (a halo cell holding a large share of the tile's
            // particles) the group's S^3 contributions are summed with warp shuffles and its first lane issues ONE
            // atomic per cell.  Only CTAs of over-populated tiles (s_agg) run this loop, so uniform inputs never pay for it,
            // and only NGP and CIC do: S^3 x 5 shuffles per group cost more than the contention they remove for TSC / PCS
            // (measured on a clustered 512^3 set: CIC tile kernel 3.86 -> 2.99 ms, PCS 21.6 -> 25.0 ms).
            constexpr int AGG_MIN = 4;
            const unsigned full = 0xffffffffu;
            const int lane = threadIdx.x & 31;
            for (int i0 = s_lo + (threadIdx.x & ~31); i0 < s_hi; i0 += TILE_THREADS) {   // warp-uniform trip count
                const int i = i0 + lane;
                float C[3][S];
                int cell0 = -1;
                float w = 1.0f;
                if (i < s_hi) {
                    const float4 q = load(i);
                    cell0 = base_cell(q, C);
                    w = q.w;
                }
                const unsigned peers = __match_any_sync(full, cell0);
                const bool heavy = cell0 >= 0 && __popc(peers) >= AGG_MIN;
                unsigned todo = __ballot_sync(full, heavy);
                while (todo) {                                   // one round per heavy group: at most 32 / AGG_MIN
                    const int leader = __ffs(todo) - 1;
                    const unsigned grp = __shfl_sync(full, peers, leader);
                    const bool mine = (grp >> lane) & 1u;
                    const int cell_l = __shfl_sync(full, cell0, leader);
#pragma unroll
                    for (int l = 0; l < S; l++)
#pragma unroll
                        for (int m = 0; m < S; m++) {
                            const float cxy = mine ? C[0][l] * C[1][m] : 0.0f;
#pragma unroll
                            for (int n = 0; n < S; n++) {
                                float v = mine ? cxy * C[2][n] : 0.0f;
                                if (HASW && mine) v *= w;
#pragma unroll
                                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(full, v, o);
                                if (lane == leader) atomicAdd(tile + cell_l + (l * TS::SY + m) * TS::SZ + n, v);
                            }
                        }
                    todo &= ~grp;
                }
                if (cell0 >= 0 && !heavy) put(cell0, C, w);
            }
        }
    } else {
        // Regroup by shared-memory bank.  The tile's particles arrive in no particular order, so the 32 cells a warp
        // updates at once fall into random banks: ~13 shared-memory wavefronts per warp update (LDS + CAS, each
        // serialised ~3.4x), and that data pipe is what bounds this kernel (95 % busy).  Here a batch of REGROUP_BATCH
        // particles is first bucketed by the bank of its base cell (one native shared atomic per particle, a 32-entry
        // scan, one 16-byte shared store); warp r then takes the r-th particle of every bank, lane = bank.  All
        // lanes of an update hit distinct banks -- every stencil offset shifts all of them alike -- and never the
        // same address: 2.7x the update rate in profiles/microbench/atomics.cu (830 -> 2200 G updates/s).
        constexpr int BATCH_P = REGROUP_BATCH;             // 1024 particles: 16 KB of staging
        constexpr int RG = BATCH_P / TILE_THREADS;
        float4 *pstage = reinterpret_cast<float4 *>(tile + (FIXED ? TS::CELLS + TS::HI_WORDS : TS::CELLS));
        __shared__ int bcnt[32], boff[32], s_rows;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int b0 = s_lo; b0 < s_hi; b0 += BATCH_P) {
            if (threadIdx.x < 32) bcnt[threadIdx.x] = 0;
            __syncthreads();
            float4 q[RG];
            int bank[RG], rk[RG];
#pragma unroll
            for (int k = 0; k < RG; k++) {
                const int i = b0 + k * TILE_THREADS + threadIdx.x;
                bank[k] = -1;
                if (i < s_hi) q[k] = load(i);
            }
#pragma unroll
            for (int k = 0; k < RG; k++) {
                const int i = b0 + k * TILE_THREADS + threadIdx.x;
                if (i < s_hi) {
                    float C[3][S];
                    const int cell0 = base_cell(q[k], C);
                    if (cell0 >= 0) {
                        bank[k] = cell0 & 31;
                        rk[k] = atomicAdd(&bcnt[bank[k]], 1);
                    }
                }
            }
            __syncthreads();
            if (threadIdx.x < 32) {            // exclusive scan of the 32 counts, and the longest list
                const int c = bcnt[lane];
                int incl = c, mx = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int y = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += y;
                    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                }
                boff[lane] = incl - c;
                if (lane == 0) s_rows = mx;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < RG; k++)
                if (bank[k] >= 0) pstage[boff[bank[k]] + rk[k]] = q[k];
            __syncthreads();
            const int mycnt = bcnt[lane], myoff = boff[lane], rows = s_rows;
            for (int r = warp; r < rows; r += TILE_THREADS / 32) {
                if (r < mycnt) {
                    const float4 p = pstage[myoff + r];
                    float C[3][S];
                    const int cell0 = base_cell(p, C);
                    put(cell0, C, p.w);
                }
            }
            __syncthreads();                   // the next batch reuses bcnt / pstage
        }
    }
    __syncthreads();
    // fixed point -> float in place (cell i: carry << 32 | lo, unit 2^-31), then the common flush
    if (FIXED) {
        for (int i = threadIdx.x; i < TS::CELLS; i += TILE_THREADS) {
            const unsigned long long t = ((unsigned long long)((hi[i >> 1] >> ((i & 1) * 16)) & 0xffffu) << 32) | lo[i];
            tile[i] = (float)t * 4.656612873077393e-10f;
        }
        __syncthreads();
    }

    // flush: local (x,y,z) -> global ((ox+x)%dims, (oy+y)%dims, (oz+z)%dims)
    const int dims = tg.dims;
    if ((dims & 3) == 0) {
        constexpr int ZV = TS::SZ / 4;
        for (int i = threadIdx.x; i < TS::SX * TS::SY * ZV; i += TILE_THREADS) {
            const int zv = i % ZV, y = (i / ZV) % TS::SY, x = i / (ZV * TS::SY);
            const float4 v = reinterpret_cast<const float4 *>(tile)[i];
            if (v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) continue;
            int gx = ox + x, gy = oy + y, gz = oz + zv * 4;
            if (tg.xext == dims) { if (gx >= dims) gx -= dims; }
            else if (gx >= tg.xext) continue;   // beyond the x window: nothing was deposited there
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
            int gx = ox + x;
            if (tg.xext == dims) gx %= dims;
            else if (gx >= tg.xext) continue;
            atomicAdd(grid + ((int64_t)gx * dims + (oy + y) % dims) * dims + (oz + z) % dims, v);
        }
    }
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
    if (G <= PART_MAXBINS) {
        // the block-local counting sort of the tiled deposit's pass 0 with the slab as its digit: one coalesced read
        // of the particles, one contiguous run per (chunk, slab) written out (1.26 ms against 1.74 ms for the two-sweep
        // scatter at 512^3 particles, G = 8)
        constexpr int PT = 256, CH = PT * PART_PER_THREAD;
        const size_t psm = sizeof(PartSmem<PT>);
        if (set_smem(bin_pass_kernel<MAS, TileS, HASW, true, PT>, psm)) return 1;
        bin_pass_kernel<MAS, TileS, HASW, true, PT><<<(unsigned)((n + CH - 1) / CH), PT, psm, st>>>(
            pos, w, wst, 0, n, ps0, ps1, inv, tg, nullptr, out, cursor, nullptr, nullptr, 0, G);
    } else {
        bin_scatter_kernel<MAS, TileS, HASW><<<P, BIN_THREADS, hist_smem, st>>>(pos, w, wst, 0, n, ps0, ps1, inv, tg, cursor, out);
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
static bool g_two_pass = true;  // binsort payload movement: two-pass block-local sort (default) or the one-pass scatter
static int g_fixed = -1;        // tile accumulation: -1 default (float, or PYLB_MA_FIXED), 0 float, 1 fixed point
void ma_tiled_force_path(int p) {
    g_fixed = -1;
    if (p >= 200) { g_fixed = 1; p -= 200; }          // 2xx: force fixed-point tiles (unweighted deposits)
    else if (p >= 100) { g_fixed = 0; p -= 100; }     // 1xx: force float tiles
    g_two_pass = !(p >= 10);
    g_force_path = p >= 10 ? p - 10 : p;
}
static bool use_fixed(int64_t np, int dims, int xext) {
    static int env = -2;
    if (env == -2) { const char *e = getenv("PYLB_MA_FIXED"); env = e ? atoi(e) : -1; }
    const int f = g_fixed >= 0 ? g_fixed : env;
    (void)np; (void)dims; (void)xext;
    // Opt-in (PYLB_MA_FIXED=1 or pylb_ma_debug_path(2xx)): measured 1.68 ms against 1.59 ms for the float CAS loop at
    // 512^3 CIC -- both saturate the shared-memory data pipe (95 % of LSU wavefronts, ~13 wavefronts per warp
    // update from bank conflicts of randomly placed cells), so the native atomic buys determinism, not speed.
    // Its 2.3e-10 absolute rounding per update needs a mean density well above 1e-4 particles per cell.
    return f > 0;
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

// Particle count from which a tile's CTAs use the warp-aggregated branch: 1.25x the mean tile population (Poisson
// fluctuations of a uniform set at >= 1000 particles per tile stay below 1.1x).  PYLB_MA_AGG=0 switches it off, =2 forces
// it for every tile (A/B runs).
static int agg_threshold(int n, int ntiles) {
    static int env = -2;
    if (env == -2) { const char *e = getenv("PYLB_MA_AGG"); env = e ? atoi(e) : 1; }
    if (env == 0) return 0;
    if (env == 2) return 1;
    const double mean = (double)n / (double)(ntiles > 0 ? ntiles : 1);
    const double thr = 1.25 * mean + 64.0;
    return thr > 2.0e9 ? 2000000000 : (int)thr;
}

static bool use_regroup() {
    static int env = -2;
    // Opt-in (PYLB_MA_REGROUP=1).  Measured at 512^3: CIC 2.31 ms against 1.59 ms in arrival order, PCS 1.50 against
    // 1.54 ms at 256^3 -- the update rate does go up 2.7x, but three barriers, the 32-counter ranking and the exposed
    // particle load per 1024-particle batch cost more than the conflicts they remove (profiles/r1_ma_regroup_ab.txt).
    if (env == -2) { const char *e = getenv("PYLB_MA_REGROUP"); env = e ? atoi(e) : 0; }
    return env != 0;
}

template <int MAS, bool HASW, class TC, bool BINSORT, bool FIXED>
static int tiled_run(const float *pos, int64_t np, int64_t ps0, int64_t ps1, float *grid, int dims, float inv,
                     const float *w, int64_t wst, int x0, int xext, TiledWs &ws, cudaStream_t st) {
    using TS = TileShape<MAS, TC>;
    const TileGeom tg = tile_geom<TC>(dims, x0, xext);
    // NGP has one update per particle: nothing to gain from regrouping
    const bool regroup = MAS != PYLB_NGP && use_regroup();
    const size_t acc_smem = sizeof(float) * (FIXED ? TS::CELLS + TS::HI_WORDS : TS::CELLS);
    const size_t tile_smem = acc_smem + (regroup ? sizeof(float4) * REGROUP_BATCH : 0);
    const size_t hist_smem = sizeof(int) * (size_t)tg.ntiles;
    const int P = sm_count();
    if (set_smem(deposit_tile_kernel<MAS, HASW, TC, BINSORT, FIXED, true>, acc_smem + sizeof(float4) * REGROUP_BATCH) ||
        set_smem(deposit_tile_kernel<MAS, HASW, TC, BINSORT, FIXED, false>, acc_smem)) return 1;
    if (BINSORT) {
        // always the maximum these kernels may ever need: the attribute is a limit, and a smaller value set
        // here would make a later, larger launch of the same instantiation fail
        if (set_smem(bin_hist_kernel<MAS, TC>, sizeof(int) * (size_t)BIN_MAX_TILES)) return 1;
        if (set_smem(bin_scatter_kernel<MAS, TC, HASW>, sizeof(int) * (size_t)BIN_MAX_TILES)) return 1;
    }
    const int nt1 = tg.ntiles + 1;
    for (int64_t first = 0; first < np; first += BATCH) {
        const int n = (int)((np - first) < BATCH ? (np - first) : BATCH);
        size_t tb = ws.tmp_bytes;
        if (BINSORT) {
            PYLB_CHECK(cudaMemsetAsync(ws.H, 0, sizeof(int) * (size_t)nt1, st));
            bin_hist_kernel<MAS, TC><<<P, BIN_THREADS, hist_smem, st>>>(pos, first, n, ps0, ps1, inv, tg, ws.H);
            PYLB_LAUNCH_CHECK();
            // tile_begin[0..ntiles] = exclusive scan of the counts (counts[ntiles] = 0)
            PYLB_CHECK(cub::DeviceScan::ExclusiveSum(ws.tmp, tb, ws.H, ws.tile_begin, nt1, st));
            count_launch(2);
            PYLB_CHECK(cudaMemcpyAsync(ws.S, ws.tile_begin, sizeof(int) * (size_t)nt1, cudaMemcpyDeviceToDevice, st));
            if (g_two_pass) {
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
                bin_scatter_kernel<MAS, TC, HASW><<<P, BIN_THREADS, hist_smem, st>>>(pos, w, wst, first, n, ps0, ps1, inv, tg,
                                                                                      ws.S, ws.sorted);
                PYLB_LAUNCH_CHECK();
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
        tile_chunks_kernel<<<(nt1 + 255) / 256, 256, 0, st>>>(ws.tile_begin, tg.ntiles, ws.nchunks);
        PYLB_LAUNCH_CHECK();
        tb = ws.tmp_bytes;
        PYLB_CHECK(cub::DeviceScan::ExclusiveSum(ws.tmp, tb, ws.nchunks, ws.chunk_off, nt1, st));
        count_launch(2);
        // upper bound on work items: every non-empty tile has at most count/CHUNK + 1 chunks
        const int64_t max_items = (int64_t)n / CHUNK + tg.ntiles;
        timing_begin(PYLB_T_TILE, st);
        if (regroup)
            deposit_tile_kernel<MAS, HASW, TC, BINSORT, FIXED, true><<<(unsigned)max_items, TC::THREADS, tile_smem, st>>>(
                pos, first, ps0, ps1, w, wst, inv, tg, ws.v1, ws.sorted, ws.tile_begin, ws.chunk_off, grid, agg_threshold(n, tg.ntiles));
        else
            deposit_tile_kernel<MAS, HASW, TC, BINSORT, FIXED, false><<<(unsigned)max_items, TC::THREADS, tile_smem, st>>>(
                pos, first, ps0, ps1, w, wst, inv, tg, ws.v1, ws.sorted, ws.tile_begin, ws.chunk_off, grid, agg_threshold(n, tg.ntiles));
        timing_end(PYLB_T_TILE, st);
        PYLB_LAUNCH_CHECK();
    }
    return 0;
}

template <int MAS, bool HASW>
static int tiled_path(int path, const float *pos, int64_t np, int64_t ps0, int64_t ps1, float *grid, int dims,
                      float inv, const float *w, int64_t wst, int x0, int xext, TiledWs &ws, cudaStream_t st) {
    if constexpr (!HASW) {
        // TileL + fixed point would need 264 KB of shared memory: fixed point only with the small tile
        if (path != PATH_BIN_L && use_fixed(np, dims, xext < 0 ? dims : xext)) {
            if (path == PATH_BIN_S) return tiled_run<MAS, HASW, TileS, true, true>(pos, np, ps0, ps1, grid, dims, inv, w, wst, x0, xext, ws, st);
            return tiled_run<MAS, HASW, TileS, false, true>(pos, np, ps0, ps1, grid, dims, inv, w, wst, x0, xext, ws, st);
        }
    }
    if (path == PATH_BIN_S) return tiled_run<MAS, HASW, TileS, true, false>(pos, np, ps0, ps1, grid, dims, inv, w, wst, x0, xext, ws, st);
    if (path == PATH_BIN_L) return tiled_run<MAS, HASW, TileL, true, false>(pos, np, ps0, ps1, grid, dims, inv, w, wst, x0, xext, ws, st);
    return tiled_run<MAS, HASW, TileS, false, false>(pos, np, ps0, ps1, grid, dims, inv, w, wst, x0, xext, ws, st);
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
