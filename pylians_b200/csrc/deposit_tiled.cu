// Tiled particle deposit for 3-D float32 grids.
//
// Particles are first brought into cell-tile order (tiles of 16x16x32 cells), then every tile is accumulated in
// shared memory and flushed with red.global.add.v4.f32.
//
//  SORT -- an MSD counting sort that moves the (x,y,z,w) payload itself, two digits of the tile id:
//     bin_hist_kernel     one CTA per SM builds a histogram in shared memory (native int ATOMS.ADD) and merges it
//                         into the global counts.  Up to 53248 tiles ("shallow": every grid up to 768^3, and x-slab
//                         windows of larger ones) the keys are the tiles themselves and this one sweep over the raw
//                         particles yields every tile's first output slot.  Beyond that ("deep", up to 2^20 tiles =
//                         a full 2048^3 grid) the sweep histograms the HI digit only, and the per-tile counts come
//                         from bin_hist2_kernel, a second sweep over the bucket-ordered payload after pass 0 (a CTA
//                         then only sees the <= 1024 tiles of one bucket): +16 B per particle, no tile-count limit.
//     cub ExclusiveSum    counts -> first output slot of every bucket / tile
//     bin_pass_kernel x2  block-local counting sort by hi digit, then by lo digit inside every bucket; streaming
//                         reads, one contiguous output run per (chunk, digit), no random gather (a random 12-byte
//                         gather costs 3.9 ms per 2^27 particles on B200, a streaming read 0.25 ms: profiles/microbench)
//
//  ACCUMULATE -- one work item = (tile, chunk of <= 8192 of its particles); persistent CTAs walk the item list:
//     deposit_tile_kernel   NGP/CIC: lane per particle, S^3 shared-memory atomicAdd(float) per lane.
//     deposit_lane_kernel   TSC/PCS: the lanes of a warp are the STENCIL POINTS of a particle (PCS: 32 lanes x 2 x-planes of
//                           one particle; TSC: 27 lanes x 2 consecutive particles) and the tile pitches put the cells of one
//                           instruction into distinct banks, so a warp instruction never touches a cell twice and its
//                           compare-and-swap only fails when another warp got there first: the update is an optimistic
//                           load / add / CAS with no spin loop on the fast path, and a second optimistic CAS on the value the
//                           failed one returned on the slow path.
//   Shared-memory float updates are what bounds every variant (there is no native fp32 add: atomicAdd is an
//   ATOMS.CAST.SPIN loop).  Measured on B200 (profiles/r2_atoms_pattern*.txt, cell updates per clock per SM at 32 warps/SM):
//   float CAS loop 2.8 on randomly placed cells, 5.6 on the stencil pattern (the optimistic pair 5.4, a 64-bit CAS on two
//   adjacent cells 5.8); native integer ATOMS.ADD 6.4 / 8.5; plain LDS+FADD+STS 4.5 / 7.7.  The PCS lane kernel reaches 4.9.
//   Built, measured and dropped this round (profiles/r2_deposit_variants.txt, profiles/variants/): integer (fixed-point)
//   tiles -- one word per cell with a per-item scale loses the sub-unit contributions of dense clumps (the 1e-5 contract
//   fails on clustered input), two words per cell are exact but need twice the atomics (10.7 ms against 8.1 ms then); and
//   column accumulators in registers (lanes = z cells, S^2 FMAs per particle and lane, one add to the tile per column
//   visit) -- no longer bound by shared memory but by its 65 warp instructions per particle (10.9 ms against 5.95 ms).
#include <cub/device/device_scan.cuh>

#include "deposit.cuh"

namespace pylb {

constexpr int TX = 16, TY = 16, TZ = 32;     // cells per tile
constexpr int CHUNK = 8192;                  // particles per work item
#ifndef PYLB_WIDE_PT0
#define PYLB_WIDE_PT0 512                      // CTA size of a sort pass with more than 256 digits
#endif
#ifndef PYLB_LANE_DEPTH
#define PYLB_LANE_DEPTH 2                     // stencil-lane kernel: steps (compare-and-swap pairs) in flight together
#endif
// Particles sorted and deposited per round.  A round flushes every tile it touches (the whole grid plus halo for a
// space-filling input: ~3 x 4 B per cell of red.global traffic), so it should hold about one particle per cell; the
// workspace (two float4 payload buffers, 32 B per particle of a round) bounds it from above.  2^28 particles (8.6 GB) up
// to 640^3 cells, 2^30 (34 GB) from 1024^3 on: 2048^3 particles onto 2048^3 cells went from 47 to 19 ms per 2^28
// particles of tile kernel when the rounds grew from 2^28 to 2^30 (profiles/r2_bench_2048_strong_1gpu*.json).
static int64_t round_particles(int dims, int xext) {
    const int64_t cells = (int64_t)(xext < 0 ? dims : xext) * dims * dims, q = 1ll << 28;
    int64_t r = (cells + q - 1) / q * q;
    if (r < q) r = q;
    if (r > 4 * q) r = 4 * q;
    return r;
}
constexpr int BIN_THREADS = 1024;            // histogram CTAs: one per SM, 32 warps
constexpr int BIN_MAX_KEYS = 53248;          // per-CTA histogram must fit shared memory (208 KB of 227 KB)
constexpr int MAX_TILES = 1 << 20;           // two digits of <= 1024 values
constexpr int FEW_KEYS = 32;                 // at most this many sort keys: count and rank per warp first (match.any)

// x0 / xext: x window held by the grid (planes x0 .. x0+xext-1 modulo dims; the whole cube when xext == dims)
struct TileGeom {
    int dims, ntx, nty, ntz, ntiles, x0, xext;
    int slab_w;   // > 0: partition mode -- the key is the x-slab (of slab_w planes) owning the particle's lowest touched cell
};

static TileGeom tile_geom(int dims, int x0 = 0, int xext = -1) {
    TileGeom t;
    t.dims = dims;
    t.x0 = x0;
    t.slab_w = 0;
    t.xext = xext < 0 ? dims : xext;
    t.ntx = (t.xext + TX - 1) / TX;
    t.nty = (dims + TY - 1) / TY;
    t.ntz = (dims + TZ - 1) / TZ;
    t.ntiles = t.ntx * t.nty * t.ntz;
    return t;
}

// key of the tile holding the particle's lowest touched cell
template <int MAS>
__device__ __forceinline__ unsigned tile_key(float x, float y, float z, float inv, const TileGeom &tg) {
    float C[Support<MAS>::S];
    int bx = wrap(axis_stencil<MAS>(x, inv, C) - tg.x0, tg.dims);
    if (tg.slab_w > 0) return (unsigned)(bx / tg.slab_w);
    if (bx >= tg.xext) bx = tg.xext - 1;   // particle routed to the wrong slab: keep the key in range (its updates are dropped)
    const int by = wrap(axis_stencil<MAS>(y, inv, C), tg.dims);
    const int bz = wrap(axis_stencil<MAS>(z, inv, C), tg.dims);
    return (unsigned)(((bx / TX) * tg.nty + (by / TY)) * tg.ntz + (bz / TZ));
}

// ------------------------------------------------------------------------------------------------
// SORT
// ------------------------------------------------------------------------------------------------
// counts[key >> shift] over the raw particles: per-CTA shared histogram, merged with one red.global per (CTA, key)
template <int MAS>
__global__ void __launch_bounds__(BIN_THREADS, 1)
bin_hist_kernel(const float *__restrict__ pos, int64_t first, int n, int64_t ps0, int64_t ps1, float inv,
                TileGeom tg, int shift, int nkeys, int *__restrict__ counts) {
    extern __shared__ int hist[];
    for (int t = threadIdx.x; t < nkeys; t += BIN_THREADS) hist[t] = 0;
    __syncthreads();
    auto key = [&](float x, float y, float z) { return tile_key<MAS>(x, y, z, inv, tg) >> shift; };
    if (nkeys <= FEW_KEYS) {
        // a handful of keys (the x-slab partition of the multi-GPU exchange): 1024 threads on 8 counters serialise on
        // the address, so every warp first counts its lanes per key (match.any) and adds once per key present
        const int64_t stride = (int64_t)gridDim.x * BIN_THREADS;
        const int lane = threadIdx.x & 31;
        for (int64_t i0 = (int64_t)blockIdx.x * BIN_THREADS + threadIdx.x - lane; i0 < n; i0 += stride) {   // warp-uniform
            const int64_t i = i0 + lane;
            int k = -1;
            if (i < n) {
                const float *p = pos + (first + i) * ps0;
                k = (int)key(__ldg(p), __ldg(p + ps1), __ldg(p + 2 * ps1));
            }
            const unsigned peers = __match_any_sync(0xffffffffu, k);
            if (k >= 0 && lane == __ffs(peers) - 1) atomicAdd(&hist[k], __popc(peers));
        }
        __syncthreads();
        for (int t = threadIdx.x; t < nkeys; t += BIN_THREADS) {
            const int c = hist[t];
            if (c) atomicAdd(&counts[t], c);
        }
        return;
    }
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
                    atomicAdd(&hist[key(a[u].x, a[u].y, a[u].z)], 1);
                    atomicAdd(&hist[key(a[u].w, b[u].x, b[u].y)], 1);
                    atomicAdd(&hist[key(b[u].z, b[u].w, c[u].x)], 1);
                    atomicAdd(&hist[key(c[u].y, c[u].z, c[u].w)], 1);
                }
        }
        if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
            const float *p = base + 3 * (int64_t)(4 * n4 + threadIdx.x);
            atomicAdd(&hist[key(p[0], p[1], p[2])], 1);
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
                if (i0 + u * stride < n) atomicAdd(&hist[key(q[u].x, q[u].y, q[u].z)], 1);
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
                if (i0 + u * stride < n) atomicAdd(&hist[key(x[u], y[u], z[u])], 1);
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nkeys; t += BIN_THREADS) {
        const int c = hist[t];
        if (c) atomicAdd(&counts[t], c);
    }
}

// ------------------------------------------------------------------------------------------------
// Two-pass block-local counting sort of the payload.
//   tile id = (hi digit << lo_bits) | lo digit, both digits <= MAXB values (256, or 1024 for the deep mode).
//   pass 0: raw particles  -> buckets of equal hi digit          (cursor = per-bucket write position)
//   pass 1: bucket by bucket -> tiles (lo digit inside a bucket) (cursor = per-tile write position)
// One CTA sorts a chunk of PT * 8 particles in shared memory (rank by a shared atomic per digit, exclusive
// scan of the counters), reserves ONE contiguous output range per digit present with a single atom.global,
// and copies the staged chunk out so that consecutive threads write consecutive addresses inside each run.
// ~40 B (pass 0) + 32 B (pass 1) of streaming traffic per particle.
// ------------------------------------------------------------------------------------------------
constexpr int PART_PER_THREAD = 8;

template <int PT, int MAXB>
struct PartSmem {
    static constexpr int PART_CHUNK = PT * PART_PER_THREAD;
    float4 stage[PART_CHUNK];
    unsigned short dig[PART_CHUNK];
    int cnt[MAXB], start[MAXB], gbase[MAXB];
    int lo, hi, bucket;
};

// bucket b = entries [b << shift, (b+1) << shift) of `begin` (the tile table in shallow mode, shift = lo_bits; the
// bucket table itself in deep mode, shift = 0).  Writes bucket_begin[0..nb0], the pass-0 cursors and the chunk list of
// pass 1: bucket b owns ceil(size_b / chunk) chunks.
__global__ void __launch_bounds__(1024)
part_buckets_kernel(const int *__restrict__ begin, int n_entries, int shift, int nb0, int chunk, int *bucket_begin,
                    int *bcursor, int *bchunk_off) {
    __shared__ int s[1024];
    const int b = threadIdx.x;
    int c = 0;
    if (b < nb0) {
        const int e0 = b << shift, e1 = min(n_entries, (b + 1) << shift);
        const int lo = begin[e0], hi = begin[e1];
        bucket_begin[b] = lo;
        if (b == nb0 - 1) bucket_begin[nb0] = hi;
        bcursor[b] = lo;
        c = (hi - lo + chunk - 1) / chunk;
    }
    s[b] = c;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int v = b >= o ? s[b - o] : 0;
        __syncthreads();
        s[b] += v;
        __syncthreads();
    }
    if (b < nb0) bchunk_off[b] = s[b] - c;
    if (b == nb0 - 1) bchunk_off[nb0] = s[b];
}

// chunk `blk` of the pass-1 chunk list -> (bucket, particle range)
__device__ __forceinline__ bool bucket_chunk(int blk, const int *__restrict__ bchunk_off, const int *__restrict__ bucket_begin,
                                             int nb0, int chunk, int &bucket, int &lo, int &hi) {
    if (blk >= bchunk_off[nb0]) return false;
    int a = 0, z = nb0;          // bchunk_off[a] <= blk < bchunk_off[z]
    while (z - a > 1) { const int mid = (a + z) >> 1; if (bchunk_off[mid] <= blk) a = mid; else z = mid; }
    bucket = a;
    lo = bucket_begin[a] + (blk - bchunk_off[a]) * chunk;
    hi = min(bucket_begin[a + 1], lo + chunk);
    return true;
}

// deep mode: per-tile counts from the bucket-ordered payload.  A CTA takes one pass-1 chunk (particles of ONE bucket),
// counts its lo digits in shared memory and adds the non-zero counters to the tiles of that bucket.
template <int MAS, int PT>
__global__ void __launch_bounds__(PT)
bin_hist2_kernel(const float4 *__restrict__ in, float inv, TileGeom tg, const int *__restrict__ bucket_begin,
                 const int *__restrict__ bchunk_off, int lo_bits, int nb0, int *__restrict__ counts) {
    __shared__ int cnt[1024];
    __shared__ int s_lo, s_hi, s_bucket;
    if (threadIdx.x == 0) {
        int b = 0, lo = 0, hi = 0;
        bucket_chunk(blockIdx.x, bchunk_off, bucket_begin, nb0, PT * PART_PER_THREAD, b, lo, hi);
        s_bucket = b; s_lo = lo; s_hi = hi;
    }
    for (int b = threadIdx.x; b < 1024; b += PT) cnt[b] = 0;
    __syncthreads();
    const int lo = s_lo, hi = s_hi;
    if (lo >= hi) return;
    const unsigned mask = (1u << lo_bits) - 1u;
    float4 q[PART_PER_THREAD];                    // the chunk holds at most PT * PART_PER_THREAD particles: all loads in flight
#pragma unroll
    for (int k = 0; k < PART_PER_THREAD; k++)
        if (lo + k * PT + (int)threadIdx.x < hi) q[k] = __ldg(in + lo + k * PT + threadIdx.x);
#pragma unroll
    for (int k = 0; k < PART_PER_THREAD; k++)
        if (lo + k * PT + (int)threadIdx.x < hi) atomicAdd(&cnt[tile_key<MAS>(q[k].x, q[k].y, q[k].z, inv, tg) & mask], 1);
    __syncthreads();
    const int cbase = s_bucket << lo_bits;
    for (int b = threadIdx.x; b < (1 << lo_bits); b += PT) {
        const int c = cnt[b];
        if (c) atomicAdd(&counts[cbase + b], c);
    }
}

template <int MAS, bool HASW, bool FIRST, int PT, int MAXB>
__global__ void __launch_bounds__(PT, 1024 / PT)
bin_pass_kernel(const float *__restrict__ pos, const float *__restrict__ W, int64_t wst, int64_t first, int n, int64_t ps0,
                int64_t ps1, float inv, TileGeom tg, const float4 *__restrict__ in, float4 *__restrict__ out,
                int *__restrict__ cursor, const int *__restrict__ bucket_begin, const int *__restrict__ bchunk_off,
                int lo_bits, int nb0) {
    extern __shared__ __align__(16) unsigned char part_raw[];
    constexpr int PART_CHUNK = PT * PART_PER_THREAD;
    PartSmem<PT, MAXB> &sm = *reinterpret_cast<PartSmem<PT, MAXB> *>(part_raw);
    const int tid = threadIdx.x;
    const int nbins = FIRST ? nb0 : (1 << lo_bits);
    if (tid == 0) {
        if (FIRST) {
            sm.lo = blockIdx.x * PART_CHUNK;
            sm.hi = min(n, sm.lo + PART_CHUNK);
            sm.bucket = 0;
        } else {
            int b = 0, lo = 0, hi = 0;
            bucket_chunk(blockIdx.x, bchunk_off, bucket_begin, nb0, PART_CHUNK, b, lo, hi);
            sm.bucket = b; sm.lo = lo; sm.hi = hi;
        }
    }
    for (int b = tid; b < MAXB; b += PT) sm.cnt[b] = 0;
    __syncthreads();
    const int lo = sm.lo, hi = sm.hi;
    if (lo >= hi) return;                       // CTA-uniform
    const int cbase = FIRST ? 0 : (sm.bucket << lo_bits);

    float4 v[PART_PER_THREAD];
    int d[PART_PER_THREAD], r[PART_PER_THREAD];
    // dense (np,3) input: a thread takes 2 x 4 consecutive particles as 3 aligned float4 each (lo is a multiple of the chunk)
    const float *rawbase = FIRST ? pos + first * ps0 : nullptr;
    const bool vec = FIRST && ps0 == 3 && ps1 == 1 && (((uintptr_t)rawbase) & 15) == 0 && hi - lo == PART_CHUNK;
    auto index_of = [&](int k) { return vec ? lo + ((k >> 2) * PT + tid) * 4 + (k & 3) : lo + k * PT + tid; };
    if (vec) {
        const float4 *p4 = reinterpret_cast<const float4 *>(rawbase);
        float4 a[2], b[2], c[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int64_t q = ((int64_t)lo >> 2) + u * PT + tid;
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
            const int i = lo + k * PT + tid;
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
    const bool few = nbins <= FEW_KEYS;          // CTA-uniform: rank per warp first, one shared atomic per digit present
#pragma unroll
    for (int k = 0; k < PART_PER_THREAD; k++) {
        const int i = index_of(k);
        d[k] = -1;
        if (i < hi) {
            const unsigned t = tile_key<MAS>(v[k].x, v[k].y, v[k].z, inv, tg);
            d[k] = FIRST ? (int)(t >> lo_bits) : (int)(t & ((1u << lo_bits) - 1u));
            if (!few) r[k] = atomicAdd(&sm.cnt[d[k]], 1);
        }
        if (few) {                               // every lane of the warp takes part (hi - lo may cut a warp)
            const int lane = tid & 31;
            const unsigned peers = __match_any_sync(0xffffffffu, d[k]);
            const int leader = __ffs(peers) - 1;
            int base = 0;
            if (d[k] >= 0 && lane == leader) base = atomicAdd(&sm.cnt[d[k]], __popc(peers));
            base = __shfl_sync(0xffffffffu, base, leader);
            r[k] = base + __popc(peers & ((1u << lane) - 1u));
        }
    }
    __syncthreads();
    // exclusive scan of the counters by warp 0 (MAXB / 32 per lane) + one global claim per digit present
    if (tid < 32) {
        constexpr int Q = MAXB / 32;
        int sum = 0;
#pragma unroll 8
        for (int q = 0; q < Q; q++) sum += sm.cnt[tid * Q + q];
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += y; }
        int run = incl - sum;
#pragma unroll 8
        for (int q = 0; q < Q; q++) { sm.start[tid * Q + q] = run; run += sm.cnt[tid * Q + q]; }
    }
    for (int b = PT - 1 - tid; b < MAXB; b += PT) {   // last warps first: warp 0 is scanning
        const int c = (b < nbins) ? sm.cnt[b] : 0;
        if (c) sm.gbase[b] = atomicAdd(&cursor[cbase + b], c);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PART_PER_THREAD; k++) {
        if (d[k] >= 0) {
            const int p = sm.start[d[k]] + r[k];
            sm.stage[p] = v[k];
            sm.dig[p] = (unsigned short)d[k];
        }
    }
    __syncthreads();
    for (int i = tid; i < hi - lo; i += PT) {
        const int dd = sm.dig[i];
        out[sm.gbase[dd] + (i - sm.start[dd])] = sm.stage[i];
    }
}

__global__ void tile_chunks_kernel(const int *__restrict__ tile_begin, int ntiles, int chunk, int *nchunks) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    nchunks[t] = (t < ntiles) ? (tile_begin[t + 1] - tile_begin[t] + chunk - 1) / chunk : 0;
}

// ------------------------------------------------------------------------------------------------
// ACCUMULATE
// ------------------------------------------------------------------------------------------------
constexpr int pad_to(int v, int r) { return v + ((r - v % 32) + 32) % 32; }   // smallest x >= v with x % 32 == r

// Shared-memory tile: SX x SY x SZV cells (tile + S-1 halo cells per axis), z fastest, row pitch PI, plane pitch PL
// (floats).  For deposit_lane_kernel (LANES) the pitches are chosen so that the cells ONE warp instruction updates fall
// into distinct banks:
//   PCS  lane = (a&1, b, c), x-planes a and a+2   bank = 16 a + 4 b + c   PI = 4, PL = 16 (mod 32)
//   TSC  lane = (a, b, c), 27 lanes               bank = 9 a + 3 b + c    PI = 3, PL = 9
template <int MAS, bool LANES = false>
struct TileShape {
    static constexpr int S = Support<MAS>::S;
    static constexpr int SX = TX + S - 1, SY = TY + S - 1, SZV = TZ + S - 1;
    static constexpr int RI = MAS == PYLB_PCS ? 4 : 3, RL = MAS == PYLB_PCS ? 16 : 9;
    static constexpr int PI = LANES ? pad_to(SZV, RI) : ((SZV + 3) & ~3);
    static constexpr int PL = LANES ? pad_to(SY * PI, RL) : SY * PI;
    static constexpr int CELLS = (SX * PL + 3) & ~3;
    static constexpr int ZV = (SZV + 3) / 4;                       // float4 groups per z-row in the flush
    static constexpr int THREADS = 256;
    static constexpr size_t SMEM = sizeof(float) * (size_t)CELLS;
    // deposit_lane_kernel: 16 warps per CTA; per-warp staging of the axis weights of 32 particles, word-major with
    // an even pitch of 34 words: a word row of two consecutive particles is one aligned 8-byte load, and both the
    // staging stores (32 consecutive words) and the per-lane row reads (rows 2 banks apart) stay conflict-free
    static constexpr int LANE_THREADS = 512;
#ifndef PYLB_TSC_LANE_CTAS
#define PYLB_TSC_LANE_CTAS 3
#endif
    static constexpr int LANE_CTAS = MAS == PYLB_PCS ? 2 : PYLB_TSC_LANE_CTAS;      // CTAs per SM (PCS: what fits shared memory)
    static constexpr int SW = 3 * S + 1;                           // wx[S], wy[S], wz[S], W
    static constexpr int SP = 34;
    static constexpr int STAGE_WORDS = SW * SP;
    static constexpr size_t LANE_SMEM = sizeof(float) * ((size_t)CELLS + (size_t)(LANE_THREADS / 32) * STAGE_WORDS);
};

__device__ __forceinline__ void red_add_v4(float *p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// work item b -> (tile, particle range)
struct WorkItem { int tile, lo, hi; };
__device__ __forceinline__ void find_work(int b, const int *__restrict__ chunk_off, const int *__restrict__ tile_begin,
                                          int ntiles, WorkItem &w) {
    int lo = 0, hi = ntiles;                    // invariant: chunk_off[lo] <= b < chunk_off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (chunk_off[mid] <= b) lo = mid; else hi = mid;
    }
    w.tile = lo;
    w.lo = tile_begin[lo] + (b - chunk_off[lo]) * CHUNK;
    w.hi = min(w.lo + CHUNK, tile_begin[lo + 1]);
}

// Flush and clear: tile-local (x,y,z) -> global ((ox+x)%dims, (oy+y)%dims, (oz+z)%dims), added with red.global (halo
// cells overlap the neighbouring tiles, and `number` is accumulate-in-place anyway); 16-byte vector reds when
// dims % 4 == 0.  The tile is left zeroed for the CTA's next work item.
template <class TS, int THREADS>
__device__ __forceinline__ void flush_tile(float *tile, float *__restrict__ grid, const TileGeom &tg, int ox, int oy, int oz) {
    const int dims = tg.dims;
    if ((dims & 3) == 0) {
        for (int i = threadIdx.x; i < TS::SX * TS::SY * TS::ZV; i += THREADS) {
            const int zv = i % TS::ZV, y = (i / TS::ZV) % TS::SY, x = i / (TS::ZV * TS::SY);
            float *row = tile + x * TS::PL + y * TS::PI + zv * 4;
            float4 v;
            if constexpr (TS::PI % 4 == 0 && TS::PL % 4 == 0) {
                v = *reinterpret_cast<float4 *>(row);              // padding words of a row are never written: zero
                if (v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) continue;
                *reinterpret_cast<float4 *>(row) = make_float4(0.f, 0.f, 0.f, 0.f);
            } else {                                               // rows of odd pitch: scalar loads, masked at the row's end
                const bool h1 = zv * 4 + 1 < TS::SZV, h2 = zv * 4 + 2 < TS::SZV, h3 = zv * 4 + 3 < TS::SZV;
                v.x = row[0]; v.y = h1 ? row[1] : 0.f; v.z = h2 ? row[2] : 0.f; v.w = h3 ? row[3] : 0.f;
                if (v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) continue;
                row[0] = 0.f;
                if (h1) row[1] = 0.f;
                if (h2) row[2] = 0.f;
                if (h3) row[3] = 0.f;
            }
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
            float *cell = tile + x * TS::PL + y * TS::PI + z;
            const float v = *cell;
            if (v == 0.f) continue;
            *cell = 0.f;
            int gx = ox + x;
            if (tg.xext == dims) gx %= dims;
            else if (gx >= tg.xext) continue;
            atomicAdd(grid + ((int64_t)gx * dims + (oy + y) % dims) * dims + (oz + z) % dims, v);
        }
    }
}

// tile-local word index of the particle's lowest touched grid point, or -1 (particle routed to the wrong x window: dropped)
template <int MAS, bool LANES = false>
__device__ __forceinline__ int base_cell(const float4 q, float inv, const TileGeom &tg, int ox, int oy, int oz,
                                         float (&C)[3][Support<MAS>::S]) {
    using TS = TileShape<MAS, LANES>;
    const int lx = wrap(axis_stencil<MAS>(q.x, inv, C[0]) - tg.x0, tg.dims) - ox;
    if (lx < 0 || lx >= TX) return -1;
    const int ly = wrap(axis_stencil<MAS>(q.y, inv, C[1]), tg.dims) - oy;
    const int lz = wrap(axis_stencil<MAS>(q.z, inv, C[2]), tg.dims) - oz;
    return lx * TS::PL + ly * TS::PI + lz;
}

// ---- lane per particle: every lane streams its own particle and issues S^3 shared atomicAdd(float) ------------------
// The 32 cells of one warp instruction fall into random banks (and, for clustered input, onto equal addresses):
// ~7 shared-memory wavefronts per warp-wide update, which is what bounds this kernel (93 % of the LSU pipe, ncu).
template <int MAS, bool HASW>
__global__ void __launch_bounds__(256)
deposit_tile_kernel(const float4 *__restrict__ sorted, float inv, TileGeom tg, const int *__restrict__ tile_begin,
                    const int *__restrict__ chunk_off, float *__restrict__ grid) {
    using TS = TileShape<MAS>;
    constexpr int S = TS::S, THREADS = TS::THREADS;
    extern __shared__ __align__(16) float tile[];
    __shared__ WorkItem s_w;
    for (int i = threadIdx.x; i < TS::CELLS / 4; i += THREADS) reinterpret_cast<float4 *>(tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int nitems = chunk_off[tg.ntiles];
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        __syncthreads();                         // the tile is clear, s_w is free
        if (threadIdx.x == 0) find_work(item, chunk_off, tile_begin, tg.ntiles, s_w);
        __syncthreads();
        const int t = s_w.tile, lo = s_w.lo, hi = s_w.hi;
        const int tz = t % tg.ntz, ty = (t / tg.ntz) % tg.nty, tx = t / (tg.ntz * tg.nty);
        const int ox = tx * TX, oy = ty * TY, oz = tz * TZ;
        for (int i = lo + threadIdx.x; i < hi; i += THREADS) {
            const float4 q = __ldg(sorted + i);
            float C[3][S];
            const int cell0 = base_cell<MAS>(q, inv, tg, ox, oy, oz, C);
            if (cell0 < 0) continue;
#pragma unroll
            for (int l = 0; l < S; l++)
#pragma unroll
                for (int m = 0; m < S; m++) {
                    const float cxy = C[0][l] * C[1][m];
#pragma unroll
                    for (int k = 0; k < S; k++) {
                        float v = cxy * C[2][k];
                        if (HASW) v *= q.w;
                        atomicAdd(tile + cell0 + l * TS::PL + m * TS::PI + k, v);
                    }
                }
        }
        __syncthreads();
        flush_tile<TS, THREADS>(tile, grid, tg, ox, oy, oz);
    }
}

// ---- stencil lanes: TSC and PCS --------------------------------------------------------------------------------------
// Per batch of 32 particles, lane p evaluates particle p's base cell and its S weights per axis exactly like the reference
// (deposit.cuh) and stages them in a per-warp scratch (word-major, even pitch 34: the staging stores are conflict-free and
// the rows of two consecutive particles are one aligned 8-byte load).  Then the warp walks the batch: every lane forms
// ((wx[a] * wy[b]) * wz[c]) * W -- the reference's left-to-right fp32 product (MAS_library.pyx:400-404, 493-497) -- for its
// own stencil point(s): PCS lane (a&1, b, c) serves the x-planes a and a+2 of one particle, TSC lane (a, b, c) (27 of 32
// lanes) one point of two consecutive particles.  Either way a lane has two independent cells per step and adds both with
// an optimistic pair of compare-and-swaps (PCS: the pairs of two particles in flight together); since one instruction never
// touches a cell twice, a CAS can only fail when another warp updated the same cell in between (or when two particles of a
// group share the cell), and is then repaired.  PCS 5.95 ms at 512^3 against 11.2 ms for the lane-per-particle kernel,
// TSC 4.15 against 4.94 ms (profiles/r2_deposit_ab_v2.txt); ncu: 0.9 of the float-update rate of shared memory.
template <int MAS, bool HASW>
__global__ void __launch_bounds__(TileShape<MAS, true>::LANE_THREADS, TileShape<MAS, true>::LANE_CTAS)
deposit_lane_kernel(const float4 *__restrict__ sorted, float inv, TileGeom tg, const int *__restrict__ tile_begin,
                    const int *__restrict__ chunk_off, float *__restrict__ grid) {
    static_assert(MAS == PYLB_PCS || MAS == PYLB_TSC, "stencil lanes: TSC and PCS");
    using TS = TileShape<MAS, true>;
    constexpr int S = TS::S, PB = 32, THREADS = TS::LANE_THREADS, NW = THREADS / 32, SP = TS::SP;
    constexpr bool PCS = MAS == PYLB_PCS;
    extern __shared__ __align__(16) float tile[];
    __shared__ WorkItem s_w;
    for (int i = threadIdx.x; i < TS::CELLS / 4; i += THREADS) reinterpret_cast<float4 *>(tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool active = PCS || lane < 27;
    const int la = PCS ? lane >> 4 : (active ? lane / 9 : 0);
    const int lb = PCS ? (lane >> 2) & 3 : (active ? (lane / 3) % 3 : 0);
    const int lc = PCS ? lane & 3 : (active ? lane % 3 : 0);
    unsigned *const cell = reinterpret_cast<unsigned *>(tile) + la * TS::PL + lb * TS::PI + lc;
    float *stage = tile + TS::CELLS + warp * TS::STAGE_WORDS;
    const float *sx0 = stage + la * SP, *sx1 = stage + (PCS ? la + 2 : la) * SP;
    const float *sy = stage + (S + lb) * SP, *sz = stage + (2 * S + lc) * SP, *sw = stage + 3 * S * SP;
    // the two cells of one step: (word offset, value) twice
    struct Step { int b0, b1; float v0, v1; };
    // the steps of the particle pair (pp, pp + 1), pp even: every staged row is read as one 8-byte pair.
    // PCS: two steps (one particle each: planes a and a+2).  TSC: one step (the two particles).
    constexpr int NSTEP = PCS ? 2 : 1;
    auto fetch2 = [&](int pp, int cell0, Step *q) {
        const float2 x0 = *reinterpret_cast<const float2 *>(sx0 + pp), y = *reinterpret_cast<const float2 *>(sy + pp),
                     z = *reinterpret_cast<const float2 *>(sz + pp);
        const int c0 = __shfl_sync(full, cell0, pp), c1 = __shfl_sync(full, cell0, pp + 1);
        if constexpr (PCS) {
            const float2 x1 = *reinterpret_cast<const float2 *>(sx1 + pp);
            q[0].b0 = c0; q[0].b1 = c0 + 2 * TS::PL;
            q[1].b0 = c1; q[1].b1 = c1 + 2 * TS::PL;
            q[0].v0 = (x0.x * y.x) * z.x; q[0].v1 = (x1.x * y.x) * z.x;
            q[1].v0 = (x0.y * y.y) * z.y; q[1].v1 = (x1.y * y.y) * z.y;
            if (HASW) {
                const float2 w = *reinterpret_cast<const float2 *>(sw + pp);
                q[0].v0 *= w.x; q[0].v1 *= w.x; q[1].v0 *= w.y; q[1].v1 *= w.y;
            }
        } else {
            q[0].b0 = c0; q[0].b1 = c1;
            q[0].v0 = (x0.x * y.x) * z.x; q[0].v1 = (x0.y * y.y) * z.y;
            if (HASW) {
                const float2 w = *reinterpret_cast<const float2 *>(sw + pp);
                q[0].v0 *= w.x; q[0].v1 *= w.y;
            }
        }
    };
    // NQ steps at a time: all 2 NQ loads, then all 2 NQ compare-and-swaps are in flight together before the first result
    // is looked at (the float CAS rate on shared memory is latency-bound: it doubles from 16 to 32 warps per SM,
    // profiles/r2_atoms_pattern.txt).  Two steps of a group may meet in a cell (equal particles): the later CAS then fails
    // on its stale value and is repaired like a collision with another warp.
    constexpr int NQ = PCS ? PYLB_LANE_DEPTH : 1, GP = NQ * (PCS ? 1 : 2);   // steps / particles per group (TSC: deeper groups spill at 3 CTAs per SM and measured slower)
    static_assert(PB % GP == 0 && GP % 2 == 0, "groups are whole particle pairs");
    auto fetch_group = [&](int pp, int cell0, Step (&q)[NQ]) {
#pragma unroll
        for (int u = 0; u < GP; u += 2) fetch2(pp + u, cell0, &q[(u / 2) * NSTEP]);
    };
    auto apply_group = [&](const Step (&q)[NQ]) {
        unsigned *p0[NQ], *p1[NQ], o0[NQ], o1[NQ], r0[NQ], r1[NQ];
#pragma unroll
        for (int u = 0; u < NQ; u++) {
            p0[u] = cell + q[u].b0;
            p1[u] = PCS ? p0[u] + 2 * TS::PL : cell + q[u].b1;
            o0[u] = *reinterpret_cast<volatile unsigned *>(p0[u]);
            o1[u] = *reinterpret_cast<volatile unsigned *>(p1[u]);
        }
        bool bad = false;
#pragma unroll
        for (int u = 0; u < NQ; u++) {
            r0[u] = atomicCAS(p0[u], o0[u], __float_as_uint(__uint_as_float(o0[u]) + q[u].v0));
            // TSC: the two particles of a step may share their base cell; the second CAS then sees a stale value
            r1[u] = atomicCAS(p1[u], o1[u], __float_as_uint(__uint_as_float(o1[u]) + q[u].v1));
        }
#pragma unroll
        for (int u = 0; u < NQ; u++) bad = bad || r0[u] != o0[u] || r1[u] != o1[u];
        // ~7 % of the CAS lose against another warp of the CTA (ncu: 0.29 repairs per particle).  One branch for the group;
        // a failed CAS has returned the cell's current value, so the repair is a second optimistic CAS on that value
        // (2 shared-memory wavefronts) and only its rare failure takes the spin loop of atomicAdd (8 wavefronts measured).
        if (bad) {
#pragma unroll
            for (int u = 0; u < NQ; u++) {
                if (r0[u] != o0[u] && atomicCAS(p0[u], r0[u], __float_as_uint(__uint_as_float(r0[u]) + q[u].v0)) != r0[u])
                    atomicAdd(reinterpret_cast<float *>(p0[u]), q[u].v0);
                if (r1[u] != o1[u] && atomicCAS(p1[u], r1[u], __float_as_uint(__uint_as_float(r1[u]) + q[u].v1)) != r1[u])
                    atomicAdd(reinterpret_cast<float *>(p1[u]), q[u].v1);
            }
        }
    };
    auto single = [&](int pp, int cell0) {                // one particle (TSC tail / partial batches)
        const int b = __shfl_sync(full, cell0, pp);
        if (!active) return;
        float v = (sx0[pp] * sy[pp]) * sz[pp];
        if (HASW) v *= sw[pp];
        atomicAdd(reinterpret_cast<float *>(cell + b), v);
        if (PCS) {
            float v1 = (sx1[pp] * sy[pp]) * sz[pp];
            if (HASW) v1 *= sw[pp];
            atomicAdd(reinterpret_cast<float *>(cell + b + 2 * TS::PL), v1);
        }
    };
    const int nitems = chunk_off[tg.ntiles];
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        __syncthreads();                         // the tile is clear, s_w is free
        if (threadIdx.x == 0) find_work(item, chunk_off, tile_begin, tg.ntiles, s_w);
        __syncthreads();
        const int t = s_w.tile, ilo = s_w.lo, ihi = s_w.hi;
        const int tz = t % tg.ntz, ty = (t / tg.ntz) % tg.nty, tx = t / (tg.ntz * tg.nty);
        const int ox = tx * TX, oy = ty * TY, oz = tz * TZ;
        float4 nxt = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ilo + warp * PB + lane < ihi) nxt = __ldg(sorted + ilo + warp * PB + lane);
        for (int i0 = ilo + warp * PB; i0 < ihi; i0 += NW * PB) {      // warp-uniform trip count
            const float4 p = nxt;
            const bool have = i0 + lane < ihi;
            if (i0 + NW * PB + lane < ihi) nxt = __ldg(sorted + i0 + NW * PB + lane);   // next batch in flight
            int cell0 = -1;
            if (have) {
                float C[3][S];
                cell0 = base_cell<MAS, true>(p, inv, tg, ox, oy, oz, C);
                if (cell0 >= 0) {
#pragma unroll
                    for (int a = 0; a < 3; a++)
#pragma unroll
                        for (int k = 0; k < S; k++) stage[(a * S + k) * SP + lane] = C[a][k];
                    if (HASW) stage[3 * S * SP + lane] = p.w;
                }
            }
            __syncwarp();
            unsigned todo = __ballot_sync(full, cell0 >= 0);
            if (todo == full) {
                // the common case, unrolled: the staging offsets become immediates; the staged values of the next step
                // are read before this step's atomics are issued (the compiler will not move a shared-memory load above
                // an atomic by itself)
                // (fully unrolled on purpose: unrolling by 8 steps only -- a loop body that fits the instruction cache -- was
                // measured slower, 8.5 against 7.9 ms at 512^3 PCS)
                Step nq[NQ];
                fetch_group(0, cell0, nq);
#pragma unroll
                for (int pp = 0; pp < PB; pp += GP) {
                    Step q[NQ];
#pragma unroll
                    for (int u = 0; u < NQ; u++) q[u] = nq[u];
                    if (pp + GP < PB) fetch_group(pp + GP, cell0, nq);
                    if (active) apply_group(q);
                }
            } else {
                while (todo) {
                    const int pp = __ffs(todo) - 1;
                    todo &= todo - 1;
                    single(pp, cell0);
                }
            }
            __syncwarp();                                            // the next batch overwrites the staging
        }
        __syncthreads();
        flush_tile<TS, THREADS>(tile, grid, tg, ox, oy, oz);
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

template <class K>
static int set_smem(K kernel, size_t bytes) {
    PYLB_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

// digits of the tile id for a grid of `ntiles` tiles
struct SortPlan {
    bool deep;          // per-tile counts need the second histogram sweep
    int lo_bits, nb0;   // lo digit = low lo_bits bits (<= 10), nb0 = number of hi digits (<= 1024)
};
static SortPlan sort_plan(int ntiles, int force) {
    SortPlan p;
    p.deep = force >= 2 || ntiles > BIN_MAX_KEYS;
    if (ntiles <= BIN_MAX_KEYS) {
        p.lo_bits = (bits_for((unsigned)(ntiles - 1)) + 1) / 2;
        if (p.lo_bits > 8) p.lo_bits = 8;
    } else {
        p.lo_bits = ntiles <= (1 << 18) ? 8 : 10;
    }
    // test hooks: the digit widths of the largest grids on a small one (3: 1024 lo digits; 4: 4 lo digits, many buckets)
    if (force == 3) p.lo_bits = 10;
    if (force == 4 && ntiles <= 4096) p.lo_bits = 2;
    p.nb0 = (ntiles + (1 << p.lo_bits) - 1) >> p.lo_bits;
    return p;
}

// ------------------------------------------------------------------------------------------------
// partition particles by owning x-slab (multi-GPU particle exchange): the sort's pass 0 with a slab key.
// out[offsets[g] .. offsets[g+1]) = (x,y,z,w) of the particles whose lowest touched x-plane lies in slab g.
// ------------------------------------------------------------------------------------------------
template <int MAS, bool HASW>
static int partition_run(const float *pos, int64_t np, int64_t ps0, int64_t ps1, const float *w, int64_t wst, int dims,
                         float inv, int G, float4 *out, int *offsets, cudaStream_t st) {
    keep_pool_memory();
    PYLB_REQUIRE(G <= 256, "pylb_partition_xslab: at most 256 slabs");
    TileGeom tg = tile_geom(dims);
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
    bin_hist_kernel<MAS><<<sm_count(), BIN_THREADS, hist_smem, st>>>(pos, 0, n, ps0, ps1, inv, tg, 0, G, counts);
    PYLB_LAUNCH_CHECK();
    PYLB_CHECK(cub::DeviceScan::ExclusiveSum(tmp, tb, counts, offsets, G + 1, st));
    count_launch(2);
    PYLB_CHECK(cudaMemcpyAsync(cursor, offsets, sizeof(int) * (G + 1), cudaMemcpyDeviceToDevice, st));
    constexpr int PT = 256, CH = PT * PART_PER_THREAD;
    const size_t psm = sizeof(PartSmem<PT, 256>);
    if (set_smem(bin_pass_kernel<MAS, HASW, true, PT, 256>, psm)) return 1;
    bin_pass_kernel<MAS, HASW, true, PT, 256><<<(unsigned)((n + CH - 1) / CH), PT, psm, st>>>(
        pos, w, wst, 0, n, ps0, ps1, inv, tg, nullptr, out, cursor, nullptr, nullptr, 0, G);
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

// tests / A-B runs (pylb_ma_debug_path): units digit 0 automatic, 2 force the deep sort, 3 / 4 deep with 1024 / 4 lo digits;
// hundreds digit 0 automatic, 1 lane-per-particle tile kernel for every scheme, 2 stencil-lane kernel where it exists (TSC, PCS)
static int g_force_sort = 0, g_force_kernel = 0;
void ma_tiled_force_path(int p) {
    if (p < 0) p = 0;
    g_force_kernel = (p / 100) % 10;
    g_force_sort = p % 100;
}
static int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

struct TiledWs {
    int *H, *S, *bucket_begin, *bcursor, *bchunk_off, *tile_begin, *nchunks, *chunk_off;
    float4 *sorted, *sorted_tmp;
    void *tmp;
    size_t tmp_bytes, total;
};

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static void plan_ws(int64_t np, int dims, int xext, TiledWs *ws, char *base) {
    const int ntiles = tile_geom(dims, 0, xext).ntiles;
    const int64_t BATCH = round_particles(dims, xext);
    const int64_t nb = np < BATCH ? np : BATCH;
    size_t o = 0, t2 = 0;
    auto take = [&](size_t bytes) { char *p = base ? base + o : nullptr; o += align_up(bytes); return p; };
    memset(ws, 0, sizeof(*ws));
    const size_t nt = (size_t)ntiles + 1026;
    ws->H = (int *)take(sizeof(int) * nt);            // per-key counts
    ws->S = (int *)take(sizeof(int) * nt);            // per-tile write cursors
    ws->tile_begin = (int *)take(sizeof(int) * nt);
    ws->nchunks = (int *)take(sizeof(int) * nt);
    ws->chunk_off = (int *)take(sizeof(int) * nt);
    ws->bucket_begin = (int *)take(sizeof(int) * 1026);
    ws->bcursor = (int *)take(sizeof(int) * 1026);
    ws->bchunk_off = (int *)take(sizeof(int) * 1026);
    ws->sorted = (float4 *)take(sizeof(float4) * nb);
    ws->sorted_tmp = (float4 *)take(sizeof(float4) * nb);
    cub::DeviceScan::ExclusiveSum(nullptr, t2, (int *)nullptr, (int *)nullptr, ntiles + 1);
    ws->tmp_bytes = t2;
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

bool ma_tiled_supported(int ndim, int dims, int grid_f64, int xext) {
    return ndim == 3 && !grid_f64 && dims >= 32 && tile_geom(dims, 0, xext).ntiles <= MAX_TILES;
}

// PT0 / PT1: CTA sizes of pass 0 and of pass 1 (and the second histogram sweep, which walks the same chunk list).  More
// than 256 digits need the 512-thread CTA (4096-particle chunks, 2 CTAs per SM); a pass with <= 256 digits runs faster with
// 256 threads and 4 CTAs per SM (4.4 against 3.7 TB/s of traffic measured), also in the deep mode.
template <int MAS, bool HASW, int PT0, int PT1, int MAXB0, int MAXB1>
static int run_passes(const float *pos, const float *w, int64_t wst, int64_t first, int n, int64_t ps0, int64_t ps1, float inv,
                      const TileGeom &tg, TiledWs &ws, const SortPlan &sp, cudaStream_t st) {
    constexpr int CH0 = PT0 * PART_PER_THREAD, CH1 = PT1 * PART_PER_THREAD;
    const int nt1 = tg.ntiles + 1;
    size_t tb = ws.tmp_bytes;
    // bucket table, pass-0 cursors and the chunk list of pass 1 from the (tile or bucket) scan in ws.tile_begin
    if (!sp.deep)
        part_buckets_kernel<<<1, 1024, 0, st>>>(ws.tile_begin, tg.ntiles, sp.lo_bits, sp.nb0, CH1, ws.bucket_begin, ws.bcursor, ws.bchunk_off);
    else
        part_buckets_kernel<<<1, 1024, 0, st>>>(ws.tile_begin, sp.nb0, 0, sp.nb0, CH1, ws.bucket_begin, ws.bcursor, ws.bchunk_off);
    PYLB_LAUNCH_CHECK();
    const size_t psm0 = sizeof(PartSmem<PT0, MAXB0>), psm1 = sizeof(PartSmem<PT1, MAXB1>);
    if (set_smem(bin_pass_kernel<MAS, HASW, true, PT0, MAXB0>, psm0) || set_smem(bin_pass_kernel<MAS, HASW, false, PT1, MAXB1>, psm1)) return 1;
    const unsigned g0 = (unsigned)((n + CH0 - 1) / CH0), g1 = (unsigned)((n + CH1 - 1) / CH1) + (unsigned)sp.nb0;
    bin_pass_kernel<MAS, HASW, true, PT0, MAXB0><<<g0, PT0, psm0, st>>>(
        pos, w, wst, first, n, ps0, ps1, inv, tg, nullptr, ws.sorted_tmp, ws.bcursor, ws.bucket_begin, ws.bchunk_off, sp.lo_bits, sp.nb0);
    PYLB_LAUNCH_CHECK();
    if (sp.deep) {
        // per-tile counts from the bucket-ordered payload, then every tile's first slot
        PYLB_CHECK(cudaMemsetAsync(ws.H, 0, sizeof(int) * (size_t)nt1, st));
        bin_hist2_kernel<MAS, PT1><<<g1, PT1, 0, st>>>(ws.sorted_tmp, inv, tg, ws.bucket_begin, ws.bchunk_off, sp.lo_bits, sp.nb0, ws.H);
        PYLB_LAUNCH_CHECK();
        PYLB_CHECK(cub::DeviceScan::ExclusiveSum(ws.tmp, tb, ws.H, ws.tile_begin, nt1, st));
        count_launch(2);
    }
    PYLB_CHECK(cudaMemcpyAsync(ws.S, ws.tile_begin, sizeof(int) * (size_t)nt1, cudaMemcpyDeviceToDevice, st));
    bin_pass_kernel<MAS, HASW, false, PT1, MAXB1><<<g1, PT1, psm1, st>>>(
        pos, w, wst, first, n, ps0, ps1, inv, tg, ws.sorted_tmp, ws.sorted, ws.S, ws.bucket_begin, ws.bchunk_off, sp.lo_bits, sp.nb0);
    PYLB_LAUNCH_CHECK();
    return 0;
}

template <int MAS, bool HASW>
static int tiled_run(const float *pos, int64_t np, int64_t ps0, int64_t ps1, float *grid, int dims, float inv,
                     const float *w, int64_t wst, int x0, int xext, TiledWs &ws, cudaStream_t st) {
    using TS = TileShape<MAS>;
    const TileGeom tg = tile_geom(dims, x0, xext);
    const SortPlan sp = sort_plan(tg.ntiles, g_force_sort);
    const int P = sm_count();
    // always the maximum the kernels may ever need: the attribute is a limit, and a smaller value set here would make
    // a later, larger launch of the same instantiation fail
    if (set_smem(deposit_tile_kernel<MAS, HASW>, TS::SMEM)) return 1;
    using TL = TileShape<MAS, true>;
    constexpr bool HAS_LANES = MAS == PYLB_PCS || MAS == PYLB_TSC;
    if constexpr (HAS_LANES) {
        if (set_smem(deposit_lane_kernel<MAS, HASW>, TL::LANE_SMEM)) return 1;
    }
    if (set_smem(bin_hist_kernel<MAS>, sizeof(int) * (size_t)BIN_MAX_KEYS)) return 1;
    const int kern = g_force_kernel > 0 ? g_force_kernel : env_int("PYLB_TILE_KERNEL", 0);
    const int nt1 = tg.ntiles + 1;
    const int64_t BATCH = round_particles(dims, xext);
    for (int64_t first = 0; first < np; first += BATCH) {
        const int n = (int)((np - first) < BATCH ? (np - first) : BATCH);
        size_t tb = ws.tmp_bytes;
        timing_begin(PYLB_T_SORT, st);
        // histogram of the tiles (shallow) or of their hi digits (deep), and its exclusive scan (counts[nkeys] = 0)
        const int shift = sp.deep ? sp.lo_bits : 0, nkeys = sp.deep ? sp.nb0 : tg.ntiles;
        PYLB_CHECK(cudaMemsetAsync(ws.H, 0, sizeof(int) * (size_t)(nkeys + 1), st));
        bin_hist_kernel<MAS><<<P, BIN_THREADS, sizeof(int) * (size_t)nkeys, st>>>(pos, first, n, ps0, ps1, inv, tg, shift, nkeys, ws.H);
        PYLB_LAUNCH_CHECK();
        PYLB_CHECK(cub::DeviceScan::ExclusiveSum(ws.tmp, tb, ws.H, ws.tile_begin, nkeys + 1, st));
        count_launch(2);
        int rc;
        const bool wide0 = sp.nb0 > 256, wide1 = (1 << sp.lo_bits) > 256;
        if (!wide0 && !wide1) rc = run_passes<MAS, HASW, 256, 256, 256, 256>(pos, w, wst, first, n, ps0, ps1, inv, tg, ws, sp, st);
        else if (!wide1) rc = run_passes<MAS, HASW, PYLB_WIDE_PT0, 256, 1024, 256>(pos, w, wst, first, n, ps0, ps1, inv, tg, ws, sp, st);
        else rc = run_passes<MAS, HASW, 512, 512, 1024, 1024>(pos, w, wst, first, n, ps0, ps1, inv, tg, ws, sp, st);
        if (rc) return rc;
        tile_chunks_kernel<<<(nt1 + 255) / 256, 256, 0, st>>>(ws.tile_begin, tg.ntiles, CHUNK, ws.nchunks);
        PYLB_LAUNCH_CHECK();
        tb = ws.tmp_bytes;
        PYLB_CHECK(cub::DeviceScan::ExclusiveSum(ws.tmp, tb, ws.nchunks, ws.chunk_off, nt1, st));
        count_launch(2);
        timing_end(PYLB_T_SORT, st);
        // persistent CTAs walk the work items (tile, chunk) round-robin; every non-empty tile has at most n/CHUNK + 1 chunks
        const int64_t max_items = (int64_t)n / CHUNK + tg.ntiles;
        timing_begin(PYLB_T_TILE, st);
        // TSC, PCS: stencil lanes with optimistic CAS pairs unless the lane-per-particle kernel is forced (A/B, tests)
        bool lanes = false;
        if constexpr (HAS_LANES) lanes = kern != 1;
        if (lanes) {
            if constexpr (HAS_LANES) {
                const int64_t g = (int64_t)P * TL::LANE_CTAS * 8;
                deposit_lane_kernel<MAS, HASW><<<(unsigned)(max_items < g ? max_items : g), TL::LANE_THREADS, TL::LANE_SMEM, st>>>(
                    ws.sorted, inv, tg, ws.tile_begin, ws.chunk_off, grid);
            }
        } else {
            const int64_t g = (int64_t)P * 4 * 8;
            deposit_tile_kernel<MAS, HASW><<<(unsigned)(max_items < g ? max_items : g), TS::THREADS, TS::SMEM, st>>>(
                ws.sorted, inv, tg, ws.tile_begin, ws.chunk_off, grid);
        }
        PYLB_LAUNCH_CHECK();
        timing_end(PYLB_T_TILE, st);
    }
    return 0;
}

int ma_tiled(const float *pos, int64_t np, int64_t ps0, int64_t ps1, float *grid, int dims, float inv, int mas,
             const float *w, int64_t wst, int x0, int xext, void *workspace, size_t workspace_bytes, cudaStream_t st) {
    if (np == 0) return 0;
    TiledWs ws;
    plan_ws(np, dims, xext, &ws, (char *)workspace);
    PYLB_REQUIRE(workspace != nullptr && workspace_bytes >= ws.total, "pylb_ma: tiled workspace too small (%zu < %zu)",
                 workspace_bytes, ws.total);
    PYLB_REQUIRE(((uintptr_t)grid & 15) == 0, "pylb_ma: grid must be 16-byte aligned");
    PYLB_REQUIRE(tile_geom(dims, x0, xext).ntiles <= MAX_TILES, "pylb_ma: more than 2^20 tiles in one window; deposit onto x-windows");
    const bool hw = w != nullptr;
    switch (mas) {
        case PYLB_NGP: return hw ? tiled_run<PYLB_NGP, true>(pos, np, ps0, ps1, grid, dims, inv, w, wst, x0, xext, ws, st)
                                 : tiled_run<PYLB_NGP, false>(pos, np, ps0, ps1, grid, dims, inv, w, wst, x0, xext, ws, st);
        case PYLB_CIC: return hw ? tiled_run<PYLB_CIC, true>(pos, np, ps0, ps1, grid, dims, inv, w, wst, x0, xext, ws, st)
                                 : tiled_run<PYLB_CIC, false>(pos, np, ps0, ps1, grid, dims, inv, w, wst, x0, xext, ws, st);
        case PYLB_TSC: return hw ? tiled_run<PYLB_TSC, true>(pos, np, ps0, ps1, grid, dims, inv, w, wst, x0, xext, ws, st)
                                 : tiled_run<PYLB_TSC, false>(pos, np, ps0, ps1, grid, dims, inv, w, wst, x0, xext, ws, st);
        case PYLB_PCS: return hw ? tiled_run<PYLB_PCS, true>(pos, np, ps0, ps1, grid, dims, inv, w, wst, x0, xext, ws, st)
                                 : tiled_run<PYLB_PCS, false>(pos, np, ps0, ps1, grid, dims, inv, w, wst, x0, xext, ws, st);
    }
    set_error("pylb_ma: unknown mass-assignment scheme %d", mas);
    return 1;
}

}  // namespace pylb
