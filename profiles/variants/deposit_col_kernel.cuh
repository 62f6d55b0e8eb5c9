// NOT COMPILED -- lab notes of round 2.  "Column accumulators in registers": the tile kernel for TSC / PCS that replaces
// the per-stencil-point shared-memory update by S^2 register FMAs per particle (lanes = 32 consecutive z cells of the (x,y)
// column being visited) and adds the accumulators to the shared tile once per column visit.  It was wired into
// deposit_tiled.cu as tile kernel 3 (pylb_ma_debug_path 300), passed every parity test, and lost:
//
//   512^3 particles -> 512^3 grid, B200            PCS       TSC
//   lane per particle (kernel 1)                   11.2 ms   4.95 ms
//   stencil lanes, optimistic CAS (kernel 2)        7.83      4.33      (6.78 / 4.23 after the weight evaluation was re-cut)
//   column accumulators (this file)                10.87      7.82      (11.9 / 8.56 with unconditional flush atomics)
//
// ncu (profiles/r2_ncu_col_kernel.txt): 65.5 warp instructions per particle at IPC 3.0 -- 33 in the walk loop (4 LDS, the
// z-window test, 4 FMUL + 8 FFMA2 and their control flow: 28 of 32 lanes multiply by zero), 16.5 in the per-particle weight
// evaluation (before it was re-cut), 11 in the column flush (16 CAS loops per ~15 particles), 5 in the in-CTA column sort.
// Shared memory is no longer the limit (LSU pipe 28 %), the issue slots are.  A register-accumulating kernel needs lanes
// that all do useful FMAs (stencil points as lanes plus a sliding z window over CELL-sorted particles), which needs a full
// in-tile cell sort; estimated 23-27 instructions per particle against 32 for the stencil-lane kernel -- not built.
//
// Needs TileShape / flush_tile / find_work / TileGeom from pylians_b200/csrc/deposit_tiled.cu.

// ---- column accumulators in registers: TSC and PCS -----------------------------------------------------------------
// Every variant above pays one shared-memory update per particle and stencil point, and shared memory takes 3-6 float
// updates per clock and SM whatever the method.  This kernel pays S^2 register FMAs per particle and warp instead and
// touches the tile once per COLUMN VISIT:
//   * the work item's <= 8192 particles are counting-sorted by (x,y) column inside the CTA (512 keys: 256 columns x
//     "z-stencil crosses the tile's last cell"), as 2-byte indices -- the payload itself is re-read through L1/L2;
//   * the sorted sequence is cut into 16 equal warp shares.  The 32 lanes of a warp are 32 consecutive z cells of the
//     column being visited; lane l holds the S x S accumulators of cells (x0+a, y0+b, z_l).  For a particle of that
//     column lane l adds (wx[a] * wz[z_l - z0]) * wy[b] (zero outside the particle's S cells in z): packed f32x2 FMAs,
//     no shared-memory update.  When the column changes, the accumulators are added to the shared tile -- 32
//     consecutive words per instruction, conflict-free -- and cleared.  At one particle per cell that is S^2
//     warp-wide updates per ~30 particles instead of 2 S^3 / 32 per particle.
//   * weights come from the same per-lane evaluation as the other kernels (deposit.cuh), staged per batch of 32.
// The product is formed as (wx * (wz * W)) * wy and accumulated with FMAs, i.e. it differs from the reference's
// ((wx * wy) * wz) * W by rounding only (1e-7 relative; the grids' contract is 1e-5).
template <int MAS>
struct ColShape {
    using TS = TileShape<MAS>;
    static constexpr int S = TS::S, THREADS = 512, NW = THREADS / 32, NKEY = TX * TY * 2, PB = 32, SP = 16;
    static constexpr int ZHI = TZ - (S - 1);               // lowest cells >= ZHI: the stencil reaches beyond lane 31
    static constexpr int PPT = CHUNK / THREADS;            // particles per thread in the sort phases
    static constexpr size_t SMEM = TS::SMEM + sizeof(float) * NW * PB * SP + sizeof(int) * (NKEY + 8) + sizeof(unsigned short) * CHUNK;
    static_assert(NKEY == THREADS, "one sort key per thread");
    static_assert(TZ == 32, "lanes are z cells");
};

template <int MAS, bool HASW>
__global__ void __launch_bounds__(ColShape<MAS>::THREADS, 2)
deposit_col_kernel(const float4 *__restrict__ sorted, float inv, TileGeom tg, const int *__restrict__ tile_begin,
                   const int *__restrict__ chunk_off, float *__restrict__ grid) {
    static_assert(MAS == PYLB_PCS || MAS == PYLB_TSC, "column kernel: TSC and PCS");
    using CS = ColShape<MAS>;
    using TS = TileShape<MAS>;
    constexpr int S = CS::S, THREADS = CS::THREADS, NW = CS::NW, PB = CS::PB, SP = CS::SP, H = S / 2;
    extern __shared__ __align__(16) float tile[];
    float *const stage = tile + TS::CELLS + (threadIdx.x >> 5) * (PB * SP);
    int *const cnt = reinterpret_cast<int *>(tile + TS::CELLS + NW * PB * SP);
    unsigned short *const idx = reinterpret_cast<unsigned short *>(cnt + CS::NKEY + 8);
    __shared__ WorkItem s_w;
    __shared__ int s_wsum[NW];
    const unsigned full = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < TS::CELLS / 4; i += THREADS) reinterpret_cast<float4 *>(tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    float2 acc2[S][H > 0 ? H : 1];
    float acc1[S];
#pragma unroll
    for (int a = 0; a < S; a++) {
        acc1[a] = 0.f;
#pragma unroll
        for (int h = 0; h < H; h++) acc2[a][h] = make_float2(0.f, 0.f);
    }
    // add the accumulators of column visit `key` to the tile (lane l: z = z_start + l) and clear them.  A visit of the
    // upper group (lowest cells ZHI .. TZ-1) reaches the 2 S - 2 cells ZHI .. TZ + S - 2 only.
    auto flush_acc = [&](int key) {
        const int col = key >> 1;
        float *base = tile + (col / TY) * TS::PL + (col % TY) * TS::PI + ((key & 1) ? CS::ZHI : 0) + lane;
        if (!(key & 1) || lane < 2 * S - 2) {
#pragma unroll
            for (int a = 0; a < S; a++) {
#pragma unroll
                for (int h = 0; h < H; h++) {
                    atomicAdd(base + a * TS::PL + (2 * h) * TS::PI, acc2[a][h].x);
                    atomicAdd(base + a * TS::PL + (2 * h + 1) * TS::PI, acc2[a][h].y);
                }
                if (S & 1) atomicAdd(base + a * TS::PL + (S - 1) * TS::PI, acc1[a]);
            }
        }
#pragma unroll
        for (int a = 0; a < S; a++) {
            acc1[a] = 0.f;
#pragma unroll
            for (int h = 0; h < H; h++) acc2[a][h] = make_float2(0.f, 0.f);
        }
    };

    const int nitems = chunk_off[tg.ntiles];
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        __syncthreads();                         // the tile is clear; s_w, cnt and idx are free
        if (tid == 0) find_work(item, chunk_off, tile_begin, tg.ntiles, s_w);
        cnt[tid] = 0;
        __syncthreads();
        const int t = s_w.tile, ilo = s_w.lo, n = s_w.hi - s_w.lo;
        const int tz = t % tg.ntz, ty = (t / tg.ntz) % tg.nty, tx = t / (tg.ntz * tg.nty);
        const int ox = tx * TX, oy = ty * TY, oz = tz * TZ;
        const float4 *const src = sorted + ilo;

        // (1) column key of every particle and its rank inside the key (native integer shared atomics)
        unsigned kr[CS::PPT];
#pragma unroll
        for (int k0 = 0; k0 < CS::PPT; k0 += 4) {
            float4 q[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                q[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if ((k0 + u) * THREADS + tid < n) q[u] = __ldg(src + (k0 + u) * THREADS + tid);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                kr[k0 + u] = 0xffffffffu;
                if ((k0 + u) * THREADS + tid < n) {
                    float C[S];
                    const int lx = wrap(axis_stencil<MAS>(q[u].x, inv, C) - tg.x0, tg.dims) - ox;
                    if (lx < 0 || lx >= TX) continue;          // routed to the wrong x window: dropped
                    const int ly = wrap(axis_stencil<MAS>(q[u].y, inv, C), tg.dims) - oy;
                    const int lz = wrap(axis_stencil<MAS>(q[u].z, inv, C), tg.dims) - oz;
                    const unsigned key = (unsigned)((lx * TY + ly) * 2 + (lz >= CS::ZHI ? 1 : 0));
                    kr[k0 + u] = (key << 16) | (unsigned)atomicAdd(&cnt[key], 1);
                }
            }
        }
        __syncthreads();
        // (2) exclusive scan of the 512 counters (one per thread)
        {
            const int c = cnt[tid];
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(full, incl, o); if (lane >= o) incl += y; }
            if (lane == 31) s_wsum[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                const int v = lane < NW ? s_wsum[lane] : 0;
                int w = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(full, w, o); if (lane >= o) w += y; }
                if (lane < NW) s_wsum[lane] = w - v;
            }
            __syncthreads();
            const int start = incl - c + s_wsum[warp];
            cnt[tid] = start;
            if (tid == THREADS - 1) cnt[CS::NKEY] = start + c;   // particles kept
        }
        __syncthreads();
        // (3) the permutation: sorted position -> index inside the chunk
#pragma unroll
        for (int k = 0; k < CS::PPT; k++)
            if (kr[k] != 0xffffffffu) idx[cnt[kr[k] >> 16] + (int)(kr[k] & 0xffffu)] = (unsigned short)(k * THREADS + tid);
        const int m = cnt[CS::NKEY];
        __syncthreads();

        // (4) every warp walks its share of the column-sorted sequence
        const int pb = (warp * m) / NW, pe = ((warp + 1) * m) / NW;
        int cur = -1;
        float4 nxt = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pb + lane < pe) nxt = __ldg(src + idx[pb + lane]);
        for (int p0 = pb; p0 < pe; p0 += PB) {
            const float4 p = nxt;
            if (p0 + PB + lane < pe) nxt = __ldg(src + idx[p0 + PB + lane]);   // next batch in flight
            if (p0 + lane < pe) {
                float C[3][4];
#pragma unroll
                for (int a = 0; a < 3; a++) C[a][3] = 0.f;
                float Cx[S], Cy[S], Cz[S];
                const int lx = wrap(axis_stencil<MAS>(p.x, inv, Cx) - tg.x0, tg.dims) - ox;
                const int ly = wrap(axis_stencil<MAS>(p.y, inv, Cy), tg.dims) - oy;
                const int lz = wrap(axis_stencil<MAS>(p.z, inv, Cz), tg.dims) - oz;
#pragma unroll
                for (int k = 0; k < S; k++) { C[0][k] = Cx[k]; C[1][k] = Cy[k]; C[2][k] = HASW ? Cz[k] * p.w : Cz[k]; }
                const int hi = lz >= CS::ZHI ? 1 : 0;
                const int meta = ((lx * TY + ly) * 2 + hi) | ((lz - (hi ? CS::ZHI : 0)) << 16);
                float4 *row = reinterpret_cast<float4 *>(stage + lane * SP);
                row[0] = make_float4(C[0][0], C[0][1], C[0][2], C[0][3]);
                row[1] = make_float4(C[1][0], C[1][1], C[1][2], C[1][3]);
                row[2] = make_float4(C[2][0], C[2][1], C[2][2], C[2][3]);
                stage[lane * SP + 12] = __int_as_float(meta);
            } else {
                stage[lane * SP + 12] = __int_as_float(-1);     // end of the warp's share
            }
            __syncwarp();
            // no software pipelining: 8 warps per scheduler hide the two dependent shared-memory loads of a particle
#pragma unroll 2
            for (int j = 0; j < PB; j++) {
                const float *sj = stage + j * SP;
                const int meta = __float_as_int(sj[12]);
                if (meta < 0) break;               // warp-uniform, like everything that steers this loop
                const int key = meta & 0xffff, zrel = lane - (meta >> 16);
                const float wz = ((unsigned)zrel < (unsigned)S) ? sj[8 + zrel] : 0.f;
                const float4 wx = *reinterpret_cast<const float4 *>(sj), wy = *reinterpret_cast<const float4 *>(sj + 4);
                if (key != cur) {
                    if (cur >= 0) flush_acc(cur);
                    cur = key;
                }
                const float wxa[4] = {wx.x, wx.y, wx.z, wx.w};
                const float wyb[4] = {wy.x, wy.y, wy.z, wy.w};
#pragma unroll
                for (int a = 0; a < S; a++) {
                    const float ta = wxa[a] * wz;
                    const float2 tt = make_float2(ta, ta);
#pragma unroll
                    for (int h = 0; h < H; h++) acc2[a][h] = __ffma2_rn(tt, make_float2(wyb[2 * h], wyb[2 * h + 1]), acc2[a][h]);
                    if (S & 1) acc1[a] = fmaf(ta, wyb[S - 1], acc1[a]);
                }
            }
            __syncwarp();                          // the next batch overwrites the staging
        }
        if (cur >= 0) flush_acc(cur);
        __syncthreads();
        flush_tile<TS, THREADS>(tile, grid, tg, ox, oy, oz);
    }
}

