// Fused MAS-deconvolution + |delta_k|^2 + k-shell / (k_par,k_per) / 1-D binning + Legendre weighting.
//
// Replaces the serial mode loops of class Pk (library/Pk_library/Pk_library.pyx:314-381) and
// class XPk (:628-737).  HBM-bound: every complex mode (8 B per field) is read exactly once.
//
// Two kernels:
//
//  ring_kernel (line of sight = z, the contiguous half-spectrum axis)
//     Rows (kx,ky) are sorted by r2 = kx^2+ky^2.  A thread owns one kz and walks a contiguous span of
//     the sorted row list, so that
//       * all rows with the same r2 share |k|, mu, the Legendre weights and every bin index
//         -> geometry is evaluated once per r2-group, the per-mode work is load + deconvolve + square;
//       * inside a ring p <= sqrt(r2) < p+1 a thread can only hit 3-D bins b0 or b0+1 with
//         b0 = floor(sqrt(p^2+kz^2)), one 2-D bin (p, kz) and one 1-D bin kz
//         -> all accumulation is in REGISTERS; no shared or global atomic in the inner loop
//            (shared fp32/fp64 atomics are CAS loops on sm_100a; only red.global is native).
//     Registers are flushed with red.global.add.f64 when the ring changes (a few times per span).
//     The self-conjugate columns kz=0 and kz=dims/2 (skip rule :326-330) are left to generic_kernel.
//
//  generic_kernel (any axis, any F <= 8, any subset of kz columns)
//     One thread per mode, red.global for everything.  Used for the two special columns above and
//     as the any-axis path.
#include <map>
#include <mutex>

#include <cub/device/device_radix_sort.cuh>

#include <vector>

#include "common.cuh"

namespace pylb {

constexpr int MAX_F = 8;

struct FieldPtrs {
    float2 *p[MAX_F];
};

struct BinGeom {
    int dims, middle, even;
    int x0, nx, y0, ny;
    long long stride_x, stride_y;
    int axis;
    int kmax_par1;  // kmax_par + 1
    int F, X;
    long long o_k3d, o_p3d, o_x3d, o_phase, o_p1d, o_x1d, o_p2d, o_x2d;
    long long o_n3d, o_n1d, o_n2d;
    double *sums;
    uint64_t *counts;
    const double *mas_tab;  // [F][middle+1]: (x/sin x)^p at |k| = 0..middle
    int mas_idx[MAX_F];
    int ximag;              // class XPk_imag: cross term im_i*re_j - re_i*im_j (:1131-1132) instead of re_i*re_j + im_i*im_j
};

// ------------------------------------------------------------------------------------------------
// MAS window table, Pk_library.pyx:86-87 and :316,:320,:324:  (x/sin x)^p, x = pi*k/dims, 1 at k=0.
// x/sin(x) is even, so |k| indexes it.
// ------------------------------------------------------------------------------------------------
__global__ void mas_table_kernel(double *tab, int middle, int dims, int F, BinGeom g) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * (middle + 1)) return;
    const int f = i / (middle + 1), k = i % (middle + 1);
    const int p = g.mas_idx[f];
    double v = 1.0;
    if (k != 0 && p != 0) {
        const double x = (M_PI / (double)dims) * (double)k;
        const double q = x / sin(x);
        v = q;
        if (p == 2) v = q * q;
        else if (p == 3) v = q * q * q;
        else if (p == 4) { const double q2 = q * q; v = q2 * q2; }
    }
    tab[i] = v;
}

// ------------------------------------------------------------------------------------------------
// Row table for the ring kernel
// ------------------------------------------------------------------------------------------------
struct __align__(16) RowEnt {
    int r2;
    short kx, ky;
    long long off;  // BYTE offset of the row start: 8*(ix*stride_x + iy*stride_y)
};

__global__ void row_keys_kernel(unsigned *keys, unsigned *vals, int nrows, BinGeom g) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    const int ix = r / g.ny, iy = r - ix * g.ny;
    const int kx = wavenumber(g.x0 + ix, g.dims, g.middle), ky = wavenumber(g.y0 + iy, g.dims, g.middle);
    keys[r] = (unsigned)(kx * kx + ky * ky);
    vals[r] = (unsigned)r;
}

__global__ void row_table_kernel(const unsigned *keys, const unsigned *vals, RowEnt *tab, int nrows, BinGeom g) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    const int r = (int)vals[i];
    const int ix = r / g.ny, iy = r - ix * g.ny;
    RowEnt e;
    e.r2 = (int)keys[i];
    e.kx = (short)wavenumber(g.x0 + ix, g.dims, g.middle);
    e.ky = (short)wavenumber(g.y0 + iy, g.dims, g.middle);
    e.off = 8ll * ((long long)ix * g.stride_x + (long long)iy * g.stride_y);
    tab[i] = e;
}

// ------------------------------------------------------------------------------------------------
// phase^2 with phase = atan2(re, |delta_k|)  (Pk_library.pyx:361 -- sic, the real part against the
// modulus).  |re|/|delta_k| <= 1 and atan^2 is even, so phase^2 = s*G(s) with s = re^2/|delta_k|^2 in
// [0,1]; G(s) = atan(sqrt(s))^2/s is analytic on [0,1] (nearest singularity s = -1) and a degree-9
// near-minimax polynomial reproduces it to 1e-8 (2e-7 with fp32 Horner rounding).  No branches,
// no range reduction: 1 MUFU.RCP + 11 FMUL/FFMA.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float phase_sq(float re, float d2) {
    float rc;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(fmaxf(d2, 1e-37f)));   // 1 MUFU; 1-ulp error is irrelevant here
    const float s = fminf((re * re) * rc, 1.0f);
    float q = -6.465150895e-03f;
    q = fmaf(q, s, 3.925943169e-02f);
    q = fmaf(q, s, -1.113286174e-01f);
    q = fmaf(q, s, 2.034298861e-01f);
    q = fmaf(q, s, -2.851652190e-01f);
    q = fmaf(q, s, 3.508593260e-01f);
    q = fmaf(q, s, -4.181177634e-01f);
    q = fmaf(q, s, 5.110430921e-01f);
    q = fmaf(q, s, -6.666647075e-01f);
    q = fmaf(q, s, 9.999999906e-01f);
    return q * s;
}

// ---- packed fp32x2 arithmetic (Blackwell FMUL2 / FFMA2): two independent IEEE-rounded lanes per instruction
__device__ __forceinline__ unsigned long long pack2(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float2 unpack2(unsigned long long v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// phase^2 of two modes at once: same polynomial as phase_sq, one FFMA2 per coefficient
__device__ __forceinline__ float2 phase_sq2(float re2a, float d2a, float re2b, float d2b) {
    float ra, rb;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ra) : "f"(fmaxf(d2a, 1e-37f)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rb) : "f"(fmaxf(d2b, 1e-37f)));
    const unsigned long long s = pack2(fminf(re2a * ra, 1.0f), fminf(re2b * rb, 1.0f));
#define PYLB_C2(c) pack2(c, c)
    unsigned long long q = PYLB_C2(-6.465150895e-03f);
    q = fma2(q, s, PYLB_C2(3.925943169e-02f));
    q = fma2(q, s, PYLB_C2(-1.113286174e-01f));
    q = fma2(q, s, PYLB_C2(2.034298861e-01f));
    q = fma2(q, s, PYLB_C2(-2.851652190e-01f));
    q = fma2(q, s, PYLB_C2(3.508593260e-01f));
    q = fma2(q, s, PYLB_C2(-4.181177634e-01f));
    q = fma2(q, s, PYLB_C2(5.110430921e-01f));
    q = fma2(q, s, PYLB_C2(-6.666647075e-01f));
    q = fma2(q, s, PYLB_C2(9.999999906e-01f));
#undef PYLB_C2
    return unpack2(mul2(q, s));
}

__device__ __forceinline__ float2 ld_stream(const float2 *p) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}

// ------------------------------------------------------------------------------------------------
// ring kernel
//
// PRECISE = false (default): |delta_k|^2, the cross products and phase^2 are formed in fp32 from the
//   fp32-deconvolved mode and summed in fp32 over the (<= a few dozen) rows sharing r2, then every
//   further accumulation is fp64.  Per-mode rounding is 6e-8 relative and unbiased -- below the
//   reference's own FFT-to-FFT differences and 100x inside the 1e-5 contract.
// PRECISE = true ("fp64 accumulation option"): every mode is squared and summed in fp64 exactly like
//   Pk_library.pyx:358-360.
// ------------------------------------------------------------------------------------------------
constexpr int RING_T = 256;         // threads per CTA = kz values per CTA
constexpr int RING_SPAN_MAX = 512;  // rows per CTA (their table entries are staged in shared memory once)

template <int F> struct RingCfg { static constexpr int D = (F == 1) ? 16 : 8; };  // cp.async pipeline depth (rows)

template <int F>
struct RingSmem {
    float2 z[RingCfg<F>::D][RING_T][F];   // per-thread prefetch ring: thread t only ever touches z[.][t][.]
    RowEnt ent[RING_SPAN_MAX];
    double cxy[RING_SPAN_MAX][F];
};

// Bulk variant: one elected lane copies whole row segments with cp.async.bulk (the TMA engine's 1-D
// path) into a ring of stages; full/empty mbarriers replace per-thread cp.async groups.  Rows are only
// 8-byte aligned and bulk copies need 16 bytes, so a row's window starts one element early when needed
// (dlt = 0|1) and a segment serves RING_T-1 kz values.
constexpr int BULK_ROWS = 4;     // rows per stage (= rows per loop iteration)
constexpr int BULK_SLOTS = 4;    // stages in the ring (16 rows in flight)

template <int F>
struct RingSmemBulk {
    float2 z[BULK_SLOTS][BULK_ROWS][F][RING_T];   // 2 KB per (row, field): 16-byte aligned windows
    RowEnt ent[RING_SPAN_MAX];
    double cxy[RING_SPAN_MAX][F];
    unsigned long long full[BULK_SLOTS], empty[BULK_SLOTS];
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, void *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <bool PRECISE> struct AccT { typedef float type; };
template <> struct AccT<true> { typedef double type; };

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int F, bool BULK> struct RingSmemSel { typedef RingSmem<F> type; };
template <int F> struct RingSmemSel<F, true> { typedef RingSmemBulk<F> type; };

template <int F, bool PHASE, bool WB, bool PRECISE, bool BULK>
__global__ void __launch_bounds__(BULK ? RING_T + 32 : RING_T, F == 1 ? (PRECISE ? 2 : 3) : (F == 2 ? 2 : 1))
ring_kernel(BinGeom g, FieldPtrs dk, const RowEnt *__restrict__ tab, int nrows, int rows_per_span, int kz_hi) {
    constexpr int X = F * (F - 1) / 2;
    constexpr int Q = F + X;
    constexpr int D = RingCfg<F>::D;
    typedef typename AccT<PRECISE>::type acc_t;
    typedef typename RingSmemSel<F, BULK>::type Smem;
    extern __shared__ __align__(128) unsigned char ring_smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(ring_smem_raw);

    const int tid = threadIdx.x;
    constexpr int LANES = BULK ? RING_T - 1 : RING_T;          // kz values served by one segment
    const int kz_first = 1 + blockIdx.y * LANES;
    const int kz = kz_first + tid;
    const bool active = tid < LANES && kz <= kz_hi;
    const int kzc = active ? kz : kz_hi;  // clamp so idle lanes stay in bounds
    const int kz2 = kzc * kzc;
    const int i0 = blockIdx.x * rows_per_span;
    const int total = min(nrows, i0 + rows_per_span) - i0;
    if (total <= 0) return;

    // stage this span's row table (and the per-row x*y MAS factors) once
    if (BULK && tid == 0) {
#pragma unroll
        for (int q = 0; q < BULK_SLOTS; q++) {
            mbar_init(&reinterpret_cast<RingSmemBulk<F> &>(sm).full[q], 1);            // the producer's expect_tx arrive
            mbar_init(&reinterpret_cast<RingSmemBulk<F> &>(sm).empty[q], RING_T / 32);  // one lane per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int j = tid; j < total; j += (int)blockDim.x) {
        RowEnt e = tab[i0 + j];
        if (BULK) {
            // window start = kz_first - dlt must sit on a 16-byte boundary: (row element offset + start) even
            const int dlt = (int)(((e.off >> 3) + kz_first) & 1);
            e.kx = (short)dlt;   // kx/ky are not needed after the MAS factors below; reuse the slot
        }
        const RowEnt e0 = tab[i0 + j];
        sm.ent[j] = e;
        e = e0;
        const int ax = e.kx < 0 ? -e.kx : e.kx, ay = e.ky < 0 ? -e.ky : e.ky;
#pragma unroll
        for (int f = 0; f < F; f++)
            sm.cxy[j][f] = g.mas_tab[f * (g.middle + 1) + ax] * g.mas_tab[f * (g.middle + 1) + ay];
    }
    __syncthreads();

    // per-thread base pointers: element kz of a row starts at base[f] + row byte offset
    const char *base[F];
#pragma unroll
    for (int f = 0; f < F; f++) base[f] = reinterpret_cast<const char *>(dk.p[f] + kzc);

    auto prefetch = [&](int j) {   // this thread's element of row j -> ring slot j % D   (non-bulk path)
        if constexpr (!BULK) {
            if (j < total) {
                const long long off = sm.ent[j].off;
#pragma unroll
                for (int f = 0; f < F; f++) cp_async8(&sm.z[j & (D - 1)][tid][f], base[f] + off);
            }
            cp_async_commit();     // one group per row, even when empty, so wait_group<N> counts rows
        }
    };
    // bulk path: thread 0 fills stage st (rows 4*st .. 4*st+3) with one cp.async.bulk per (row, field)
    const int nstages = (total + BULK_ROWS - 1) / BULK_ROWS;
    const int kz_last = min(kz_hi, kz_first + LANES - 1);
    auto produce = [&](int st) {
        if constexpr (BULK) {
            if (st < nstages) {
                RingSmemBulk<F> &sb = reinterpret_cast<RingSmemBulk<F> &>(sm);
                const int slot = st % BULK_SLOTS;
                if (st >= BULK_SLOTS) mbar_wait(&sb.empty[slot], (unsigned)((st / BULK_SLOTS - 1) & 1));
                const int r0 = st * BULK_ROWS, nr = min(BULK_ROWS, total - r0);
                unsigned bytes[BULK_ROWS], tot_bytes = 0;
                for (int u = 0; u < nr; u++) {
                    const int dlt = sb.ent[r0 + u].kx;
                    const int cnt = kz_last - (kz_first - dlt) + 1;           // elements needed from the window start
                    bytes[u] = (unsigned)(((cnt + 1) & ~1) * 8);               // even count: multiple of 16 bytes, stays inside the row (even dims)
                    tot_bytes += bytes[u] * F;
                }
                mbar_expect_tx(&sb.full[slot], tot_bytes);
                for (int u = 0; u < nr; u++) {
                    const int dlt = sb.ent[r0 + u].kx;
                    const long long off = sb.ent[r0 + u].off + 8ll * (kz_first - dlt);
#pragma unroll
                    for (int f = 0; f < F; f++)
                        bulk_g2s(&sb.z[slot][u][f][0], reinterpret_cast<const char *>(dk.p[f]) + off, bytes[u], &sb.full[slot]);
                }
            }
        }
    };
    if constexpr (BULK) {
        if (tid >= RING_T) {                    // dedicated producer warp: one elected lane streams every stage
            if (tid == RING_T)
                for (int st = 0; st < nstages; st++) produce(st);
            return;
        }
    } else {
        for (int j = 0; j < D; j++) prefetch(j);
    }

    double cz[F];
#pragma unroll
    for (int f = 0; f < F; f++) cz[f] = g.mas_tab[f * (g.middle + 1) + kzc];
    const int mid2 = g.middle * g.middle;

    // ring state (bins b0 / b0+1, 2-D bin (p, kz)) and span state (1-D bin kz)
    double lo3[3][Q], hi3[3][Q], lok = 0, hik = 0, loph = 0, hiph = 0, a2[Q], a1[Q];
    int locn = 0, hicn = 0, c2 = 0, c1 = 0;
#pragma unroll
    for (int q = 0; q < Q; q++) {
        a2[q] = 0; a1[q] = 0;
#pragma unroll
        for (int l = 0; l < 3; l++) { lo3[l][q] = 0; hi3[l][q] = 0; }
    }
    // group state (rows sharing r2)
    acc_t gq[Q], gph = 0;
    int gcnt = 0;
#pragma unroll
    for (int q = 0; q < Q; q++) gq[q] = 0;
    int cur_r2 = -1, ring_p = -1, ring_hi = 0, b0 = 0, thr = 0;
    bool sel = false, in1d = false;
    double w2 = 0, w4 = 0, kk = 0;
    float w2f = 0, w4f = 0;   // default mode: Legendre weights in fp32 (1e-7), applied to the fp32 group sums

    auto apply_group = [&]() {
        if (gcnt == 0) return;
        double v0[Q], v1[Q], v2[Q];
#pragma unroll
        for (int q = 0; q < Q; q++) {
            v0[q] = (double)gq[q];
            if (PRECISE) { v1[q] = v0[q] * w2; v2[q] = v0[q] * w4; }
            else { v1[q] = (double)((float)gq[q] * w2f); v2[q] = (double)((float)gq[q] * w4f); }
        }
        if (sel) {
#pragma unroll
            for (int q = 0; q < Q; q++) { hi3[0][q] += v0[q]; hi3[1][q] += v1[q]; hi3[2][q] += v2[q]; }
            hik += (double)gcnt * kk; hicn += gcnt; hiph += (double)gph;
        } else {
#pragma unroll
            for (int q = 0; q < Q; q++) { lo3[0][q] += v0[q]; lo3[1][q] += v1[q]; lo3[2][q] += v2[q]; }
            lok += (double)gcnt * kk; locn += gcnt; loph += (double)gph;
        }
#pragma unroll
        for (int q = 0; q < Q; q++) { a2[q] += v0[q]; if (in1d) a1[q] += v0[q]; gq[q] = 0; }
        c2 += gcnt;
        if (in1d) c1 += gcnt;
        gcnt = 0; gph = 0;
    };

    auto flush_bin3 = [&](int b, double (&s3)[3][Q], double ks, double ph, int cn) {
        if (cn == 0) return;
        red_add(g.sums + g.o_k3d + b, ks);
        red_add_u64(g.counts + g.o_n3d + b, (uint64_t)cn);
#pragma unroll
        for (int l = 0; l < 3; l++) {
#pragma unroll
            for (int f = 0; f < F; f++) red_add(g.sums + g.o_p3d + ((long long)b * 3 + l) * F + f, s3[l][f]);
#pragma unroll
            for (int x = 0; x < X; x++) red_add(g.sums + g.o_x3d + ((long long)b * 3 + l) * X + x, s3[l][F + x]);
        }
        if (PHASE) red_add(g.sums + g.o_phase + b, ph);
    };

    auto flush_ring = [&]() {
        if (active && c2 > 0) {
            flush_bin3(b0, lo3, lok, loph, locn);
            flush_bin3(b0 + 1, hi3, hik, hiph, hicn);
            const long long i2 = (long long)g.kmax_par1 * ring_p + kz;  // (kmax_par+1)*k_per + k_par, :371
            red_add_u64(g.counts + g.o_n2d + i2, (uint64_t)c2);
#pragma unroll
            for (int f = 0; f < F; f++) red_add(g.sums + g.o_p2d + i2 * F + f, a2[f]);
#pragma unroll
            for (int x = 0; x < X; x++) red_add(g.sums + g.o_x2d + i2 * X + x, a2[F + x]);
        }
        lok = hik = loph = hiph = 0; locn = hicn = c2 = 0;
#pragma unroll
        for (int q = 0; q < Q; q++) {
            a2[q] = 0;
#pragma unroll
            for (int l = 0; l < 3; l++) { lo3[l][q] = 0; hi3[l][q] = 0; }
        }
    };

    // new r2 group (CTA-uniform branch): bins, |k|, mu^2 and the Legendre weights, once per group
    auto new_group = [&](int r2) {
        apply_group();
        cur_r2 = r2;
        if (r2 >= ring_hi) {
            flush_ring();
            ring_p = isqrt_exact(r2);
            ring_hi = (ring_p + 1) * (ring_p + 1);
            b0 = isqrt_exact(ring_p * ring_p + kz2);
            thr = (b0 + 1) * (b0 + 1);
        }
        const int n = r2 + kz2;                 // >= 1 because kz >= 1
        sel = n >= thr;
        in1d = n <= mid2;                       // k <= middle, :364
        const double dn = (double)n;
        if (PRECISE) {
            kk = sqrt(dn);                      // :334
            const double mu = (double)kzc / kk; // :347
            const double mu2 = mu * mu;
            w2 = (3.0 * mu2 - 1.0) / 2.0;                       // :378
            w4 = (35.0 * mu2 * mu2 - 30.0 * mu2 + 3.0) / 8.0;   // :379
        } else {
            // Short dependency chains (every warp of the CTA takes this branch at the same time, so its
            // latency is exposed):  |k| = fp32 sqrt + one fp64 Newton correction (3e-14 relative);
            // mu^2 = kz^2/n and the Legendre weights in fp32 (1e-7 absolute on the weights).
            const float nf = (float)n;
            float rs, rc;
            asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(nf));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(nf));
            const double k0 = (double)(nf * rs);
            kk = fma(fma(-k0, k0, dn), (double)(0.5f * rs), k0);
            float mu2f = (float)kz2 * rc;
            mu2f = fminf(mu2f, 1.0f);
            w2f = fmaf(1.5f, mu2f, -0.5f);                                        // (3 mu^2 - 1)/2
            w4f = fmaf(fmaf(4.375f, mu2f, -3.75f), mu2f, 0.375f);                 // (35 mu^4 - 30 mu^2 + 3)/8
        }
    };

    // Per-mode work is split in two so that the arithmetic of several rows can overlap:
    //   mode_math : deconvolve, square, phase^2 -- depends only on the loaded value and the row's MAS
    //               factor, never on the binning state, so RING_B rows are evaluated back to back (ILP);
    //   mode_bin  : the (CTA-uniform) group / ring bookkeeping and the accumulation into the group sums.
    struct ModeVals { acc_t q[Q]; acc_t ph; };

    auto mode_math = [&](int j, const float2 (&z)[F], ModeVals &v) {
        float re[F], im[F];
#pragma unroll
        for (int f = 0; f < F; f++) {
            const float mf = (float)(sm.cxy[j][f] * cz[f]);  // double product -> float, :354
            re[f] = __fmul_rn(z[f].x, mf);                    // complex64 *= float, :355
            im[f] = __fmul_rn(z[f].y, mf);
            if (WB && active)
                *reinterpret_cast<float2 *>(const_cast<char *>(base[f]) + sm.ent[j].off) = make_float2(re[f], im[f]);
        }
        float d2_0 = 0.f;
#pragma unroll
        for (int f = 0; f < F; f++) {
            if (PRECISE) v.q[f] = (acc_t)((double)re[f] * (double)re[f] + (double)im[f] * (double)im[f]);  // :358-360
            else {
                const float d2 = fmaf(re[f], re[f], im[f] * im[f]);
                if (f == 0) d2_0 = d2;
                v.q[f] = (acc_t)d2;
            }
        }
        if (X > 0) {
            int ix = 0;
#pragma unroll
            for (int a = 0; a < F; a++)
#pragma unroll
                for (int b = a + 1; b < F; b++) {  // :721-722
                    if (g.ximag) {
                        if (PRECISE) v.q[F + ix] = (acc_t)((double)im[a] * (double)re[b] - (double)re[a] * (double)im[b]);
                        else v.q[F + ix] = (acc_t)fmaf(im[a], re[b], -(re[a] * im[b]));
                    } else if (PRECISE) v.q[F + ix] = (acc_t)((double)re[a] * (double)re[b] + (double)im[a] * (double)im[b]);
                    else v.q[F + ix] = (acc_t)fmaf(re[a], re[b], im[a] * im[b]);
                    ix++;
                }
        }
        v.ph = 0;
        if (PHASE) {
            if (PRECISE) d2_0 = fmaf(re[0], re[0], im[0] * im[0]);
            v.ph = (acc_t)phase_sq(re[0], d2_0);
        }
    };

    auto mode_bin = [&](int j, const ModeVals &v) {
        const int r2 = sm.ent[j].r2;
        if (r2 != cur_r2) new_group(r2);   // CTA-uniform
#pragma unroll
        for (int q = 0; q < Q; q++) gq[q] += v.q[q];
        if (PHASE) gph += v.ph;
        gcnt++;
    };

    constexpr int RING_B = BULK ? BULK_ROWS : ((F == 1) ? 4 : 2);   // rows per loop iteration
    static_assert(RING_B < D, "pipeline depth must exceed the batch");
    int j = 0;
    auto batch_math = [&](float2 (&z)[RING_B][F], ModeVals (&v)[RING_B]) {
        if (F == 1 && !PRECISE && !WB) {
            // packed path: (re,im) *= mf and (re^2, im^2) as one FMUL2 each, phase^2 of two rows per FFMA2 chain
#pragma unroll
            for (int u = 0; u < RING_B; u += 2) {
                const float mfa = (float)(sm.cxy[j + u][0] * cz[0]);          // double product -> float, :354
                const float mfb = (float)(sm.cxy[j + u + 1][0] * cz[0]);
                const unsigned long long da = mul2(pack2(z[u][0].x, z[u][0].y), pack2(mfa, mfa));          // :355
                const unsigned long long db = mul2(pack2(z[u + 1][0].x, z[u + 1][0].y), pack2(mfb, mfb));
                const float2 sa = unpack2(mul2(da, da)), sb2 = unpack2(mul2(db, db));                       // (re^2, im^2)
                const float d2a = sa.x + sa.y, d2b = sb2.x + sb2.y;
                v[u].q[0] = (acc_t)d2a;
                v[u + 1].q[0] = (acc_t)d2b;
                if (PHASE) {
                    const float2 ph = phase_sq2(sa.x, d2a, sb2.x, d2b);
                    v[u].ph = (acc_t)ph.x;
                    v[u + 1].ph = (acc_t)ph.y;
                } else {
                    v[u].ph = 0; v[u + 1].ph = 0;
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < RING_B; u++) mode_math(j + u, z[u], v[u]);
        }
    };

    if constexpr (BULK) {
        RingSmemBulk<F> &sb = reinterpret_cast<RingSmemBulk<F> &>(sm);
        const int lane = tid & 31;
        for (int st = 0; st * BULK_ROWS + BULK_ROWS <= total; st++, j += BULK_ROWS) {
            const int slot = st % BULK_SLOTS;
            mbar_wait(&sb.full[slot], (unsigned)((st / BULK_SLOTS) & 1));
            float2 z[RING_B][F];
#pragma unroll
            for (int u = 0; u < RING_B; u++) {
                const int idx = (active ? tid : 0) + sb.ent[j + u].kx;   // element kz sits at window index tid + dlt
#pragma unroll
                for (int f = 0; f < F; f++) z[u][f] = sb.z[slot][u][f][idx];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sb.empty[slot]);       // this warp is done with the stage
            ModeVals v[RING_B];
            batch_math(z, v);
#pragma unroll
            for (int u = 0; u < RING_B; u++) mode_bin(j + u, v[u]);
        }
        if (j < total) {   // last, partial stage
            const int st = j / BULK_ROWS, slot = st % BULK_SLOTS;
            mbar_wait(&sb.full[slot], (unsigned)((st / BULK_SLOTS) & 1));
            for (int u = 0; j < total; j++, u++) {
                float2 z0[F];
                const int idx = (active ? tid : 0) + sb.ent[j].kx;
#pragma unroll
                for (int f = 0; f < F; f++) z0[f] = sb.z[slot][u][f][idx];
                ModeVals v0;
                mode_math(j, z0, v0);
                mode_bin(j, v0);
            }
        }
    } else {
    for (; j + RING_B <= total; j += RING_B) {
        cp_async_wait<D - RING_B>();   // rows complete in order: rows j .. j+RING_B-1 have landed
        float2 z[RING_B][F];
#pragma unroll
        for (int u = 0; u < RING_B; u++)
#pragma unroll
            for (int f = 0; f < F; f++) z[u][f] = sm.z[(j + u) & (D - 1)][tid][f];
#pragma unroll
        for (int u = 0; u < RING_B; u++) prefetch(j + D + u);
        ModeVals v[RING_B];
        batch_math(z, v);
#pragma unroll
        for (int u = 0; u < RING_B; u++) mode_bin(j + u, v[u]);
    }
    if (j < total) cp_async_wait<0>();
    for (; j < total; j++) {
        float2 z0[F];
#pragma unroll
        for (int f = 0; f < F; f++) z0[f] = sm.z[j & (D - 1)][tid][f];
        ModeVals v0;
        mode_math(j, z0, v0);
        mode_bin(j, v0);
    }
    }
    cp_async_wait<0>();
    apply_group();
    flush_ring();
    if (active && c1 > 0) {
        red_add_u64(g.counts + g.o_n1d + kz, (uint64_t)c1);
#pragma unroll
        for (int f = 0; f < F; f++) red_add(g.sums + g.o_p1d + (long long)kz * F + f, a1[f]);
#pragma unroll
        for (int x = 0; x < X; x++) red_add(g.sums + g.o_x1d + (long long)kz * X + x, a1[F + x]);
    }
}

// ------------------------------------------------------------------------------------------------
// ring2 kernel: the production path for one field (class Pk), fp32 mode math, no write-back.
//
// Same idea as ring_kernel (rows sorted by r2, register accumulation, geometry once per r2-group), re-cut
// around what limited it on sm_100a (ncu: short-scoreboard/MIO stalls from ~7 shared-memory instructions and
// 2 XU conversions per mode, 13-19 % of the SM time idle in the tail of a static 2-wave grid):
//   * a thread owns TWO adjacent kz and every row access is one aligned 16-byte element pair (cp.async 16).
//     Rows are 8-byte aligned only in general (N/2+1 complex per row), so the row table is split by the parity
//     of the row's start element: in the even table a thread owns kz = (2t, 2t+1), in the odd table
//     (2t+1, 2t+2); both are then 16-byte aligned.  Pk() pads its own k-space rows to an even pitch
//     (pylb_fft_r2c_pitched), so there everything is in the even table and r2-groups keep their full length.
//     All per-row work (table reads, group test, loop control) is shared by two modes and the mode arithmetic
//     runs as packed f32x2 (FMUL2/FFMA2) across the two kz;
//   * the MAS factor (float)(Cx*Cy*Cz) (Pk_library.pyx:354) is formed from float-float splits of Cx*Cy and Cz
//     with 5 packed FMAs: error 2^-46 before the final rounding, i.e. the same float as the double product
//     except for ~2e-7 of the modes, which then differ by one fp32 ulp; no DMUL, no F2F;
//   * n = r2 + kz^2 grows along the sorted rows, so each kz walks its 3-D bins monotonically: ONE running 3-D
//     accumulator set per kz, flushed (red.global) when the bin changes; the 2-D bin (k_per = floor(sqrt(r2)))
//     changes CTA-uniformly; the 1-D bin is the kz itself;
//   * persistent CTAs with guided self-scheduling: a CTA takes spans of consecutive table rows from an atomic
//     counter; the first span of every CTA is long (pipeline fill/drain, the table prologue and the epilogue's
//     ~12 red.global per kz are paid per span: 443-row spans run 1.8x faster than 100-row spans), later spans
//     shrink geometrically, so the SMs finish together.  (Interleaving short chunks statically instead makes
//     every CTA flush the same few bins at the same time: 4x slower, red.global serialises per address.)
// Skipped here and left to special_kernel: kz = 0 and kz = N/2 (self-conjugate columns).  Even dims only.
// ------------------------------------------------------------------------------------------------
struct __align__(8) Row2 {
    long long off;   // BYTE offset of the row start
    int r2;
    float chi, clo;  // Cx*Cy split into two floats: chi = (float)c, clo = (float)(c - chi)
    short kx, ky;    // for special_kernel (skip rule, per-axis MAS factors)
};

constexpr int R2_T = 128;          // threads = kz pairs per CTA
constexpr int R2_ROWS = 4;         // rows per loop iteration
constexpr int R2_D = 16;           // rows in flight per thread (cp.async ring depth)
constexpr int R2_SPAN_MAX = 1024;  // rows per span (their table entries are staged in shared memory once)
constexpr int R2_SPAN_MIN = 32;
constexpr int R2_LEVELS = 16;
constexpr int R2_T3_VALS = 6;       // P0, P2, P4 sums, sum |k|, mode count, phase^2 per (3-D bin, kz)

struct Ring2Smem {
    float4 z[R2_D][R2_T];                     // 32 KB: thread t only ever touches z[.][t]
    long long off[R2_SPAN_MAX];
    int r2[R2_SPAN_MAX + 4];
    float2 c[R2_SPAN_MAX];
    int item;
};

// Guided schedule of one parity table: level l holds `per_level` spans of size[l] rows starting at base[l].
struct Ring2Sched {
    int nlevels, per_level;                   // per_level = spans per level and (parity, kz segment) group
    int base[2][R2_LEVELS + 1], size[2][R2_LEVELS];
};

// parity of the row's first element address in units of 8 bytes (bp = that of the field base pointer)
__device__ __forceinline__ int row_parity(int ix, int iy, const BinGeom &g, int bp) {
    return (int)(((long long)ix * g.stride_x + (long long)iy * g.stride_y + bp) & 1);
}

__global__ void row_keys2_kernel(unsigned *keys, unsigned *vals, int nrows, BinGeom g, int bp, int par_bit) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    const int ix = r / g.ny, iy = r - ix * g.ny;
    const int kx = wavenumber(g.x0 + ix, g.dims, g.middle), ky = wavenumber(g.y0 + iy, g.dims, g.middle);
    keys[r] = (unsigned)(kx * kx + ky * ky) | ((unsigned)row_parity(ix, iy, g, bp) << par_bit);
    vals[r] = (unsigned)r;
}

// cext[(f-1) * nrows + i] = the same split of Cx*Cy for field f = 1 .. F-1 (XPk: per-field MAS), in table order
__global__ void row_table2_kernel(const unsigned *keys, const unsigned *vals, Row2 *tab, int nrows, BinGeom g, int par_bit,
                                  float2 *cext) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    const int r = (int)vals[i];
    const int ix = r / g.ny, iy = r - ix * g.ny;
    const int kx = wavenumber(g.x0 + ix, g.dims, g.middle), ky = wavenumber(g.y0 + iy, g.dims, g.middle);
    for (int f = 1; f < g.F && cext != nullptr; f++) {
        const int m1 = g.middle + 1;
        const double cf = g.mas_tab[f * m1 + (kx < 0 ? -kx : kx)] * g.mas_tab[f * m1 + (ky < 0 ? -ky : ky)];
        const float hi = (float)cf;
        cext[(size_t)(f - 1) * nrows + i] = make_float2(hi, (float)(cf - (double)hi));
    }
    const double c = g.mas_tab[kx < 0 ? -kx : kx] * g.mas_tab[ky < 0 ? -ky : ky];   // field 0
    Row2 e;
    e.off = 8ll * ((long long)ix * g.stride_x + (long long)iy * g.stride_y);
    e.r2 = (int)(keys[i] & ((1u << par_bit) - 1u));
    e.chi = (float)c;
    e.clo = (float)(c - (double)e.chi);
    e.kx = (short)kx; e.ky = (short)ky;
    tab[i] = e;
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}

template <bool PHASE>
__global__ void __launch_bounds__(R2_T, 4)
ring2_kernel(BinGeom g, const float2 *__restrict__ dk, const Row2 *__restrict__ tab_all, int n0, int kz_hi, int nseg,
             int npar, Ring2Sched sc, int *__restrict__ counter, double *__restrict__ t3, int t3_kz, long long *__restrict__ trace) {
    extern __shared__ __align__(128) unsigned char ring2_raw[];
    Ring2Smem &sm = *reinterpret_cast<Ring2Smem *>(ring2_raw);
    const int tid = threadIdx.x;
    const int groups = nseg * npar;
    const int nitems = sc.nlevels * sc.per_level * groups;
    const int mid2 = g.middle * g.middle;
    const unsigned long long m1 = pack2(-1.0f, -1.0f);

    for (;;) {
        // ---- next span: item -> (level, span k, group = (parity, kz segment))
        __syncthreads();                        // everyone is done with the previous span's tables
        if (tid == 0) sm.item = atomicAdd(counter, 1);
        __syncthreads();
        const int item = sm.item;
        if (item >= nitems) break;
        const int lev = item / (sc.per_level * groups), w = item - lev * (sc.per_level * groups);
        const int grp = w % groups, k = w / groups;
        const int seg = grp % nseg;
        const int P = npar == 2 ? grp / nseg : (sc.base[0][sc.nlevels] > 0 ? 0 : 1);   // parity table of this span
        const int i0 = sc.base[P][lev] + k * sc.size[P][lev];
        const int total = min(sc.size[P][lev], sc.base[P][lev + 1] - i0);
        if (total <= 0) continue;
        const int npairs = P ? (kz_hi - 1) / 2 + 1 : kz_hi / 2 + 1;
        const int first_pair = seg * R2_T;
        const int cnt = min(R2_T, npairs - first_pair);            // kz pairs served by this span
        if (cnt <= 0) continue;
        const Row2 *tab = tab_all + (P ? n0 : 0) + i0;
        long long t_begin = 0;
        if (trace && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
        for (int j = tid; j < total; j += R2_T) {
            const Row2 e = tab[j];
            sm.off[j] = e.off; sm.r2[j] = e.r2; sm.c[j] = make_float2(e.chi, e.clo);
        }
        __syncthreads();

        // this thread's 16 bytes inside a row: kz = 2*(first_pair + tid) + P, 16-byte aligned by construction
        const char *colbase = reinterpret_cast<const char *>(dk) + 8ll * (2 * (first_pair + tid) + P);
        const bool loader = tid < cnt;
        auto prefetch = [&](int j) {
            if (j < total && loader) cp_async16(&sm.z[j & (R2_D - 1)][tid], colbase + sm.off[j]);
            cp_async_commit();                  // one group per row, even when empty, so wait_group<N> counts rows
        };
        for (int j = 0; j < R2_D; j++) prefetch(j);

        // ---- per-kz constants (s = 0: kzA, s = 1: kzB = kzA + 1)
        const int kzA = 2 * (first_pair + tid) + P;
        int kz[2], kz2[2];
        bool act[2];
        float kz2f[2], czh[2], czl[2];
#pragma unroll
        for (int s = 0; s < 2; s++) {
            kz[s] = kzA + s;
            act[s] = loader && kz[s] >= 1 && kz[s] <= kz_hi;
            const int kc = min(kz[s], g.middle);
            kz2[s] = kc * kc;
            kz2f[s] = (float)kz2[s];
            const double czd = g.mas_tab[kc];
            czh[s] = (float)czd;
            czl[s] = (float)(czd - (double)czh[s]);
        }
        const unsigned long long czh2 = pack2(czh[0], czh[1]);
        const unsigned long long nczh2 = pack2(-czh[0], -czh[1]), nczl2 = pack2(-czl[0], -czl[1]);

        // ---- accumulators
        double s3[2][3], ks[2], ph[2], a2[2], a1[2], kk[2];
        int cn[2], c1[2], bin[2], thr[2];
        float w2f[2], w4f[2], gq[2], gph[2];
        bool in1d[2];
#pragma unroll
        for (int s = 0; s < 2; s++) {
            s3[s][0] = s3[s][1] = s3[s][2] = 0; ks[s] = ph[s] = a2[s] = a1[s] = 0; kk[s] = 0;
            cn[s] = c1[s] = 0; bin[s] = 0; thr[s] = 0; w2f[s] = w4f[s] = 0; gq[s] = gph[s] = 0; in1d[s] = false;
        }
        int gcnt = 0, c2 = 0, cur_r2 = -1, ring_p = 0, ring_hi = 0;

        auto apply_group = [&]() {
            if (gcnt == 0) return;
            const double dg = (double)gcnt;
#pragma unroll
            for (int s = 0; s < 2; s++) {
                const double v0 = (double)gq[s];
                s3[s][0] += v0;
                s3[s][1] += (double)(gq[s] * w2f[s]);
                s3[s][2] += (double)(gq[s] * w4f[s]);
                ks[s] = fma(dg, kk[s], ks[s]);
                cn[s] += gcnt;
                if (PHASE) ph[s] += (double)gph[s];
                a2[s] += v0;
                if (in1d[s]) { a1[s] += v0; c1[s] += gcnt; }
                gq[s] = 0; gph[s] = 0;
            }
            c2 += gcnt;
            gcnt = 0;
        };
        // 3-D bins go to a kz-private table t3[bin][value][kz] (no two lanes, and hardly any two CTAs, ever hit
        // the same address; red.global on the shared bins themselves serialises: 47 G/s measured, and a kernel
        // whose CTAs flush together then spends longer draining atomics than streaming data).  ring2_finish_kernel
        // folds the table into the bins.
        auto flush3 = [&](int s) {
            if (act[s] && cn[s] > 0) {
                double *t = t3 + ((long long)bin[s] * R2_T3_VALS) * t3_kz + kz[s];
                red_add(t, s3[s][0]);
                red_add(t + t3_kz, s3[s][1]);
                red_add(t + 2 * t3_kz, s3[s][2]);
                red_add(t + 3 * t3_kz, ks[s]);
                red_add(t + 4 * t3_kz, (double)cn[s]);        // exact: counts < 2^53
                if (PHASE) red_add(t + 5 * t3_kz, ph[s]);
            }
            s3[s][0] = s3[s][1] = s3[s][2] = 0; ks[s] = 0; ph[s] = 0; cn[s] = 0;
        };
        auto flush2 = [&]() {
            if (c2 > 0) {
#pragma unroll
                for (int s = 0; s < 2; s++)
                    if (act[s]) {
                        const long long i2 = (long long)g.kmax_par1 * ring_p + kz[s];   // (kmax_par+1)*k_per + k_par, :371
                        red_add_u64(g.counts + g.o_n2d + i2, (uint64_t)c2);
                        red_add(g.sums + g.o_p2d + i2, a2[s]);
                    }
            }
            a2[0] = a2[1] = 0; c2 = 0;
        };
        // new r2 group: CTA-uniform branch.  Bins, |k|, mu^2 and the Legendre weights once per (group, kz).
        auto new_group = [&](int r2) {
            apply_group();
            cur_r2 = r2;
            if (r2 >= ring_hi) {                   // k_per = floor(sqrt(r2)) changes, uniform
                flush2();
                ring_p = isqrt_exact(r2);
                ring_hi = (ring_p + 1) * (ring_p + 1);
            }
#pragma unroll
            for (int s = 0; s < 2; s++) {
                const int n = r2 + kz2[s];         // >= 1 for active lanes (kz >= 1)
                if (n >= thr[s]) {                 // k_index changes for this kz (divergent, about once per ring)
                    flush3(s);
                    bin[s] = isqrt_exact(n);
                    thr[s] = (bin[s] + 1) * (bin[s] + 1);
                }
                in1d[s] = n <= mid2;               // k <= middle, :364
                const float nf = (float)n;
                float rs;
                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(fmaxf(nf, 1.0f)));
                const float h = nf * rs;                                              // ~sqrt(n)
                const double k0 = (double)h;
                kk[s] = fma(fma(-k0, k0, (double)n), (double)(0.5f * rs), k0);       // fp32 sqrt + one fp64 Newton step: 3e-14
                const float rs2 = rs * rs;
                const float rn = fmaf(rs2, fmaf(-h, rs, 1.0f), rs2);                  // 1/n = rs^2 (1 + (1 - n rs^2)): ~1 ulp, no second MUFU
                const float mu2f = fminf(kz2f[s] * rn, 1.0f);                         // mu^2 = kz^2/n, :347
                w2f[s] = fmaf(1.5f, mu2f, -0.5f);                                     // (3 mu^2 - 1)/2, :378
                w4f[s] = fmaf(fmaf(4.375f, mu2f, -3.75f), mu2f, 0.375f);              // (35 mu^4 - 30 mu^2 + 3)/8, :379
            }
        };

        // deconvolve + square (+ phase^2) of one row's element pair
        auto row_math = [&](const float4 z, const float2 c, float (&d2)[2], float (&p2)[2]) {
            const unsigned long long chi2 = pack2(c.x, c.x), clo2 = pack2(c.y, c.y);
            const unsigned long long p = mul2(chi2, czh2);
            unsigned long long t = fma2(chi2, nczh2, p);            // -(chi*czh - p), exact
            t = fma2(chi2, nczl2, t);
            t = fma2(clo2, nczh2, t);
            const float2 mf = unpack2(fma2(t, m1, p));              // (float)(Cx*Cy*Cz) for kzA, kzB, :354
            const unsigned long long dA = mul2(pack2(z.x, z.y), pack2(mf.x, mf.x));   // complex64 *= float, :355
            const unsigned long long dB = mul2(pack2(z.z, z.w), pack2(mf.y, mf.y));
            const float2 sA = unpack2(mul2(dA, dA)), sB = unpack2(mul2(dB, dB));      // (re^2, im^2)
            d2[0] = sA.x + sA.y;
            d2[1] = sB.x + sB.y;
            if (PHASE) {
                const float2 q = phase_sq2(sA.x, d2[0], sB.x, d2[1]);
                p2[0] = q.x; p2[1] = q.y;
            } else {
                p2[0] = p2[1] = 0.f;
            }
        };
        auto row_bin = [&](int r2, const float (&d2)[2], const float (&p2)[2]) {
            if (r2 != cur_r2) new_group(r2);       // CTA-uniform
            gq[0] += d2[0]; gq[1] += d2[1];
            if (PHASE) { gph[0] += p2[0]; gph[1] += p2[1]; }
            gcnt++;
        };

        int j = 0;
        for (; j + R2_ROWS <= total; j += R2_ROWS) {
            cp_async_wait<R2_D - R2_ROWS>();      // rows complete in order: rows j .. j+3 have landed
            float4 z[R2_ROWS];
#pragma unroll
            for (int u = 0; u < R2_ROWS; u++) z[u] = sm.z[(j + u) & (R2_D - 1)][tid];
#pragma unroll
            for (int u = 0; u < R2_ROWS; u++) prefetch(j + R2_D + u);
            const int4 r2v = *reinterpret_cast<const int4 *>(&sm.r2[j]);
            const float4 c01 = *reinterpret_cast<const float4 *>(&sm.c[j]);
            const float4 c23 = *reinterpret_cast<const float4 *>(&sm.c[j + 2]);
            float d2[R2_ROWS][2], p2[R2_ROWS][2];
            row_math(z[0], make_float2(c01.x, c01.y), d2[0], p2[0]);
            row_math(z[1], make_float2(c01.z, c01.w), d2[1], p2[1]);
            row_math(z[2], make_float2(c23.x, c23.y), d2[2], p2[2]);
            row_math(z[3], make_float2(c23.z, c23.w), d2[3], p2[3]);
            row_bin(r2v.x, d2[0], p2[0]);
            row_bin(r2v.y, d2[1], p2[1]);
            row_bin(r2v.z, d2[2], p2[2]);
            row_bin(r2v.w, d2[3], p2[3]);
        }
        cp_async_wait<0>();
        for (; j < total; j++) {
            float d2[2], p2[2];
            row_math(sm.z[j & (R2_D - 1)][tid], sm.c[j], d2, p2);
            row_bin(sm.r2[j], d2, p2);
        }
        apply_group();
        flush3(0);
        flush3(1);
        flush2();
#pragma unroll
        for (int s = 0; s < 2; s++)
            if (act[s] && c1[s] > 0) {
                red_add_u64(g.counts + g.o_n1d + kz[s], (uint64_t)c1[s]);
                red_add(g.sums + g.o_p1d + kz[s], a1[s]);
            }
        if (trace && tid == 0) {                // PYLB_RING2_TRACE: per-span (start ns, end ns, SM, rows, first r2)
            long long t_end;
            unsigned smid;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            trace[5 * item + 0] = t_begin; trace[5 * item + 1] = t_end; trace[5 * item + 2] = smid;
            trace[5 * item + 3] = total; trace[5 * item + 4] = sm.r2[0];
        }
    }
}

// t3[bin][value][kz] -> the 3-D bins.  One warp per (bin, value).
__global__ void __launch_bounds__(256)
ring2_finish_kernel(BinGeom g, const double *__restrict__ t3, int t3_kz, int nbins, int want_phase) {
    const int wid = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (wid >= nbins * R2_T3_VALS) return;
    const int b = wid / R2_T3_VALS, v = wid - b * R2_T3_VALS;
    if (v == 5 && !want_phase) return;
    const double *row = t3 + (long long)wid * t3_kz;
    double acc = 0;
    // n = r2 + kz^2 >= kz^2: bin b only ever holds kz <= b
    const int kz_end = min(t3_kz, b + 1);
    for (int k = lane; k < kz_end; k += 32) acc += row[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane != 0 || acc == 0.0) return;
    if (v < 3) red_add(g.sums + g.o_p3d + (long long)b * 3 + v, acc);
    else if (v == 3) red_add(g.sums + g.o_k3d + b, acc);
    else if (v == 4) red_add_u64(g.counts + g.o_n3d + b, (uint64_t)(acc + 0.5));
    else red_add(g.sums + g.o_phase + b, acc);
}

// ------------------------------------------------------------------------------------------------
// ring2x kernel: ring2 for F = 2 or 3 fields (class XPk): auto spectra of every field and the cross spectra of every
// pair, per-field MAS factors, no phase term, no write-back.  Same row tables, span schedule, two kz per thread, 16-byte
// cp.async loads (one per field and row) and kz-private flush table as ring2_kernel; every accumulator exists Q = F + X
// times (X = F(F-1)/2 pairs in the reference's order (0,1),(0,2),(1,2), Pk_library.pyx:628-737).  |delta_k|^2 and the
// cross terms Re(d_a conj d_b) are formed in fp32 from the fp32-deconvolved modes -- the reference's operands -- summed
// over the rows of one r2-group in fp32 and in fp64 from there on.
// ------------------------------------------------------------------------------------------------
constexpr int R2X_D = 8;            // rows in flight per thread and field

template <int F>
struct Ring2xSmem {
    float4 z[F][R2X_D][R2_T];
    long long off[R2_SPAN_MAX];
    int r2[R2_SPAN_MAX + 4];
    float2 c[F][R2_SPAN_MAX];
    int item;
};

template <int F>
__global__ void __launch_bounds__(R2_T, 2)
ring2x_kernel(BinGeom g, FieldPtrs dk, const Row2 *__restrict__ tab_all, const float2 *__restrict__ cext, int nrows_all, int n0,
              int kz_hi, int nseg, int npar, Ring2Sched sc, int *__restrict__ counter, double *__restrict__ t3, int t3_kz) {
    constexpr int X = F * (F - 1) / 2, Q = F + X, NV = 3 * Q + 2;
    extern __shared__ __align__(128) unsigned char ring2x_raw[];
    Ring2xSmem<F> &sm = *reinterpret_cast<Ring2xSmem<F> *>(ring2x_raw);
    const int tid = threadIdx.x;
    const int groups = nseg * npar;
    const int nitems = sc.nlevels * sc.per_level * groups;
    const int mid2 = g.middle * g.middle, m1i = g.middle + 1;
    const unsigned long long m1 = pack2(-1.0f, -1.0f);

    for (;;) {
        __syncthreads();                        // everyone is done with the previous span's tables
        if (tid == 0) sm.item = atomicAdd(counter, 1);
        __syncthreads();
        const int item = sm.item;
        if (item >= nitems) break;
        const int lev = item / (sc.per_level * groups), w = item - lev * (sc.per_level * groups);
        const int grp = w % groups, k = w / groups;
        const int seg = grp % nseg;
        const int P = npar == 2 ? grp / nseg : (sc.base[0][sc.nlevels] > 0 ? 0 : 1);   // parity table of this span
        const int i0 = sc.base[P][lev] + k * sc.size[P][lev];
        const int total = min(sc.size[P][lev], sc.base[P][lev + 1] - i0);
        if (total <= 0) continue;
        const int npairs = P ? (kz_hi - 1) / 2 + 1 : kz_hi / 2 + 1;
        const int first_pair = seg * R2_T;
        const int cnt = min(R2_T, npairs - first_pair);            // kz pairs served by this span
        if (cnt <= 0) continue;
        const int t0 = (P ? n0 : 0) + i0;                          // first table row of this span
        for (int j = tid; j < total; j += R2_T) {
            const Row2 e = tab_all[t0 + j];
            sm.off[j] = e.off; sm.r2[j] = e.r2; sm.c[0][j] = make_float2(e.chi, e.clo);
#pragma unroll
            for (int f = 1; f < F; f++) sm.c[f][j] = cext[(size_t)(f - 1) * nrows_all + t0 + j];
        }
        __syncthreads();

        const long long coloff = 8ll * (2 * (first_pair + tid) + P);   // this thread's 16 bytes inside a row
        const bool loader = tid < cnt;
        auto prefetch = [&](int j) {
            if (j < total && loader) {
#pragma unroll
                for (int f = 0; f < F; f++)
                    cp_async16(&sm.z[f][j & (R2X_D - 1)][tid], reinterpret_cast<const char *>(dk.p[f]) + coloff + sm.off[j]);
            }
            cp_async_commit();                  // one group per row, even when empty, so wait_group<N> counts rows
        };
        for (int j = 0; j < R2X_D; j++) prefetch(j);

        // ---- per-kz constants (s = 0: kzA, s = 1: kzB = kzA + 1), per field
        const int kzA = 2 * (first_pair + tid) + P;
        int kz[2], kz2[2];
        bool act[2];
        float kz2f[2];
        unsigned long long czh2[F], nczh2[F], nczl2[F];
#pragma unroll
        for (int s = 0; s < 2; s++) {
            kz[s] = kzA + s;
            act[s] = loader && kz[s] >= 1 && kz[s] <= kz_hi;
            const int kc = min(kz[s], g.middle);
            kz2[s] = kc * kc;
            kz2f[s] = (float)kz2[s];
        }
#pragma unroll
        for (int f = 0; f < F; f++) {
            float h[2], l[2];
#pragma unroll
            for (int s = 0; s < 2; s++) {
                const double czd = g.mas_tab[f * m1i + min(kz[s], g.middle)];
                h[s] = (float)czd;
                l[s] = (float)(czd - (double)h[s]);
            }
            czh2[f] = pack2(h[0], h[1]); nczh2[f] = pack2(-h[0], -h[1]); nczl2[f] = pack2(-l[0], -l[1]);
        }

        // ---- accumulators
        double s3[2][Q][3], a2[2][Q], a1[2][Q], ks[2], kk[2];
        float gq[2][Q], w2f[2], w4f[2];
        int cn[2], c1[2], bin[2], thr[2];
        bool in1d[2];
#pragma unroll
        for (int s = 0; s < 2; s++) {
#pragma unroll
            for (int q = 0; q < Q; q++) { s3[s][q][0] = s3[s][q][1] = s3[s][q][2] = 0; a2[s][q] = a1[s][q] = 0; gq[s][q] = 0; }
            ks[s] = 0; kk[s] = 0; cn[s] = c1[s] = 0; bin[s] = 0; thr[s] = 0; w2f[s] = w4f[s] = 0; in1d[s] = false;
        }
        int gcnt = 0, c2 = 0, cur_r2 = -1, ring_p = 0, ring_hi = 0;

        auto apply_group = [&]() {
            if (gcnt == 0) return;
            const double dg = (double)gcnt;
#pragma unroll
            for (int s = 0; s < 2; s++) {
#pragma unroll
                for (int q = 0; q < Q; q++) {
                    const double v0 = (double)gq[s][q];
                    s3[s][q][0] += v0;
                    s3[s][q][1] += (double)(gq[s][q] * w2f[s]);
                    s3[s][q][2] += (double)(gq[s][q] * w4f[s]);
                    a2[s][q] += v0;
                    if (in1d[s]) a1[s][q] += v0;
                    gq[s][q] = 0;
                }
                ks[s] = fma(dg, kk[s], ks[s]);
                cn[s] += gcnt;
                if (in1d[s]) c1[s] += gcnt;
            }
            c2 += gcnt;
            gcnt = 0;
        };
        auto flush3 = [&](int s) {
            if (act[s] && cn[s] > 0) {
                double *t = t3 + ((long long)bin[s] * NV) * t3_kz + kz[s];
#pragma unroll
                for (int q = 0; q < Q; q++)
#pragma unroll
                    for (int l = 0; l < 3; l++) red_add(t + (long long)(q * 3 + l) * t3_kz, s3[s][q][l]);
                red_add(t + (long long)(3 * Q) * t3_kz, ks[s]);
                red_add(t + (long long)(3 * Q + 1) * t3_kz, (double)cn[s]);     // exact: counts < 2^53
            }
#pragma unroll
            for (int q = 0; q < Q; q++) s3[s][q][0] = s3[s][q][1] = s3[s][q][2] = 0;
            ks[s] = 0; cn[s] = 0;
        };
        auto flush2 = [&]() {
            if (c2 > 0) {
#pragma unroll
                for (int s = 0; s < 2; s++)
                    if (act[s]) {
                        const long long i2 = (long long)g.kmax_par1 * ring_p + kz[s];   // (kmax_par+1)*k_per + k_par
                        red_add_u64(g.counts + g.o_n2d + i2, (uint64_t)c2);
#pragma unroll
                        for (int f = 0; f < F; f++) red_add(g.sums + g.o_p2d + i2 * F + f, a2[s][f]);
#pragma unroll
                        for (int x = 0; x < X; x++) red_add(g.sums + g.o_x2d + i2 * X + x, a2[s][F + x]);
                    }
            }
#pragma unroll
            for (int s = 0; s < 2; s++)
#pragma unroll
                for (int q = 0; q < Q; q++) a2[s][q] = 0;
            c2 = 0;
        };
        auto new_group = [&](int r2) {
            apply_group();
            cur_r2 = r2;
            if (r2 >= ring_hi) {                   // k_per = floor(sqrt(r2)) changes, uniform
                flush2();
                ring_p = isqrt_exact(r2);
                ring_hi = (ring_p + 1) * (ring_p + 1);
            }
#pragma unroll
            for (int s = 0; s < 2; s++) {
                const int n = r2 + kz2[s];
                if (n >= thr[s]) {
                    flush3(s);
                    bin[s] = isqrt_exact(n);
                    thr[s] = (bin[s] + 1) * (bin[s] + 1);
                }
                in1d[s] = n <= mid2;
                const float nf = (float)n;
                float rs;
                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(fmaxf(nf, 1.0f)));
                const float h = nf * rs;
                const double k0 = (double)h;
                kk[s] = fma(fma(-k0, k0, (double)n), (double)(0.5f * rs), k0);
                const float rs2 = rs * rs;
                const float rn = fmaf(rs2, fmaf(-h, rs, 1.0f), rs2);
                const float mu2f = fminf(kz2f[s] * rn, 1.0f);
                w2f[s] = fmaf(1.5f, mu2f, -0.5f);
                w4f[s] = fmaf(fmaf(4.375f, mu2f, -3.75f), mu2f, 0.375f);
            }
        };
        // one row: deconvolve every field's element pair, then the Q products for kzA (s = 0) and kzB (s = 1)
        auto row = [&](int j) {
            unsigned long long dA[F], dB[F];
#pragma unroll
            for (int f = 0; f < F; f++) {
                const float4 z = sm.z[f][j & (R2X_D - 1)][tid];
                const float2 c = sm.c[f][j];
                const unsigned long long chi2 = pack2(c.x, c.x), clo2 = pack2(c.y, c.y);
                const unsigned long long p = mul2(chi2, czh2[f]);
                unsigned long long t = fma2(chi2, nczh2[f], p);
                t = fma2(chi2, nczl2[f], t);
                t = fma2(clo2, nczh2[f], t);
                const float2 mf = unpack2(fma2(t, m1, p));              // (float)(Cx*Cy*Cz) of field f for kzA, kzB
                dA[f] = mul2(pack2(z.x, z.y), pack2(mf.x, mf.x));       // complex64 *= float
                dB[f] = mul2(pack2(z.z, z.w), pack2(mf.y, mf.y));
            }
            const int r2 = sm.r2[j];
            if (r2 != cur_r2) new_group(r2);       // CTA-uniform
#pragma unroll
            for (int f = 0; f < F; f++) {
                const float2 sA = unpack2(mul2(dA[f], dA[f])), sB = unpack2(mul2(dB[f], dB[f]));
                gq[0][f] += sA.x + sA.y;
                gq[1][f] += sB.x + sB.y;
            }
            int ix = 0;
#pragma unroll
            for (int a = 0; a < F; a++)
#pragma unroll
                for (int b = a + 1; b < F; b++) {
                    const float2 xA = unpack2(mul2(dA[a], dA[b])), xB = unpack2(mul2(dB[a], dB[b]));   // (re re, im im)
                    gq[0][F + ix] += xA.x + xA.y;
                    gq[1][F + ix] += xB.x + xB.y;
                    ix++;
                }
            gcnt++;
        };

        int j = 0;
        for (; j + 2 <= total; j += 2) {
            cp_async_wait<R2X_D - 2>();           // rows complete in order: rows j, j+1 have landed
            row(j);
            row(j + 1);
            prefetch(j + R2X_D);
            prefetch(j + R2X_D + 1);
        }
        cp_async_wait<0>();
        for (; j < total; j++) row(j);
        apply_group();
        flush3(0);
        flush3(1);
        flush2();
#pragma unroll
        for (int s = 0; s < 2; s++)
            if (act[s] && c1[s] > 0) {
                red_add_u64(g.counts + g.o_n1d + kz[s], (uint64_t)c1[s]);
#pragma unroll
                for (int f = 0; f < F; f++) red_add(g.sums + g.o_p1d + (long long)kz[s] * F + f, a1[s][f]);
#pragma unroll
                for (int x = 0; x < X; x++) red_add(g.sums + g.o_x1d + (long long)kz[s] * X + x, a1[s][F + x]);
            }
    }
}

// t3[bin][value][kz] -> the 3-D bins of ring2x.  One warp per (bin, value); values: Q x 3 multipole sums, sum |k|, count.
template <int F>
__global__ void __launch_bounds__(256)
ring2x_finish_kernel(BinGeom g, const double *__restrict__ t3, int t3_kz, int nbins) {
    constexpr int X = F * (F - 1) / 2, Q = F + X, NV = 3 * Q + 2;
    const int wid = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (wid >= nbins * NV) return;
    const int b = wid / NV, v = wid - b * NV;
    const double *rowp = t3 + (long long)wid * t3_kz;
    double acc = 0;
    const int kz_end = min(t3_kz, b + 1);          // n = r2 + kz^2 >= kz^2: bin b only ever holds kz <= b
    for (int k = lane; k < kz_end; k += 32) acc += rowp[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane != 0 || acc == 0.0) return;
    if (v < 3 * Q) {
        const int q = v / 3, l = v - 3 * q;
        if (q < F) red_add(g.sums + g.o_p3d + ((long long)b * 3 + l) * F + q, acc);
        else red_add(g.sums + g.o_x3d + ((long long)b * 3 + l) * X + (q - F), acc);
    } else if (v == 3 * Q) red_add(g.sums + g.o_k3d + b, acc);
    else red_add_u64(g.counts + g.o_n3d + b, (uint64_t)(acc + 0.5));
}

// Running per-lane sums for the special-column kernels.  A warp walks a CONTIGUOUS range of the r2-sorted row table, 32 rows
// per step; with kz fixed both k_index = floor(sqrt(r2 + kz^2)) and k_per = floor(sqrt(r2)) are non-decreasing along the
// table and the 32 rows of a step hardly ever span more than two bins.  Every lane therefore keeps sums for the bin of the
// step's first row (`cur`) and for the next one; only when `cur` moves are both summed over the warp and added to the bins
// (one red.global per value) -- a handful of times per warp instead of once per 32 rows and distinct bin, which is what
// bounded these kernels: red.global on the same few hot bins from every warp serialises.
template <int NV>
struct RunSums {
    int cur;
    double a[NV], b[NV];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int q = 0; q < NV; q++) { a[q] = 0; b[q] = 0; }
    }
    __device__ __forceinline__ void init() { cur = -2; clear(); }
    // all lanes; flush(key, sums) runs on lane 0 for every key with a non-zero LAST value (the mode count)
    template <class FLUSH>
    __device__ __forceinline__ void flush_all(FLUSH flush) {
        const unsigned full = 0xffffffffu;
#pragma unroll
        for (int half = 0; half < 2; half++) {
            double t[NV];
#pragma unroll
            for (int q = 0; q < NV; q++) {
                t[q] = half ? b[q] : a[q];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) t[q] += __shfl_xor_sync(full, t[q], o);
            }
            if ((threadIdx.x & 31) == 0 && t[NV - 1] > 0.5) flush(cur + half, t);
        }
        clear();
    }
    // kmin: the key of the step's first row (warp-uniform)
    template <class FLUSH>
    __device__ __forceinline__ void step(int kmin, bool valid, int key, const double (&v)[NV], FLUSH flush) {
        if (kmin != cur) {
            if (cur >= 0) flush_all(flush);
            cur = kmin;
        }
        if (valid) {
            if (key == cur) {
#pragma unroll
                for (int q = 0; q < NV; q++) a[q] += v[q];
            } else if (key == cur + 1) {
#pragma unroll
                for (int q = 0; q < NV; q++) b[q] += v[q];
            } else {
                flush(key, v);                  // more than two bins in one step: only among the first few hundred rows
            }
        }
    }
};

// rows [lo, hi) of warp `w` of `nw`: equal contiguous shares, whole steps of 32 rows
__device__ __forceinline__ void warp_rows(int nrows, int &lo, int &hi) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long steps = (nrows + 31) / 32, per = (steps + nw - 1) / nw;
    lo = (int)min((long long)nrows, w * per * 32);
    hi = (int)min((long long)nrows, (w + 1) * per * 32);
}

// special columns of the ring2x path: special2_kernel for F = 2 or 3 fields (no phase term)
template <int F>
__global__ void __launch_bounds__(256)
special2x_kernel(BinGeom g, FieldPtrs dk, const Row2 *__restrict__ tab, int nrows) {
    constexpr int X = F * (F - 1) / 2, Q = F + X, N3 = 3 * Q + 2, N2 = Q + 1;
    const int kz = blockIdx.y == 0 ? 0 : g.middle, m1 = g.middle + 1;
    const int lane = threadIdx.x & 31;
    double v1[N2];                                             // the 1-D bin is kz itself: one running sum per thread
#pragma unroll
    for (int q = 0; q < N2; q++) v1[q] = 0;
    RunSums<N3> r3;
    RunSums<N2> r2s;
    r3.init(); r2s.init();
    auto flush3 = [&](int b, const double (&t)[N3]) {
#pragma unroll
        for (int l = 0; l < 3; l++) {
#pragma unroll
            for (int f = 0; f < F; f++) red_add(g.sums + g.o_p3d + ((long long)b * 3 + l) * F + f, t[3 * f + l]);
#pragma unroll
            for (int x = 0; x < X; x++) red_add(g.sums + g.o_x3d + ((long long)b * 3 + l) * X + x, t[3 * (F + x) + l]);
        }
        red_add(g.sums + g.o_k3d + b, t[3 * Q]);
        red_add_u64(g.counts + g.o_n3d + b, (uint64_t)(t[3 * Q + 1] + 0.5));
    };
    auto flush2 = [&](int b, const double (&t)[N2]) {
        const long long i2 = (long long)g.kmax_par1 * b + kz;
#pragma unroll
        for (int f = 0; f < F; f++) red_add(g.sums + g.o_p2d + i2 * F + f, t[f]);
#pragma unroll
        for (int x = 0; x < X; x++) red_add(g.sums + g.o_x2d + i2 * X + x, t[F + x]);
        red_add_u64(g.counts + g.o_n2d + i2, (uint64_t)(t[Q] + 0.5));
    };
    int lo, hi;
    warp_rows(nrows, lo, hi);
    for (int i0 = lo; i0 < hi; i0 += 32) {                     // warp-uniform
        const int i = i0 + lane;
        bool valid = i < hi;
        Row2 e;
        e.off = 0; e.r2 = 0; e.chi = e.clo = 0.f; e.kx = e.ky = 0;
        if (valid) e = tab[i];
        const int kx = e.kx, ky = e.ky;
        // keep one of each conjugate pair, Pk_library.pyx:326-330 / :640-644
        if (kx < 0) valid = false;
        if ((kx == 0 || (kx == g.middle && g.even)) && ky < 0) valid = false;
        const int n = e.r2 + kz * kz;
        const int b3 = isqrt_exact(n), b2 = isqrt_exact(e.r2);
        double v3[N3], v2[N2];
#pragma unroll
        for (int q = 0; q < N3; q++) v3[q] = 0;
#pragma unroll
        for (int q = 0; q < N2; q++) v2[q] = 0;
        if (valid) {
            const double k = sqrt((double)n);
            const double mu = (n == 0) ? 0.0 : (double)kz / k;
            const double mu2 = mu * mu;
            const double w2 = (3.0 * mu2 - 1.0) / 2.0, w4 = (35.0 * mu2 * mu2 - 30.0 * mu2 + 3.0) / 8.0;
            const int ax = kx, ay = ky < 0 ? -ky : ky;
            double re[F], im[F], d[Q];
#pragma unroll
            for (int f = 0; f < F; f++) {
                const float mf = (float)(g.mas_tab[f * m1 + ax] * g.mas_tab[f * m1 + ay] * g.mas_tab[f * m1 + kz]);
                const float2 z = *(reinterpret_cast<const float2 *>(reinterpret_cast<const char *>(dk.p[f]) + e.off) + kz);
                re[f] = (double)__fmul_rn(z.x, mf); im[f] = (double)__fmul_rn(z.y, mf);
                d[f] = re[f] * re[f] + im[f] * im[f];
            }
            int ix = 0;
#pragma unroll
            for (int a = 0; a < F; a++)
#pragma unroll
                for (int b = a + 1; b < F; b++) { d[F + ix] = re[a] * re[b] + im[a] * im[b]; ix++; }
            const bool in1d = n <= g.middle * g.middle;
#pragma unroll
            for (int q = 0; q < Q; q++) {
                v3[3 * q] = d[q]; v3[3 * q + 1] = d[q] * w2; v3[3 * q + 2] = d[q] * w4;
                v2[q] = d[q];
                if (in1d) v1[q] += d[q];
            }
            v3[3 * Q] = k; v3[3 * Q + 1] = 1.0;
            v2[Q] = 1.0;
            if (in1d) v1[Q] += 1.0;
        }
        r3.step(__shfl_sync(0xffffffffu, b3, 0), valid, b3, v3, flush3);
        r2s.step(__shfl_sync(0xffffffffu, b2, 0), valid, b2, v2, flush2);
    }
    if (r3.cur >= 0) r3.flush_all(flush3);
    if (r2s.cur >= 0) r2s.flush_all(flush2);
    warp_reduce_by_key<N2>(0, true, v1, [&](int, const double (&t)[N2]) {
        if (t[Q] > 0.5) {
#pragma unroll
            for (int f = 0; f < F; f++) red_add(g.sums + g.o_p1d + (long long)kz * F + f, t[f]);
#pragma unroll
            for (int x = 0; x < X; x++) red_add(g.sums + g.o_x1d + (long long)kz * X + x, t[F + x]);
            red_add_u64(g.counts + g.o_n1d + kz, (uint64_t)(t[Q] + 0.5));
        }
    });
}

static double ring2_first_share() {
    static double f = 0;
    if (f == 0) {
        const char *e = getenv("PYLB_RING2_SHARE");
        f = e ? atof(e) : 0.8;
        if (f < 0.05) f = 0.05;
        if (f > 1.0) f = 1.0;
    }
    return f;
}

// level sizes for a table of `rows` rows worked on by `per_level` CTAs: the first level takes `share` of the
// rows, every later level the same share of what is left, down to R2_SPAN_MIN rows per span
static void ring2_levels(int rows, int per_level, int P, Ring2Sched &sc, int &nlev) {
    int done = 0, l = 0;
    const double share = ring2_first_share();
    while (done < rows && l < R2_LEVELS) {
        const int left = rows - done;
        long long sz = (long long)(share * left / per_level) + 1;
        if (l == R2_LEVELS - 1) sz = (left + per_level - 1) / per_level;
        sz = (sz + R2_ROWS - 1) / R2_ROWS * R2_ROWS;
        if (sz < R2_SPAN_MIN) sz = R2_SPAN_MIN;
        if (sz > R2_SPAN_MAX) sz = R2_SPAN_MAX;
        sc.base[P][l] = done;
        sc.size[P][l] = (int)sz;
        const long long cover = sz * per_level;
        done = cover >= left ? rows : done + (int)cover;
        l++;
    }
    for (int q = l; q <= R2_LEVELS; q++) sc.base[P][q] = done;     // done == rows unless the table is huge (checked by the caller)
    for (int q = l; q < R2_LEVELS; q++) sc.size[P][q] = R2_SPAN_MIN;
    if (l > nlev) nlev = l;
}

template <bool PHASE>
static int launch_ring2(const BinGeom &g, const float2 *dk, const Row2 *tab, int n0, int n1, int kz_hi, int *counter, int nbins3,
                       double *t3, cudaStream_t st) {
    const size_t smem = sizeof(Ring2Smem);
    static bool attr_set = false;
    if (!attr_set) {
        PYLB_CHECK(cudaFuncSetAttribute(ring2_kernel<PHASE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ring2_kernel<PHASE>, R2_T, smem);
    if (occ < 1) occ = 1;
    const int npairs = kz_hi / 2 + 1;                           // parity 0 (never fewer than parity 1)
    const int nseg = (npairs + R2_T - 1) / R2_T;
    const int npar = (n0 > 0) + (n1 > 0);
    const int ctas = sm_count() * occ;                          // persistent: one resident wave
    int per_level = ctas / (nseg * npar);
    if (per_level < 1) per_level = 1;
    {   // large tables: more spans per level than resident CTAs, so that level 0 still covers its share
        const long long need = (long long)(ring2_first_share() * (n0 > n1 ? n0 : n1) / R2_SPAN_MAX) + 1;
        if (need > per_level) per_level = (int)need;
    }
    Ring2Sched sc;
    memset(&sc, 0, sizeof(sc));
    sc.per_level = per_level;
    int nlev = 0;
    ring2_levels(n0, per_level, 0, sc, nlev);
    ring2_levels(n1, per_level, 1, sc, nlev);
    sc.nlevels = nlev;
    PYLB_REQUIRE(sc.base[0][nlev] == n0 && sc.base[1][nlev] == n1, "ring2: row table too large for the span schedule");
    PYLB_CHECK(cudaMemsetAsync(counter, 0, sizeof(int), st));
    const int t3_kz = (kz_hi + 1 + 3) & ~3;                     // kz = 0 .. kz_hi, rows padded to 32 bytes; t3 comes zeroed
    long long *trace = nullptr;
    const char *trace_path = getenv("PYLB_RING2_TRACE");
    const int nitems = sc.nlevels * sc.per_level * nseg * npar;
    if (trace_path) {
        PYLB_CHECK(cudaMalloc(&trace, sizeof(long long) * 5 * (size_t)nitems));
        PYLB_CHECK(cudaMemset(trace, 0, sizeof(long long) * 5 * (size_t)nitems));
    }
    timing_begin(PYLB_T_RING, st);
    ring2_kernel<PHASE><<<ctas, R2_T, smem, st>>>(g, dk, tab, n0, kz_hi, nseg, npar, sc, counter, t3, t3_kz, trace);
    timing_end(PYLB_T_RING, st);
    PYLB_LAUNCH_CHECK();
    ring2_finish_kernel<<<(unsigned)(((long long)nbins3 * R2_T3_VALS * 32 + 255) / 256), 256, 0, st>>>(g, t3, t3_kz, nbins3, PHASE ? 1 : 0);
    PYLB_LAUNCH_CHECK();
    if (trace) {                                // debugging aid: dump the per-span timeline (synchronises)
        std::vector<long long> h(5 * (size_t)nitems);
        PYLB_CHECK(cudaStreamSynchronize(st));
        PYLB_CHECK(cudaMemcpy(h.data(), trace, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
        cudaFree(trace);
        if (FILE *f = fopen(trace_path, "w")) {
            fprintf(f, "# item start_ns end_ns smid rows first_r2   (levels %d, spans per level %d, groups %d)\n", sc.nlevels,
                    sc.per_level, nseg * npar);
            long long t0 = 0;
            for (int i = 0; i < nitems; i++) if (h[5 * i + 3] && (!t0 || h[5 * i] < t0)) t0 = h[5 * i];
            for (int i = 0; i < nitems; i++)
                if (h[5 * i + 3])
                    fprintf(f, "%d %lld %lld %lld %lld %lld\n", i, h[5 * i] - t0, h[5 * i + 1] - t0, h[5 * i + 2], h[5 * i + 3], h[5 * i + 4]);
            fclose(f);
        }
    }
    return 0;
}

template <int F>
static int launch_ring2x(const BinGeom &g, const FieldPtrs &dk, const Row2 *tab, const float2 *cext, int nrows, int n0, int n1,
                         int kz_hi, int *counter, int nbins3, double *t3, cudaStream_t st) {
    constexpr int NV = 3 * (F + F * (F - 1) / 2) + 2;
    const size_t smem = sizeof(Ring2xSmem<F>);
    PYLB_CHECK(cudaFuncSetAttribute(ring2x_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ring2x_kernel<F>, R2_T, smem);
    if (occ < 1) occ = 1;
    const int npairs = kz_hi / 2 + 1;                           // parity 0 (never fewer than parity 1)
    const int nseg = (npairs + R2_T - 1) / R2_T;
    const int npar = (n0 > 0) + (n1 > 0);
    const int ctas = sm_count() * occ;                          // persistent: one resident wave
    int per_level = ctas / (nseg * npar);
    if (per_level < 1) per_level = 1;
    {
        const long long need = (long long)(ring2_first_share() * (n0 > n1 ? n0 : n1) / R2_SPAN_MAX) + 1;
        if (need > per_level) per_level = (int)need;
    }
    Ring2Sched sc;
    memset(&sc, 0, sizeof(sc));
    sc.per_level = per_level;
    int nlev = 0;
    ring2_levels(n0, per_level, 0, sc, nlev);
    ring2_levels(n1, per_level, 1, sc, nlev);
    sc.nlevels = nlev;
    PYLB_REQUIRE(sc.base[0][nlev] == n0 && sc.base[1][nlev] == n1, "ring2x: row table too large for the span schedule");
    PYLB_CHECK(cudaMemsetAsync(counter, 0, sizeof(int), st));
    const int t3_kz = (kz_hi + 1 + 3) & ~3;                     // t3 comes zeroed
    timing_begin(PYLB_T_RING, st);
    ring2x_kernel<F><<<ctas, R2_T, smem, st>>>(g, dk, tab, cext, nrows, n0, kz_hi, nseg, npar, sc, counter, t3, t3_kz);
    timing_end(PYLB_T_RING, st);
    PYLB_LAUNCH_CHECK();
    ring2x_finish_kernel<F><<<(unsigned)(((long long)nbins3 * NV * 32 + 255) / 256), 256, 0, st>>>(g, t3, t3_kz, nbins3);
    PYLB_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// special columns kz = 0 and kz = dims/2 (even dims) for the ring path (line of sight = z).
// These are the self-conjugate planes: one of each conjugate pair is kept (:326-330).  A thread owns
// one (span of the r2-sorted row list, special kz).  With kz fixed, n = r2 + kz^2 grows along the
// sorted list, so k_index and k_per are monotone: ONE running 3-D bin and ONE running 2-D bin per
// thread, flushed with red.global when they change.  fp64 per mode (2*N^2 modes in total: irrelevant).
// ------------------------------------------------------------------------------------------------
constexpr int SPECIAL_ROWS = 16;   // rows per thread

template <int F, class ROW>
__global__ void __launch_bounds__(128)
special_kernel(BinGeom g, FieldPtrs dk, const ROW *__restrict__ tab, int nrows, int nplanes, int want_phase, int write_back) {
    constexpr int X = F * (F - 1) / 2;
    constexpr int Q = F + X;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int plane = gid % nplanes, span = gid / nplanes;
    const int i0 = span * SPECIAL_ROWS, i1 = min(nrows, i0 + SPECIAL_ROWS);
    if (i0 >= i1) return;
    const int kz = plane == 0 ? 0 : g.middle;
    const int kz2 = kz * kz, mid2 = g.middle * g.middle, m1 = g.middle + 1;
    double cz[F];
#pragma unroll
    for (int f = 0; f < F; f++) cz[f] = g.mas_tab[f * m1 + kz];

    int b3 = -1, b2 = -1, c3 = 0, c2 = 0, c1 = 0;
    double s3[3][Q], ks = 0, ph = 0, s2[Q], s1[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) { s2[q] = 0; s1[q] = 0; s3[0][q] = s3[1][q] = s3[2][q] = 0; }

    auto flush3 = [&]() {
        if (c3) {
            red_add(g.sums + g.o_k3d + b3, ks);
            red_add_u64(g.counts + g.o_n3d + b3, (uint64_t)c3);
#pragma unroll
            for (int l = 0; l < 3; l++) {
#pragma unroll
                for (int f = 0; f < F; f++) red_add(g.sums + g.o_p3d + ((long long)b3 * 3 + l) * F + f, s3[l][f]);
#pragma unroll
                for (int x = 0; x < X; x++) red_add(g.sums + g.o_x3d + ((long long)b3 * 3 + l) * X + x, s3[l][F + x]);
            }
            if (want_phase) red_add(g.sums + g.o_phase + b3, ph);
        }
        c3 = 0; ks = 0; ph = 0;
#pragma unroll
        for (int q = 0; q < Q; q++) s3[0][q] = s3[1][q] = s3[2][q] = 0;
    };
    auto flush2 = [&]() {
        if (c2) {
            const long long i2 = (long long)g.kmax_par1 * b2 + kz;
            red_add_u64(g.counts + g.o_n2d + i2, (uint64_t)c2);
#pragma unroll
            for (int f = 0; f < F; f++) red_add(g.sums + g.o_p2d + i2 * F + f, s2[f]);
#pragma unroll
            for (int x = 0; x < X; x++) red_add(g.sums + g.o_x2d + i2 * X + x, s2[F + x]);
        }
        c2 = 0;
#pragma unroll
        for (int q = 0; q < Q; q++) s2[q] = 0;
    };

    for (int i = i0; i < i1; i++) {
        const ROW e = tab[i];
        const int kx = e.kx, ky = e.ky;
        // keep one of each conjugate pair, :326-330
        if (kx < 0) continue;
        if ((kx == 0 || (kx == g.middle && g.even)) && ky < 0) continue;
        const int n = e.r2 + kz2;
        const int k_index = isqrt_exact(n), k_per = isqrt_exact(e.r2);
        if (k_index != b3) { flush3(); b3 = k_index; }
        if (k_per != b2) { flush2(); b2 = k_per; }
        const double k = sqrt((double)n);
        const double mu = (n == 0) ? 0.0 : (double)kz / k;
        const double mu2 = mu * mu;
        const double w2 = (3.0 * mu2 - 1.0) / 2.0, w4 = (35.0 * mu2 * mu2 - 30.0 * mu2 + 3.0) / 8.0;
        const bool in1d = n <= mid2;
        const int ax = kx, ay = ky < 0 ? -ky : ky;
        double re[F], im[F], v[Q];
#pragma unroll
        for (int f = 0; f < F; f++) {
            const float mf = (float)(g.mas_tab[f * m1 + ax] * g.mas_tab[f * m1 + ay] * cz[f]);
            float2 *zp = reinterpret_cast<float2 *>(reinterpret_cast<char *>(dk.p[f]) + e.off) + kz;
            const float2 z = *zp;
            const float r = __fmul_rn(z.x, mf), q = __fmul_rn(z.y, mf);
            if (write_back) *zp = make_float2(r, q);
            re[f] = (double)r; im[f] = (double)q;
            v[f] = re[f] * re[f] + im[f] * im[f];
        }
        if (X > 0) {
            int ix = 0;
#pragma unroll
            for (int a = 0; a < F; a++)
#pragma unroll
                for (int b = a + 1; b < F; b++) { v[F + ix] = g.ximag ? im[a] * re[b] - re[a] * im[b] : re[a] * re[b] + im[a] * im[b]; ix++; }
        }
        if (want_phase) ph += (double)phase_sq((float)re[0], (float)v[0]);   // atan2(re, |delta_k|)^2, :361 (1e-7 polynomial)
#pragma unroll
        for (int q = 0; q < Q; q++) {
            s3[0][q] += v[q]; s3[1][q] += v[q] * w2; s3[2][q] += v[q] * w4;
            s2[q] += v[q];
            if (in1d) s1[q] += v[q];
        }
        ks += k; c3++; c2++;
        if (in1d) c1++;
    }
    flush3();
    flush2();
    if (c1) {
        red_add_u64(g.counts + g.o_n1d + kz, (uint64_t)c1);
#pragma unroll
        for (int f = 0; f < F; f++) red_add(g.sums + g.o_p1d + (long long)kz * F + f, s1[f]);
#pragma unroll
        for (int x = 0; x < X; x++) red_add(g.sums + g.o_x1d + (long long)kz * X + x, s1[F + x]);
    }
}

// ------------------------------------------------------------------------------------------------
// special columns for the ring2 path (one field, no write-back): a warp walks a contiguous range of the r2-sorted table,
// one thread per row and step, with running sums for the current and the next bin (RunSums above) -- red.global on the
// hot bins serialises: the 16-rows-per-thread kernel above spends 50-75 us there at 512^3, a warp-level reduce-by-key
// per 32 rows 84 us at 1024^3.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
special2_kernel(BinGeom g, const float2 *__restrict__ dk, const Row2 *__restrict__ tab, int nrows, int want_phase) {
    const int kz = blockIdx.y == 0 ? 0 : g.middle;
    const int lane = threadIdx.x & 31;
    double v1[2] = {0, 0};                                     // the 1-D bin is kz itself: one running sum per thread
    RunSums<6> r3;                                             // P0, P2, P4 sums, phase^2, |k|, count (count last)
    RunSums<2> r2s;
    r3.init(); r2s.init();
    auto flush3 = [&](int b, const double (&t)[6]) {
        red_add(g.sums + g.o_p3d + (long long)b * 3 + 0, t[0]);
        red_add(g.sums + g.o_p3d + (long long)b * 3 + 1, t[1]);
        red_add(g.sums + g.o_p3d + (long long)b * 3 + 2, t[2]);
        if (want_phase) red_add(g.sums + g.o_phase + b, t[3]);
        red_add(g.sums + g.o_k3d + b, t[4]);
        red_add_u64(g.counts + g.o_n3d + b, (uint64_t)(t[5] + 0.5));
    };
    auto flush2 = [&](int b, const double (&t)[2]) {
        const long long i2 = (long long)g.kmax_par1 * b + kz;
        red_add(g.sums + g.o_p2d + i2, t[0]);
        red_add_u64(g.counts + g.o_n2d + i2, (uint64_t)(t[1] + 0.5));
    };
    int lo, hi;
    warp_rows(nrows, lo, hi);
    for (int i0 = lo; i0 < hi; i0 += 32) {                     // warp-uniform
        const int i = i0 + lane;
        bool valid = i < hi;
        Row2 e;
        e.off = 0; e.r2 = 0; e.chi = e.clo = 0.f; e.kx = e.ky = 0;
        if (valid) e = tab[i];
        const int kx = e.kx, ky = e.ky;
        // keep one of each conjugate pair, :326-330
        if (kx < 0) valid = false;
        if ((kx == 0 || (kx == g.middle && g.even)) && ky < 0) valid = false;
        const int n = e.r2 + kz * kz;
        const int b3 = isqrt_exact(n), b2 = isqrt_exact(e.r2);
        double v3[6] = {0, 0, 0, 0, 0, 0}, v2[2] = {0, 0};
        if (valid) {
            const double k = sqrt((double)n);
            const double mu = (n == 0) ? 0.0 : (double)kz / k;
            const double mu2 = mu * mu;
            const int ax = kx, ay = ky < 0 ? -ky : ky;
            const float mf = (float)(g.mas_tab[ax] * g.mas_tab[ay] * g.mas_tab[kz]);   // :354
            const float2 z = *(reinterpret_cast<const float2 *>(reinterpret_cast<const char *>(dk) + e.off) + kz);
            const float r = __fmul_rn(z.x, mf), q = __fmul_rn(z.y, mf);                // :355
            const double d2 = (double)r * (double)r + (double)q * (double)q;           // :358-360
            v3[0] = d2;
            v3[1] = d2 * ((3.0 * mu2 - 1.0) / 2.0);
            v3[2] = d2 * ((35.0 * mu2 * mu2 - 30.0 * mu2 + 3.0) / 8.0);
            v3[3] = want_phase ? (double)phase_sq(r, (float)d2) : 0.0;                 // atan2(re, |delta_k|)^2, :361
            v3[4] = k;
            v3[5] = 1.0;
            v2[0] = d2; v2[1] = 1.0;
            if (n <= g.middle * g.middle) { v1[0] += d2; v1[1] += 1.0; }
        }
        r3.step(__shfl_sync(0xffffffffu, b3, 0), valid, b3, v3, flush3);
        r2s.step(__shfl_sync(0xffffffffu, b2, 0), valid, b2, v2, flush2);
    }
    if (r3.cur >= 0) r3.flush_all(flush3);
    if (r2s.cur >= 0) r2s.flush_all(flush2);
    warp_reduce_by_key<2>(0, true, v1, [&](int, const double (&t)[2]) {
        if (t[1] > 0.5) {
            red_add(g.sums + g.o_p1d + kz, t[0]);
            red_add_u64(g.counts + g.o_n1d + kz, (uint64_t)(t[1] + 0.5));
        }
    });
}

template <int F, class ROW>
static int launch_special(const BinGeom &g, const FieldPtrs &dk, const ROW *tab, int nrows, int want_phase,
                          int write_back, cudaStream_t st) {
    const int nplanes = (g.even && g.middle > 0) ? 2 : 1;
    const long long nthreads = (long long)((nrows + SPECIAL_ROWS - 1) / SPECIAL_ROWS) * nplanes;
    if (nthreads == 0) return 0;
    special_kernel<F, ROW><<<(unsigned)((nthreads + 127) / 128), 128, 0, st>>>(g, dk, tab, nrows, nplanes, want_phase, write_back);
    PYLB_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// generic kernel: one thread per mode, kzz = kz_start + j*kz_step for j < kz_num
// ------------------------------------------------------------------------------------------------
template <bool WB>
__global__ void __launch_bounds__(256)
generic_kernel(BinGeom g, FieldPtrs dk, long long nmodes, int kz_start, int kz_step, int kz_num, int want_phase) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nmodes) return;
    const int j = (int)(idx % kz_num);
    const int r = (int)(idx / kz_num);
    const int kzz = kz_start + j * kz_step;
    const int ix = r / g.ny, iy = r - ix * g.ny;
    const int kx = wavenumber(g.x0 + ix, g.dims, g.middle);
    const int ky = wavenumber(g.y0 + iy, g.dims, g.middle);
    const int kz = kzz;  // kzz <= middle always
    // one of each conjugate pair on the self-conjugate planes, :326-330
    if (kz == 0 || (kz == g.middle && g.even)) {
        if (kx < 0) return;
        if (kx == 0 || (kx == g.middle && g.even)) {
            if (ky < 0) return;
        }
    }
    const int n = kx * kx + ky * ky + kz * kz;
    const int k_index = isqrt_exact(n);  // <int>sqrt(...), :334-335
    const double k = sqrt((double)n);
    int k_par, k_per;  // :338-343
    if (g.axis == 0) { k_par = kx; k_per = isqrt_exact(ky * ky + kz * kz); }
    else if (g.axis == 1) { k_par = ky; k_per = isqrt_exact(kx * kx + kz * kz); }
    else { k_par = kz; k_per = isqrt_exact(kx * kx + ky * ky); }
    const double mu = (n == 0) ? 0.0 : (double)k_par / k;  // :346-347
    const double mu2 = mu * mu;
    const double w2 = (3.0 * mu2 - 1.0) / 2.0;
    const double w4 = (35.0 * mu2 * mu2 - 30.0 * mu2 + 3.0) / 8.0;
    if (k_par < 0) k_par = -k_par;  // :351
    const bool in1d = n <= g.middle * g.middle;
    const long long i2 = (long long)g.kmax_par1 * k_per + k_par;
    const int F = g.F, X = g.X, m1 = g.middle + 1;
    const int ax = kx < 0 ? -kx : kx, ay = ky < 0 ? -ky : ky;

    if (in1d) red_add_u64(g.counts + g.o_n1d + k_par, 1);
    red_add_u64(g.counts + g.o_n2d + i2, 1);
    red_add_u64(g.counts + g.o_n3d + k_index, 1);
    red_add(g.sums + g.o_k3d + k_index, k);

    const long long off = (long long)ix * g.stride_x + (long long)iy * g.stride_y + kzz;
    float re[MAX_F], im[MAX_F];
    for (int f = 0; f < F; f++) {
        const double c = g.mas_tab[f * m1 + ax] * g.mas_tab[f * m1 + ay] * g.mas_tab[f * m1 + kz];
        const float mf = (float)c;
        const float2 z = dk.p[f][off];
        re[f] = __fmul_rn(z.x, mf);
        im[f] = __fmul_rn(z.y, mf);
        if (WB) dk.p[f][off] = make_float2(re[f], im[f]);
        const double d2 = (double)re[f] * (double)re[f] + (double)im[f] * (double)im[f];
        if (f == 0 && want_phase) {
            const double ph = atan2((double)re[0], sqrt(d2));  // :361
            red_add(g.sums + g.o_phase + k_index, ph * ph);
        }
        if (in1d) red_add(g.sums + g.o_p1d + (long long)k_par * F + f, d2);
        red_add(g.sums + g.o_p2d + i2 * F + f, d2);
        red_add(g.sums + g.o_p3d + ((long long)k_index * 3 + 0) * F + f, d2);
        red_add(g.sums + g.o_p3d + ((long long)k_index * 3 + 1) * F + f, d2 * w2);
        red_add(g.sums + g.o_p3d + ((long long)k_index * 3 + 2) * F + f, d2 * w4);
    }
    int xi = 0;
    for (int a = 0; a < F; a++)
        for (int b = a + 1; b < F; b++) {
            const double dx = g.ximag ? (double)im[a] * (double)re[b] - (double)re[a] * (double)im[b]
                                      : (double)re[a] * (double)re[b] + (double)im[a] * (double)im[b];
            if (in1d) red_add(g.sums + g.o_x1d + (long long)k_par * X + xi, dx);
            red_add(g.sums + g.o_x2d + i2 * X + xi, dx);
            red_add(g.sums + g.o_x3d + ((long long)k_index * 3 + 0) * X + xi, dx);
            red_add(g.sums + g.o_x3d + ((long long)k_index * 3 + 1) * X + xi, dx * w2);
            red_add(g.sums + g.o_x3d + ((long long)k_index * 3 + 2) * X + xi, dx * w4);
            xi++;
        }
}

static int launch_generic(const BinGeom &g, const FieldPtrs &dk, int kz_start, int kz_step, int kz_num,
                          int want_phase, int write_back, cudaStream_t st) {
    const long long nmodes = (long long)g.nx * g.ny * kz_num;
    if (nmodes == 0) return 0;
    const unsigned blocks = (unsigned)((nmodes + 255) / 256);
    if (write_back) generic_kernel<true><<<blocks, 256, 0, st>>>(g, dk, nmodes, kz_start, kz_step, kz_num, want_phase);
    else generic_kernel<false><<<blocks, 256, 0, st>>>(g, dk, nmodes, kz_start, kz_step, kz_num, want_phase);
    PYLB_LAUNCH_CHECK();
    return 0;
}

template <int F, bool PHASE, bool WB, bool PRECISE, bool BULK>
static int launch_ring_v(const BinGeom &g, const FieldPtrs &dk, const RowEnt *tab, int nrows, int kz_hi, cudaStream_t st) {
    constexpr int LANES = BULK ? RING_T - 1 : RING_T;
    const int nseg = (kz_hi + LANES - 1) / LANES;
    const size_t smem = BULK ? sizeof(RingSmemBulk<F>) : sizeof(RingSmem<F>);
    static bool attr_set = false;
    if (!attr_set) {
        PYLB_CHECK(cudaFuncSetAttribute(ring_kernel<F, PHASE, WB, PRECISE, BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    // two waves of resident CTAs; spans of 32..RING_SPAN_MAX rows
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ring_kernel<F, PHASE, WB, PRECISE, BULK>, BULK ? RING_T + 32 : RING_T, smem);
    if (occ < 1) occ = 1;
    int nspan = (sm_count() * occ * 2 + nseg - 1) / nseg;
    if (nspan > (nrows + 31) / 32) nspan = (nrows + 31) / 32;
    if (nspan < 1) nspan = 1;
    int rows_per_span = (nrows + nspan - 1) / nspan;
    if (rows_per_span > RING_SPAN_MAX) rows_per_span = RING_SPAN_MAX;
    nspan = (nrows + rows_per_span - 1) / rows_per_span;
    dim3 grid(nspan, nseg);
    timing_begin(PYLB_T_RING, st);
    ring_kernel<F, PHASE, WB, PRECISE, BULK><<<grid, BULK ? RING_T + 32 : RING_T, smem, st>>>(g, dk, tab, nrows, rows_per_span, kz_hi);
    timing_end(PYLB_T_RING, st);
    PYLB_LAUNCH_CHECK();
    return 0;
}

template <int F, bool BULK>
static int launch_ring_b(const BinGeom &g, const FieldPtrs &dk, const RowEnt *tab, int nrows, int kz_hi,
                         int want_phase, int write_back, int precise, cudaStream_t st) {
    const int sel = (want_phase ? 4 : 0) | (write_back ? 2 : 0) | (precise ? 1 : 0);
    switch (sel) {
        case 0: return launch_ring_v<F, false, false, false, BULK>(g, dk, tab, nrows, kz_hi, st);
        case 1: return launch_ring_v<F, false, false, true, BULK>(g, dk, tab, nrows, kz_hi, st);
        case 2: return launch_ring_v<F, false, true, false, BULK>(g, dk, tab, nrows, kz_hi, st);
        case 3: return launch_ring_v<F, false, true, true, BULK>(g, dk, tab, nrows, kz_hi, st);
        case 4: return launch_ring_v<F, true, false, false, BULK>(g, dk, tab, nrows, kz_hi, st);
        case 5: return launch_ring_v<F, true, false, true, BULK>(g, dk, tab, nrows, kz_hi, st);
        case 6: return launch_ring_v<F, true, true, false, BULK>(g, dk, tab, nrows, kz_hi, st);
        default: return launch_ring_v<F, true, true, true, BULK>(g, dk, tab, nrows, kz_hi, st);
    }
}

// bulk (TMA 1-D) loads need even dims (the window may take one extra element, which exists only then),
// 16-byte aligned field pointers and 8-byte-multiple row strides (always true)
static bool g_allow_bulk = false;   // opt-in (PYLB_BIN_BULK): matches cp.async for Pk, loses for F >= 2 (profiles/r1_ring_variants.txt)
template <int F>
static int launch_ring_f(const BinGeom &g, const FieldPtrs &dk, const RowEnt *tab, int nrows, int kz_hi,
                         int want_phase, int write_back, int precise, cudaStream_t st) {
    bool bulk = g_allow_bulk && g.even;
    for (int f = 0; f < F; f++) bulk = bulk && (((uintptr_t)dk.p[f] & 15) == 0);
    if (bulk) return launch_ring_b<F, true>(g, dk, tab, nrows, kz_hi, want_phase, write_back, precise, st);
    return launch_ring_b<F, false>(g, dk, tab, nrows, kz_hi, want_phase, write_back, precise, st);
}

static int bits_for(unsigned v) {
    int b = 1;
    while (b < 32 && (v >> b)) b++;
    return b;
}

static bool g_allow_ring2 = true;

struct Ring2Key {
    int dims, x0, nx, y0, ny;
    long long stride_x, stride_y;
    int bp, F, mas[3];
};
struct Ring2Cache {
    Ring2Key key = {};
    Row2 *tab = nullptr;
    float2 *cext = nullptr;          // [(F-1)][nrows]: Cx*Cy splits of fields 1 .. F-1, in table order
    cudaEvent_t ready = nullptr;
};
// A few tables per device (a caller alternating between two grid sizes, or Pk and a slab engine, no longer rebuilds one
// every call), replaced round-robin; guarded by a mutex because host threads may share a device.
constexpr int R2_CACHE_WAYS = 4;
struct Ring2DeviceCache {
    Ring2Cache way[R2_CACHE_WAYS];
    int next = 0;
};
static std::mutex g_ring2_mutex;
static std::map<int, Ring2DeviceCache> g_ring2_cache;      // keyed by the full device ordinal

// one field, even dims, default arithmetic, no write-back: parity-split row table + ring2_kernel
static int run_ring2(const BinGeom &g, const FieldPtrs &dk, int want_phase, cudaStream_t st) {
    const int nrows = g.nx * g.ny;
    const int kz_hi = g.middle - 1;
    const int bp = (int)(((uintptr_t)dk.p[0] >> 3) & 1);
    // rows whose first element sits on an odd multiple of 8 bytes
    const long long sxo = g.stride_x & 1, syo = g.stride_y & 1;
    const long long ox = sxo ? g.nx / 2 : 0, ex = g.nx - ox, oy = syo ? g.ny / 2 : 0, ey = g.ny - oy;
    long long n1 = ox * ey + ex * oy;
    if (bp) n1 = (long long)nrows - n1;
    const int n0 = (int)(nrows - n1);

    // The row table depends only on the k-space window, the strides, the pointer parity and the MAS exponent:
    // keep the last one per device (24 B per row; rebuilding costs a radix sort of N^2 keys and ~8 launches).
    int dev = 0;
    PYLB_CHECK(cudaGetDevice(&dev));
    Ring2Key key;
    memset(&key, 0, sizeof(key));               // the struct has padding and is compared with memcmp
    key.dims = g.dims; key.x0 = g.x0; key.nx = g.nx; key.y0 = g.y0; key.ny = g.ny;
    key.stride_x = g.stride_x; key.stride_y = g.stride_y; key.bp = bp; key.F = g.F;
    for (int f = 0; f < g.F && f < 3; f++) key.mas[f] = g.mas_idx[f];
    std::lock_guard<std::mutex> lock(g_ring2_mutex);       // held until the kernels using the table are queued on `st`
    Ring2DeviceCache &dc = g_ring2_cache[dev];
    Ring2Cache *hit = nullptr;
    for (auto &w : dc.way)
        if (w.tab && memcmp(&w.key, &key, sizeof(key)) == 0) hit = &w;
    if (!hit) {
        Ring2Cache &c = dc.way[dc.next];
        dc.next = (dc.next + 1) % R2_CACHE_WAYS;
        if (c.tab) { cudaFree(c.tab); c.tab = nullptr; }          // synchronises: nobody is reading it any more
        if (c.cext) { cudaFree(c.cext); c.cext = nullptr; }
        if (!c.ready) PYLB_CHECK(cudaEventCreateWithFlags(&c.ready, cudaEventDisableTiming));
        unsigned *buf = nullptr;
        void *tmp = nullptr;
        size_t tmp_bytes = 0;
        const int kmaxsq = 2 * (g.dims / 2 + 1) * (g.dims / 2 + 1);
        const int par_bit = bits_for((unsigned)kmaxsq);
        PYLB_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (unsigned *)nullptr, (unsigned *)nullptr,
                                                   (unsigned *)nullptr, (unsigned *)nullptr, nrows, 0, par_bit + 1, st));
        PYLB_CHECK(cudaMalloc(&c.tab, sizeof(Row2) * (size_t)nrows));
        if (g.F > 1) PYLB_CHECK(cudaMalloc(&c.cext, sizeof(float2) * (size_t)(g.F - 1) * nrows));
        ScratchGuard guard(st);
        PYLB_CHECK(cudaMallocAsync(&buf, sizeof(unsigned) * 4 * (size_t)nrows, st));
        guard.add(buf);
        PYLB_CHECK(cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 16, st));
        guard.add(tmp);
        unsigned *k_in = buf, *v_in = buf + nrows, *k_out = buf + 2 * (size_t)nrows, *v_out = buf + 3 * (size_t)nrows;
        const unsigned blocks = (unsigned)((nrows + 255) / 256);
        row_keys2_kernel<<<blocks, 256, 0, st>>>(k_in, v_in, nrows, g, bp, par_bit);
        PYLB_LAUNCH_CHECK();
        PYLB_CHECK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, v_in, v_out, nrows, 0, par_bit + 1, st));
        count_launch(3);
        row_table2_kernel<<<blocks, 256, 0, st>>>(k_out, v_out, c.tab, nrows, g, par_bit, c.cext);
        PYLB_LAUNCH_CHECK();
        PYLB_CHECK(cudaEventRecord(c.ready, st));
        c.key = key;
        hit = &c;
    } else {
        PYLB_CHECK(cudaStreamWaitEvent(st, hit->ready, 0));        // built on another stream, perhaps
    }
    Ring2Cache &c = *hit;
    Row2 *tab = c.tab;
    int *counter = nullptr;
    ScratchGuard guard(st);
    PYLB_CHECK(cudaMallocAsync(&counter, 16, st));
    guard.add(counter);

    const int nbins3 = isqrt_exact(3 * g.middle * g.middle) + 1;     // kmax + 1
    // the ring kernel's kz-private flush table t3[bin][value][kz] (22 MB at 1024^3, 87 MB at 2048^3)
    const int nvals = g.F == 1 ? R2_T3_VALS : 3 * (g.F + g.F * (g.F - 1) / 2) + 2;
    const int t3_kz = (kz_hi + 1 + 3) & ~3;
    const size_t t3_bytes = kz_hi >= 1 ? sizeof(double) * (size_t)nbins3 * nvals * t3_kz : 0;
    double *t3 = nullptr;
    if (t3_bytes) {
        PYLB_CHECK(cudaMallocAsync(&t3, t3_bytes, st));
        guard.add(t3);
        PYLB_CHECK(cudaMemsetAsync(t3, 0, t3_bytes, st));
    }
    // The self-conjugate columns first, on the same stream.  Measured and dropped: (i) a side stream next to the ring kernel
    // (the ring kernel slows down by more than the special kernel takes: 0.885 against 0.771 ms at 1024^3); (ii) clearing t3
    // with an extra row of CTAs of the special kernel instead of the memset (no gain at 2048^3, 13 us slower at 1024^3).
    // What bounds the special kernel is its 2 N^2 / G column gathers, one 32-byte sector per 8 KB row: ~35 G per second.
    const unsigned sblocks = (unsigned)((nrows + 255) / 256), scap = 4u * (unsigned)sm_count();
    const dim3 sgrid(sblocks < scap ? sblocks : scap, (g.middle > 0) ? 2 : 1);
    if (g.F == 1) special2_kernel<<<sgrid, 256, 0, st>>>(g, dk.p[0], tab, nrows, want_phase);
    else if (g.F == 2) special2x_kernel<2><<<sgrid, 256, 0, st>>>(g, dk, tab, nrows);
    else special2x_kernel<3><<<sgrid, 256, 0, st>>>(g, dk, tab, nrows);
    PYLB_LAUNCH_CHECK();
    if (kz_hi < 1) return 0;
    if (g.F == 1) return want_phase ? launch_ring2<true>(g, dk.p[0], tab, n0, (int)n1, kz_hi, counter, nbins3, t3, st)
                                    : launch_ring2<false>(g, dk.p[0], tab, n0, (int)n1, kz_hi, counter, nbins3, t3, st);
    if (g.F == 2) return launch_ring2x<2>(g, dk, tab, c.cext, nrows, n0, (int)n1, kz_hi, counter, nbins3, t3, st);
    return launch_ring2x<3>(g, dk, tab, c.cext, nrows, n0, (int)n1, kz_hi, counter, nbins3, t3, st);
}

static int run_ring(const BinGeom &g, const FieldPtrs &dk, int want_phase, int write_back, int precise, cudaStream_t st) {
    const int nrows = g.nx * g.ny;
    const int kz_hi = g.even ? g.middle - 1 : g.middle;  // columns 1..kz_hi carry no skip rule
    if (nrows == 0) return 0;
    if (g_allow_ring2 && g.F <= 3 && !write_back && !precise && !g.ximag && g.even && g.middle >= 2) {
        // one row table serves every field: they must agree in the parity of their base pointers (in units of 8 bytes)
        bool ok = ((uintptr_t)dk.p[0] & 7) == 0 && (g.F == 1 || !want_phase);
        for (int f = 1; f < g.F; f++) ok = ok && ((((uintptr_t)dk.p[f] ^ (uintptr_t)dk.p[0]) & 15) == 0);
        if (ok) return run_ring2(g, dk, want_phase, st);
    }

    // scratch: keys/vals (double-buffered for the radix sort), row table, cub temp
    unsigned *buf = nullptr;
    RowEnt *tab = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    const int kmaxsq = 2 * (g.dims / 2 + 1) * (g.dims / 2 + 1);
    const int end_bit = bits_for((unsigned)kmaxsq);
    PYLB_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (unsigned *)nullptr, (unsigned *)nullptr,
                                               (unsigned *)nullptr, (unsigned *)nullptr, nrows, 0, end_bit, st));
    ScratchGuard guard(st);
    PYLB_CHECK(cudaMallocAsync(&buf, sizeof(unsigned) * 4 * (size_t)nrows, st));
    guard.add(buf);
    PYLB_CHECK(cudaMallocAsync(&tab, sizeof(RowEnt) * (size_t)nrows, st));
    guard.add(tab);
    PYLB_CHECK(cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 16, st));
    guard.add(tmp);
    unsigned *k_in = buf, *v_in = buf + nrows, *k_out = buf + 2 * (size_t)nrows, *v_out = buf + 3 * (size_t)nrows;
    const unsigned blocks = (unsigned)((nrows + 255) / 256);
    row_keys_kernel<<<blocks, 256, 0, st>>>(k_in, v_in, nrows, g);
    PYLB_LAUNCH_CHECK();
    PYLB_CHECK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, v_in, v_out, nrows, 0, end_bit, st));
    count_launch(3);
    row_table_kernel<<<blocks, 256, 0, st>>>(k_out, v_out, tab, nrows, g);
    PYLB_LAUNCH_CHECK();

    // special columns (kz = 0 and, for even dims, kz = middle), then the bulk kz in [1, kz_hi]
    int rc = 1;
    switch (g.F) {
        case 1: rc = launch_special<1, RowEnt>(g, dk, tab, nrows, want_phase, write_back, st) ||
                     (kz_hi >= 1 && launch_ring_f<1>(g, dk, tab, nrows, kz_hi, want_phase, write_back, precise, st)); break;
        case 2: rc = launch_special<2, RowEnt>(g, dk, tab, nrows, want_phase, write_back, st) ||
                     (kz_hi >= 1 && launch_ring_f<2>(g, dk, tab, nrows, kz_hi, want_phase, write_back, precise, st)); break;
        case 3: rc = launch_special<3, RowEnt>(g, dk, tab, nrows, want_phase, write_back, st) ||
                     (kz_hi >= 1 && launch_ring_f<3>(g, dk, tab, nrows, kz_hi, want_phase, write_back, precise, st)); break;
        default: set_error("ring binning supports 1..3 fields, got %d", g.F);
    }
    return rc;
}

}  // namespace pylb

using namespace pylb;

extern "C" int pylb_pk_get_layout(int dims, int F, pylb_pk_layout *L) {
    PYLB_REQUIRE(L != nullptr && dims >= 2 && F >= 1 && F <= MAX_F, "pylb_pk_get_layout: need dims >= 2 and 1 <= F <= %d", MAX_F);
    const int middle = dims / 2;  // python-2 integer division in frequencies(), Pk_library.pyx:60
    L->dims = dims; L->F = F; L->X = F * (F - 1) / 2; L->middle = middle;
    L->kmax_par = middle;
    L->kmax_per = isqrt_exact(2 * middle * middle);      // int(sqrt(middle^2+middle^2)), :62
    L->kmax = isqrt_exact(3 * middle * middle);          // :63
    L->B2 = (int64_t)(L->kmax_par + 1) * (L->kmax_per + 1);
    int64_t o = 0;
    const int64_t n3 = L->kmax + 1, n1 = L->kmax_par + 1;
    L->o_k3d = o; o += n3;
    L->o_p3d = o; o += n3 * 3 * F;
    L->o_x3d = o; o += n3 * 3 * L->X;
    L->o_phase = o; o += n3;
    L->o_p1d = o; o += n1 * F;
    L->o_x1d = o; o += n1 * L->X;
    L->o_p2d = o; o += L->B2 * F;
    L->o_x2d = o; o += L->B2 * L->X;
    L->n_doubles = o;
    o = 0;
    L->o_n3d = o; o += n3;
    L->o_n1d = o; o += n1;
    L->o_n2d = o; o += L->B2;
    L->n_counts = o;
    return 0;
}

extern "C" int pylb_pk_bin(void *const *dk, int F, const pylb_kspace *ks, int axis, const int *mas_index,
                           int want_phase, int write_back, int algo, int accumulate, double *sums,
                           uint64_t *counts, void *stream) {
    PYLB_REQUIRE(dk && ks && mas_index && sums && counts, "pylb_pk_bin: NULL argument");
    PYLB_REQUIRE(F >= 1 && F <= MAX_F, "pylb_pk_bin: 1 <= F <= %d required, got %d", MAX_F, F);
    PYLB_REQUIRE(axis >= 0 && axis <= 2, "pylb_pk_bin: axis must be 0, 1 or 2");
    PYLB_REQUIRE(ks->dims >= 2 && ks->dims <= 32768, "pylb_pk_bin: dims out of range");
    PYLB_REQUIRE(ks->nx >= 0 && ks->ny >= 0 && ks->x0 >= 0 && ks->y0 >= 0 && ks->x0 + ks->nx <= ks->dims &&
                 ks->y0 + ks->ny <= ks->dims, "pylb_pk_bin: k-space window out of range");
    PYLB_REQUIRE((long long)ks->nx * ks->ny < (1ll << 31), "pylb_pk_bin: too many rows");
    keep_pool_memory();
    cudaStream_t st = (cudaStream_t)stream;
    pylb_pk_layout L;
    if (pylb_pk_get_layout(ks->dims, F, &L)) return 1;

    BinGeom g;
    g.dims = ks->dims; g.middle = L.middle; g.even = (ks->dims % 2 == 0);
    g.x0 = ks->x0; g.nx = ks->nx; g.y0 = ks->y0; g.ny = ks->ny;
    g.stride_x = ks->stride_x; g.stride_y = ks->stride_y;
    g.axis = axis; g.kmax_par1 = L.kmax_par + 1; g.F = F; g.X = L.X;
    g.o_k3d = L.o_k3d; g.o_p3d = L.o_p3d; g.o_x3d = L.o_x3d; g.o_phase = L.o_phase;
    g.o_p1d = L.o_p1d; g.o_x1d = L.o_x1d; g.o_p2d = L.o_p2d; g.o_x2d = L.o_x2d;
    g.o_n3d = L.o_n3d; g.o_n1d = L.o_n1d; g.o_n2d = L.o_n2d;
    g.sums = sums; g.counts = counts;
    FieldPtrs fp;
    for (int f = 0; f < MAX_F; f++) { fp.p[f] = nullptr; g.mas_idx[f] = 0; }
    for (int f = 0; f < F; f++) {
        PYLB_REQUIRE(dk[f] != nullptr, "pylb_pk_bin: field %d is NULL", f);
        PYLB_REQUIRE(mas_index[f] >= 0 && mas_index[f] <= 4, "pylb_pk_bin: MAS index %d out of range", mas_index[f]);
        fp.p[f] = (float2 *)dk[f];
        g.mas_idx[f] = mas_index[f];
    }
    timing_begin(PYLB_T_BIN, st);
    if (!accumulate) {
        PYLB_CHECK(cudaMemsetAsync(sums, 0, sizeof(double) * (size_t)L.n_doubles, st));
        PYLB_CHECK(cudaMemsetAsync(counts, 0, sizeof(uint64_t) * (size_t)L.n_counts, st));
    }
    double *tab = nullptr;
    const int ntab = F * (L.middle + 1);
    ScratchGuard guard(st);
    PYLB_CHECK(cudaMallocAsync(&tab, sizeof(double) * (size_t)ntab, st));
    guard.add(tab);
    g.mas_tab = tab;
    mas_table_kernel<<<(ntab + 127) / 128, 128, 0, st>>>(tab, L.middle, ks->dims, F, g);
    PYLB_LAUNCH_CHECK();

    g.ximag = (algo & PYLB_BIN_XIMAG) ? 1 : 0;
    const int precise = (algo & PYLB_BIN_PRECISE) ? 1 : 0;
    g_allow_bulk = (algo & PYLB_BIN_BULK) != 0;
    g_allow_ring2 = (algo & PYLB_BIN_RING1) == 0;
    algo &= ~(PYLB_BIN_PRECISE | PYLB_BIN_BULK | PYLB_BIN_RING1 | PYLB_BIN_XIMAG);
    if (algo == PYLB_BIN_AUTO) algo = (axis == 2 && F <= 3) ? PYLB_BIN_RING : PYLB_BIN_GENERIC;
    int rc;
    if (algo == PYLB_BIN_RING) {
        if (axis != 2) { set_error("pylb_pk_bin: the ring kernel needs axis=2 (transpose the field for other axes)"); rc = 1; }
        else rc = run_ring(g, fp, want_phase, write_back, precise, st);
    } else {
        timing_begin(PYLB_T_GENERIC, st);
        rc = launch_generic(g, fp, 0, 1, L.middle + 1, want_phase, write_back, st);
        timing_end(PYLB_T_GENERIC, st);
    }
    timing_end(PYLB_T_BIN, st);
    return rc;
}
