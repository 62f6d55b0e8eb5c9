/*
 * pylians_b200.h -- C ABI of the B200-native density-field -> power-spectrum path.
 *
 * Boundary contract
 * -----------------
 *  * plain C linkage, plain pointers and sizes, no torch / C++ types;
 *  * every `pylb_*` function returns 0 on success, non-zero on failure; the message is available
 *    from pylb_last_error() (thread-local);
 *  * unless a function name ends in `_host`, every data pointer is a DEVICE pointer and every call
 *    is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream);
 *  * the library never falls back to the CPU: without a CUDA device every compute entry fails.
 *
 * What each entry point replaces in the reference (paths relative to the Pylians tree):
 *  * NGP/CIC/TSC/PCS (host pointers)   = library/MAS_library/MAS_c.h:3-10, the reference's own C ABI,
 *                                        bound by MAS_c.pxd:1-10 and called from MAS_library.pyx:1136-1220;
 *  * pylb_ma                           = the Cython kernels behind MASL.MA, MAS_library.pyx:57-112
 *                                        (NGP :273, CIC :123, TSC :369, PCS :463 and their W variants,
 *                                        NGPW_d :338 / CICW_d :229 for the float64 grid);
 *  * pylb_fft_r2c                      = FFT3Dr_f, library/Pk_library/Pk_library.pyx:120-133 (pyfftw/FFTW);
 *  * pylb_pk_bin                       = the mode loops of class Pk :314-381 and class XPk :628-737;
 *  * pylb_pk_get_layout                = frequencies() :59-64 (bin counts);
 *  * pylb_overdensity                  = `delta /= mean; delta -= 1` done by every caller,
 *                                        e.g. library/Pk_library/Pk_snapshot.py:88,194;
 *  * pylb_pos_redshift_space           = library/redshift_space_library.pyx:29-43;
 *  * pylb_slab_pack / pylb_fft_*slab*  = new (the reference is single-process): the slab-decomposed FFT.
 */
#ifndef PYLIANS_B200_H
#define PYLIANS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PYLB_VERSION 100

/* mass-assignment scheme ids (deposit).  The Pk-side deconvolution exponent is MAS_function()'s
 * index (Pk_library.pyx:75-81): 0 none, 1 NGP, 2 CIC, 3 TSC, 4 PCS, i.e. id+1. */
enum { PYLB_NGP = 0, PYLB_CIC = 1, PYLB_TSC = 2, PYLB_PCS = 3 };

/* deposit algorithm selector for pylb_ma */
enum { PYLB_MA_AUTO = 0, PYLB_MA_DIRECT = 1, PYLB_MA_TILED = 2 };

/* binning algorithm selector for pylb_pk_bin.  OR in PYLB_BIN_PRECISE for the fp64 option: every
 * mode is squared and accumulated in float64 exactly like Pk_library.pyx:358-360 (the default forms
 * |delta_k|^2 in fp32 and accumulates in fp64 from the first group sum on; the generic kernel is
 * always fp64). */
enum { PYLB_BIN_AUTO = 0, PYLB_BIN_GENERIC = 1, PYLB_BIN_RING = 2, PYLB_BIN_PRECISE = 16,
       PYLB_BIN_BULK = 32 /* ring kernel: cp.async.bulk + mbarrier row loads by a producer warp instead of
                             per-thread cp.async (even dims only) */,
       PYLB_BIN_RING1 = 64 /* one field: use the one-kz-per-thread ring kernel instead of ring2 */,
       PYLB_BIN_XIMAG = 256 /* cross terms im_i*re_j - re_i*im_j (class XPk_imag :1131-1132) */ };

int pylb_version(void);
const char *pylb_last_error(void);
/* number of CUDA kernels this library has launched so far in this process (bench accounting) */
int64_t pylb_launch_count(void);

/* Per-kernel device timing for roofline reports.  When enabled, the dominant kernels / stages are
 * bracketed by CUDA events on their launching stream: which = 0 the binning ring kernel alone,
 * 1 the tiled-deposit tile kernel, 2 the direct deposit kernel, 3 the generic binning kernel,
 * 4 the WHOLE fused binning of one pylb_pk_bin call (bin zeroing, MAS table, self-conjugate columns,
 * ring kernel, finish pass), 5 the cuFFT execution of one transform call, 6 the sort stage of the tiled
 * deposit (histogram, scans, both counting-sort passes).
 * pylb_timing_collect synchronises the pending events, returns the summed milliseconds and the
 * number of brackets since the last collect, and resets the tally. */
enum { PYLB_T_RING = 0, PYLB_T_TILE = 1, PYLB_T_DIRECT = 2, PYLB_T_GENERIC = 3, PYLB_T_BIN = 4, PYLB_T_FFT = 5,
       PYLB_T_SORT = 6, PYLB_T_COUNT = 7 };
void pylb_timing_enable(int on);
int pylb_timing_collect(int which, double *total_ms, int *launches);

/* ---------------------------------------------------------------------------------------------
 * Mass assignment
 * ------------------------------------------------------------------------------------------- */

/* Reference-compatible host entry points: identical signatures to MAS_c.h:3-10.
 * pos: (particles, axes) row-major host floats; number: dims^axes row-major host floats, accumulated
 * in place; W may be NULL; `threads` is accepted and ignored.  No return value and no error channel,
 * exactly like the reference; on failure the message is left in pylb_last_error() and `number`
 * is untouched.  2-D calls reproduce MAS_c.c's single update per cell (n_max = 1). */
void NGP(float *pos, float *number, float *W, long particles, int dims, int axes, float BoxSize, int threads);
void CIC(float *pos, float *number, float *W, long particles, int dims, int axes, float BoxSize, int threads);
void TSC(float *pos, float *number, float *W, long particles, int dims, int axes, float BoxSize, int threads);
void PCS(float *pos, float *number, float *W, long particles, int dims, int axes, float BoxSize, int threads);

/* Device entry point behind MASL.MA.
 *   pos          device float32, element (i, a) at pos[i*pos_stride0 + a*pos_stride1]
 *   ndim         2 or 3 (= pos.shape[1] = number.ndim)
 *   grid         device, dims^ndim, C-contiguous, float32 (grid_f64 = 0) or float64 (grid_f64 = 1);
 *                ACCUMULATED INTO (+=), never cleared
 *   mas          PYLB_NGP..PYLB_PCS
 *   w            device float32 weights or NULL
 *   z_repeat     2-D only: multiply each update by this factor (the Cython 2-D path adds every
 *                contribution 2/3/4 times before MA() divides the array, MAS_library.pyx:84-110);
 *                pass 1 for the MAS_c.c behaviour
 *   algo         PYLB_MA_AUTO | PYLB_MA_DIRECT | PYLB_MA_TILED
 *   workspace    device scratch of at least pylb_ma_workspace_bytes(...) bytes (may be NULL when
 *                that function returns 0) */
size_t pylb_ma_workspace_bytes(int64_t np, int ndim, int dims, int mas, int has_w, int grid_f64, int algo);
int pylb_ma(const float *pos, int64_t np, int ndim, int64_t pos_stride0, int64_t pos_stride1,
            void *grid, int grid_f64, int dims, float box, int mas, const float *w, int z_repeat,
            int algo, void *workspace, size_t workspace_bytes, void *stream);

/* Deposit onto an x-window of the cube: `grid` holds planes x0 .. x0+xext-1 (mod dims), shape
 * (xext, dims, dims) float32.  Updates that fall outside the window are dropped.  `w` is read with
 * stride w_stride, so particles may come as packed (x,y,z,w) records (pos stride 4, w = pos+3, w_stride 4).
 * Used by the multi-GPU particle-exchange mode: every rank deposits only the particles it owns onto its
 * slab plus S-1 halo planes. */
size_t pylb_ma_window_workspace_bytes(int64_t np, int dims, int xext, int mas, int algo);
int pylb_ma_window(const float *pos, int64_t np, int64_t pos_stride0, int64_t pos_stride1, float *grid, int dims,
                   int x0, int xext, float box, int mas, const float *w, int64_t w_stride, int algo,
                   void *workspace, size_t workspace_bytes, void *stream);

/* Group particles by the x-slab (of dims/G planes) that owns their lowest touched x-plane.
 * out_xyzw: device float4[np]; offsets: device int[G+1] (offsets[g]..offsets[g+1] = slab g's particles). */
int pylb_partition_xslab(const float *pos, int64_t np, int64_t pos_stride0, int64_t pos_stride1, const float *w,
                         int64_t w_stride, int dims, float box, int mas, int G, void *out_xyzw, int *offsets,
                         void *stream);

/* dst[i] += src[i]  (adding received halo planes) */
int pylb_add_f32(float *dst, const float *src, int64_t n, void *stream);

/* Testing hook for the tiled deposit.  path = 100*kernel + sort.
 * sort:   0 automatic, 2 force the deep sort (second histogram sweep after pass 0), 3 / 4 deep sort with 1024 / 4 lo digits
 * kernel: 0 automatic, 1 lane-per-particle tile kernel for every scheme, 2 stencil-lane tile kernel where it exists (PCS);
 * negative = everything automatic. */
void pylb_ma_debug_path(int path);

/* grid[i] /= divisor  (the `number2 /= 2|3|4` of MAS_library.pyx:90-107; applies to the WHOLE array) */
int pylb_divide(float *grid, int64_t n, float divisor, void *stream);

/* Host (dims,dims,dims) float32 -> device padded in-place-FFT layout (dims,dims,2*(dims/2+1)), one
 * strided copy (cudaMemcpy2DAsync).  Lets Pk() transform a host field with a single device buffer. */
int pylb_h2d_padded(const float *host, float *dev, int dims, void *stream);
/* Same with an explicit device row pitch (floats per z-row), for pylb_fft_r2c_pitched. */
int pylb_h2d_pitched(const float *host, float *dev, int dims, int64_t dev_pitch, void *stream);

/* delta = grid/mean(grid) - 1 in place, mean accumulated in float64 (np.mean(dtype=float64) in the
 * callers).  `scratch` is a device double[2].  The two halves are exposed separately for slab-
 * decomposed grids: pylb_grid_sum adds sum(grid) into *sum (device double, caller zeroes it and may
 * all-reduce it), pylb_overdensity_apply maps grid -> grid*(n_total/(*sum)) - 1. */
int pylb_overdensity(float *grid, int64_t n, double *scratch, void *stream);
int pylb_grid_sum(const float *grid, int64_t n, double *sum, void *stream);
int pylb_overdensity_apply(float *grid, int64_t n, const double *sum, int64_t n_total, void *stream);

/* pos[:,axis] += vel[:,axis]*(1+z)/H with the reference's wrap rule; pos/vel (np,3) C-contiguous */
int pylb_pos_redshift_space(float *pos, const float *vel, int64_t np, float box, float hubble,
                            float redshift, int axis, void *stream);

/* ---------------------------------------------------------------------------------------------
 * FFT (cuFFT, plans cached inside the library per shape)
 * ------------------------------------------------------------------------------------------- */

/* 3-D real-to-complex, unnormalised, float32 (dims,dims,dims) -> complex64 (dims,dims,dims/2+1).
 * in == out is allowed when `in` uses the padded in-place layout (row length 2*(dims/2+1) floats).
 * work: device scratch of pylb_fft_r2c_work_bytes(dims) bytes. */
size_t pylb_fft_r2c_work_bytes(int dims, int inplace);
int pylb_fft_r2c(const float *in, void *out, int dims, int inplace, void *work, size_t work_bytes, void *stream);
/* Same transform with explicit z-row pitches: in_pitch floats per input row, out_pitch complex elements per
 * output row (>= dims/2+1).  in == out needs in_pitch == 2*out_pitch.  Pk() uses out_pitch = dims/2+2 when
 * dims/2+1 is odd, so that every k-space row starts on a 16-byte boundary (the binning kernel then reads
 * aligned element pairs from a single row table). */
size_t pylb_fft_r2c_pitched_work_bytes(int dims, int64_t in_pitch, int64_t out_pitch);
int pylb_fft_r2c_pitched(const float *in, int64_t in_pitch, void *out, int64_t out_pitch, int dims, void *work,
                         size_t work_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Siblings of Pk/XPk that share the FFT and the mode loop (library/Pk_library/Pk_library.pyx)
 * ------------------------------------------------------------------------------------------- */

/* IFFT3Dr_f :152-165 (backward c2r; `in` is destroyed), FFT2Dr_f :184-197, IFFT2Dr_f :216-229.
 * normalise != 0 scales the result by (float)(1/N) like pyfftw's FFTW.__call__ default (normalise_idft=True), which
 * is what the reference's IFFT*Dr_f return; normalise == 0 is the raw FFTW/cuFFT backward transform. */
int pylb_fft_c2r(void *in, float *out, int dims, int normalise, void *stream);
int pylb_fft2d_r2c(const float *in, void *out, int dims, void *stream);
int pylb_fft2d_c2r(void *in, float *out, int dims, int normalise, void *stream);

/* In-place pass over a (dims,dims,dims/2+1) complex64 field.  mode 0: the loop of correct_MAS :1770-1797
 * (independent modes times the MAS factor; the self-conjugate planes are left Hermitian-averaged, which is what
 * FFTW/pocketfft's c2r make of the reference's half-corrected planes).  mode 1: the first loop of Xi :2063-2083
 * (every stored mode becomes (|M delta_k|^2, 0)). */
int pylb_mas_correct(void *dk, int dims, int mas_index, int mode, void *stream);

/* Mode loop of Pk_theta :1283-1325.  sums: device double[3][kmax+1] = sum |k|, sum |theta|^2, Nmodes. */
int pylb_theta_bin(const void *vx, const void *vy, const void *vz, int dims, int mas_index, double *sums, void *stream);

/* Mode loops of Pk_plane :472-502 and XPk_plane :871-925 on (dims,dims/2+1) complex64 fields (d2 may be NULL).
 * sums: device double[5][kmax2d+1] = sum |k|, sum |d1|^2, sum |d2|^2, sum Re(d1 conj d2), Nmodes. */
int pylb_plane_bin(const void *d1, const void *d2, int dims, int mas1, int mas2, double *sums, void *stream);

/* Real-space loop of Xi :2097-2133 over the inverse transform.
 * sums: device double[5][kmax+1] = sum r, sum xi, sum xi*L2(mu), sum xi*L4(mu), Nmodes. */
int pylb_xi_bin(const float *xi, int dims, int axis, double *sums, void *stream);

/* Real-space axis swap feeding the FFT: out(i,j,k) = in(k,j,i) for axis 0, in(i,k,j) for axis 1, a copy
 * for axis 2.  `out` rows have out_pitch floats (dims for a dense cube, 2*(dims/2+1) for the padded
 * in-place R2C layout).  The power spectrum with the line of sight along `axis` equals the spectrum of
 * the swapped field with the line of sight along z (the binning ring kernel's fast direction). */
int pylb_swap_axes(const float *in, float *out, int dims, int axis, int64_t out_pitch, void *stream);

/* Slab-decomposed pieces (multi-GPU).  Real slab [nx_local][dims][dims] -> batched 2-D R2C over
 * (y,z) -> complex [nx_local][dims][pitch]; then, after the all-to-all, 1-D C2C along x on the
 * transposed layout [dims][ny_local][pitch] (in place).  `pitch` = complex elements per kz-row, >= dims/2+1; the slab
 * engine uses the next even number, so that every row starts on a 16-byte boundary for the binning kernel (the padding
 * column travels with its row and is never read as a mode). */
size_t pylb_fft_slab_yz_work_bytes(int dims, int nx_local, int64_t out_pitch);
int pylb_fft_slab_yz(const float *in, void *out, int dims, int nx_local, int64_t out_pitch, void *work, size_t work_bytes,
                     void *stream);
size_t pylb_fft_slab_x_work_bytes(int dims, int ny_local, int64_t pitch);
int pylb_fft_slab_x(void *data, int dims, int ny_local, int64_t pitch, void *work, size_t work_bytes, void *stream);

/* Slab transpose pack: src complex [nx_local][dims][pitch] -> dst [G][nx_local][dims/G][pitch], i.e. the
 * send buffer of the all-to-all, block g going to rank g. */
int pylb_slab_pack(const void *src, void *dst, int dims, int nx_local, int G, int64_t pitch, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Power-spectrum binning
 * ------------------------------------------------------------------------------------------- */

/* Bin geometry for a dims^3 grid and F fields (X = F(F-1)/2 cross pairs in the reference's order
 * (0,1),(0,2),..,(1,2),..).  All accumulators live in two device buffers the caller allocates and
 * the library zeroes: `sums` (double[n_doubles]) and `counts` (uint64[n_counts]).
 * Raw sums only -- units, (2l+1), /Nmodes and the DC-bin handling of Pk_library.pyx:387-421 are
 * host bookkeeping on a few KB done by the caller.
 *   k3d   [kmax+1]            sum of |k| (in units of kF)
 *   p3d   [kmax+1][3][F]      sum of |delta_k|^2 * {1, L2(mu), L4(mu)}
 *   x3d   [kmax+1][3][X]
 *   phase [kmax+1]            sum of atan2(re, |delta_k|)^2, field 0 (Pk only)
 *   p1d   [kmax_par+1][F]     x1d [kmax_par+1][X]       (modes with |k| <= dims/2)
 *   p2d   [B2][F]             x2d [B2][X]               index (kmax_par+1)*k_per + |k_par|
 *   n3d [kmax+1], n1d [kmax_par+1], n2d [B2]   integer mode counts (bit-exact contract) */
typedef struct {
    int dims, F, X, middle, kmax_par, kmax_per, kmax;
    int64_t B2;
    int64_t o_k3d, o_p3d, o_x3d, o_phase, o_p1d, o_x1d, o_p2d, o_x2d, n_doubles;
    int64_t o_n3d, o_n1d, o_n2d, n_counts;
} pylb_pk_layout;
int pylb_pk_get_layout(int dims, int F, pylb_pk_layout *out);

/* Which part of k-space a device buffer holds.  Element (kxx, kyy, kzz) of field f lives at
 * dk[f][(kxx-x0)*stride_x + (kyy-y0)*stride_y + kzz] (complex64 elements), kzz in [0, dims/2].
 * Single GPU: x0=y0=0, nx=ny=dims, stride_y=dims/2+1, stride_x=dims*stride_y.
 * Slab FFT output on rank r of G: x0=0, nx=dims, y0=r*dims/G, ny=dims/G,
 * stride_y=dims/2+1, stride_x=ny*stride_y. */
typedef struct {
    int dims, x0, nx, y0, ny;
    int64_t stride_x, stride_y;
} pylb_kspace;

/* Deconvolve (MAS window), square, bin.  dk: HOST array of F device pointers.
 *   mas_index[f]  deconvolution exponent of field f (0..4)
 *   axis          line of sight (0,1,2)
 *   want_phase    accumulate `phase` (class Pk does; XPk does not)
 *   write_back    store the deconvolved modes back into dk (keep_deltak; the reference always
 *                 overwrites its private copy, Pk_library.pyx:355)
 *   algo          PYLB_BIN_AUTO | PYLB_BIN_GENERIC | PYLB_BIN_RING
 * sums/counts are zeroed by this call unless accumulate != 0. */
int pylb_pk_bin(void *const *dk, int F, const pylb_kspace *ks, int axis, const int *mas_index,
                int want_phase, int write_back, int algo, int accumulate, double *sums,
                uint64_t *counts, void *stream);

/* After pylb_pk_bin (and, multi-GPU, after the all-reduce of sums/counts): the 2-D table's normalisation
 * Pk2D = sum * (fact / Nmodes2D) (Pk_library.pyx:397-406, 771-786) for every field and pair, in place, and every
 * mode count rewritten in place as a float64 (the reference's Nmodes arrays are float64), so that the host reads
 * the finished tables with one copy.  fact = (BoxSize/dims^2)^3.  Bins with no mode become inf/nan; the caller
 * checks min(Nmodes2D) first (ZeroDivisionError in the reference). */
int pylb_pk_finish_tables(double *sums, uint64_t *counts, int dims, int F, double fact, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Callers and consumers either side of the path; csrc/consumers.cu
 * ------------------------------------------------------------------------------------------- */
/* x[i] *= mul in fp32: the `data *= math.sqrt(time)` of the snapshot reader (library/readsnap.py:376) */
int pylb_scale_f32(float *x, int64_t n, float mul, void *stream);
/* delta /= mean; delta -= 1.0 with the caller's mean, e.g. Np/dims^3 (Pk_library/Pk_snapshot.py:84-88, 191-194) */
int pylb_overdensity_mean(float *grid, int64_t n, float mean, void *stream);
/* dst[i] += a*src[i]: delta_tot += Omega*delta (Pk_snapshot.py:248-254) */
int pylb_axpy_f32(float *dst, const float *src, float a, int64_t n, void *stream);

/* smoothing_library.FT_filter (smoothing_library/smoothing_library.pyx:19-83) up to its FFT: fills the (dims^3)
 * float32 grid with the Top-Hat (kind 0) or Gaussian (kind 1) filter of squared radius R2 (grid cells, float32),
 * normalised to unit sum.  scratch: device double[1]. */
int pylb_filter_real(float *field, int dims, float R2, int kind, double *scratch, void *stream);
/* a[i] *= b[i], complex64: the mode loop of field_smoothing (:104-108) */
int pylb_cmul_c64(void *a, const void *b, int64_t n, void *stream);

/* void_library.gaussian_smoothing (void_library/void_library.pyx:45-80): dk *= top-hat window of kR = prefact*|k|,
 * prefact = (float)(R*2*pi/BoxSize); complex64 (dims,dims,dims/2+1), in place, DC mode untouched */
int pylb_tophat_k(void *dk, int dims, float prefact, void *stream);

/* bispectrum_library.Bk (Pk_library/bispectrum_library.pyx:88-130): out_d = MAS-deconvolved delta_k on the shell
 * kmin <= |k| < kmax (units of kF) and 0 elsewhere; out_i = 1 on the shell, 0 elsewhere.  complex64
 * (dims,dims,dims/2+1) buffers; dk is not modified. */
int pylb_bk_shell(const void *dk, void *out_d, void *out_i, int dims, int mas_index, double kmin, double kmax,
                  void *stream);
/* *out = sum_i a[i]*b[i][*c[i]] (c may be NULL): fp32 products summed in float64 (:141-146, :187-193).  out: device double */
int pylb_prod_sum(const float *a, const float *b, const float *c, int64_t n, double *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PYLIANS_B200_H */
