/*
 * pyl_oracle.c -- CPU restatement of the Pylians density-field -> power-spectrum hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA product path in
 * pylians_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it.  It is never on the product path.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks every function below against
 * the unmodified reference compiled from /root/reference (oracle/_ref, see oracle/build_ref.py),
 * and tests/golden/ holds reference-generated vectors (tests/golden/make_golden.py) so the pin
 * also holds on machines without /root/reference.
 *
 * Every function cites the reference lines it restates (paths relative to /root/reference).
 * The code is a serial, single-threaded restatement -- deliberately simple.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------------------------------------
 * Mass assignment.  library/MAS_library/MAS_library.pyx
 *   NGP :273-292 / NGPW :305-325      CIC :123-166 / CICW :179-216
 *   TSC :369-404 / TSCW :417-452      PCS :463-497 / PCSW :510-545
 * Shared conventions (all kernels):
 *   inv_cell_size = (float)(dims / BoxSize)  -- `cdef float inv_cell_size = dims/BoxSize` with
 *       BoxSize a C float: int/float division in fp32 (:135,:282,:380,:473);
 *   dist = pos * inv_cell_size in fp32 (:152,:290,:392,:485);
 *   2-D inputs use index 0 / weight 1 on the third axis (:138-139,:285,:383-385,:476-478);
 *   `number` is accumulated into (+=), never cleared.
 * pos is addressed with element strides so C- and Fortran-ordered inputs both work, like the
 * reference's strided memoryviews.  grid is C-contiguous dims^ndim.
 * mas: 0 NGP, 1 CIC, 2 TSC, 3 PCS.   W may be NULL.
 * ------------------------------------------------------------------------------------------- */

static inline int wrap_pos(int i, int dims) { /* (i + dims) % dims with C remainder, :291,:395,:488 */
    return (i + dims) % dims;
}

#define ORC_MA_BODY(GRID_T)                                                                        \
    const float inv = (float)dims / box;                                                           \
    const long s0 = (ndim == 3) ? (long)dims * dims : dims, s1 = (ndim == 3) ? dims : 1,            \
               s2 = (ndim == 3) ? 1 : 0;                                                            \
    for (long i = 0; i < np; i++) {                                                                \
        int idx[3][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};                                  \
        float C[3][4] = {{1, 1, 1, 1}, {1, 1, 1, 1}, {1, 1, 1, 1}};                                  \
        /* 2-D: the third axis keeps weight 1 / index 0 but is still looped over its full support \
           (2,3,4 updates of the same cell), which MA() then divides out  :84-110 */                \
        const int sup = (mas == 0) ? 1 : mas + 1;                                                  \
        int n_ax[3] = {sup, sup, sup};                                                             \
        for (int ax = 0; ax < ndim; ax++) {                                                        \
            const float dist = pos[i * ps0 + ax * ps1] * inv;                                      \
            if (mas == 0) { /* NGP: <int>(dist + 0.5) in double, then (i+dims)%dims  :290-291 */    \
                int c = (int)((double)dist + 0.5);                                                 \
                idx[ax][0] = wrap_pos(c, dims);                                                    \
                n_ax[ax] = 1;                                                                      \
            } else if (mas == 1) { /* CIC :152-157 (no +dims: valid domain 0<=pos<=BoxSize) */     \
                const int id = (int)dist;                                                          \
                const float u = dist - (float)id;                                                  \
                const float d = (float)(1.0 - (double)u);                                          \
                idx[ax][0] = id % dims;                                                            \
                idx[ax][1] = (idx[ax][0] + 1) % dims;                                              \
                C[ax][0] = d;                                                                      \
                C[ax][1] = u;                                                                      \
                n_ax[ax] = 2;                                                                      \
            } else if (mas == 2) { /* TSC :392-399 */                                              \
                const int m = (int)floor((double)dist - 1.5);                                      \
                for (int j = 0; j < 3; j++) {                                                      \
                    idx[ax][j] = wrap_pos(m + j + 1, dims);                                        \
                    const float diff = (float)fabs((double)((float)(m + j + 1) - dist));           \
                    if (diff < 0.5)                                                                \
                        C[ax][j] = (float)(0.75 - (double)(diff * diff));                          \
                    else if (diff < 1.5)                                                           \
                        C[ax][j] = (float)(0.5 * (1.5 - diff) * (1.5 - diff));                     \
                    else                                                                           \
                        C[ax][j] = 0.0f;                                                           \
                }                                                                                  \
                n_ax[ax] = 3;                                                                      \
            } else { /* PCS :485-492 */                                                            \
                const int m = (int)floor((double)dist - 2.0);                                      \
                for (int j = 0; j < 4; j++) {                                                      \
                    idx[ax][j] = wrap_pos(m + j + 1, dims);                                        \
                    const float diff = (float)fabs((double)((float)(m + j + 1) - dist));           \
                    if (diff < 1.0)                                                                \
                        C[ax][j] = (float)((4.0 - 6.0 * diff * diff + 3.0 * diff * diff * diff) / 6.0); \
                    else if (diff < 2.0)                                                           \
                        C[ax][j] = (float)((2.0 - diff) * (2.0 - diff) * (2.0 - diff) / 6.0);      \
                    else                                                                           \
                        C[ax][j] = 0.0f;                                                           \
                }                                                                                  \
                n_ax[ax] = 4;                                                                      \
            }                                                                                      \
        }                                                                                          \
        /* 1 / 8 / 27 / 64 updates; weight product left to right in fp32, then *W  :159-166,       \
           :209-216, :404, :452, :497, :545.  NGP adds 1.0 or W[i]  :292,:325. */                   \
        for (int l = 0; l < n_ax[0]; l++)                                                          \
            for (int m2 = 0; m2 < n_ax[1]; m2++)                                                   \
                for (int n = 0; n < n_ax[2]; n++) {                                                \
                    float w = C[0][l] * C[1][m2] * C[2][n];                                        \
                    if (W) w = w * W[i];                                                           \
                    grid[idx[0][l] * s0 + idx[1][m2] * s1 + idx[2][n] * s2] += (GRID_T)w;          \
                }                                                                                  \
    }

/* float32 grid: the kernels reached from MA(), MAS_library.pyx:57-112 */
void orc_ma_f32(const float *pos, long np, int ndim, long ps0, long ps1, float *grid, int dims,
                float box, int mas, const float *W) {
    ORC_MA_BODY(float)
}

/* float64 grid: NGPW_d :338-358 and CICW_d :229-266 (the reference's fp64-accumulation precedent;
 * generalised here to all four schemes so the CUDA fp64-accumulate option has a checker). */
void orc_ma_f64(const float *pos, long np, int ndim, long ps0, long ps1, double *grid, int dims,
                float box, int mas, const float *W) {
    ORC_MA_BODY(double)
}

/* ---------------------------------------------------------------------------------------------
 * Redshift-space displacement.  library/redshift_space_library.pyx:29-43
 *   factor = (float)((1+z)/H);  pos[:,axis] += vel[:,axis]*factor;  wrap with (p+Box) fmod Box
 *   when p > Box or p < 0.
 * ------------------------------------------------------------------------------------------- */
void orc_pos_redshift_space(float *pos, const float *vel, long np, float box, float hubble,
                            float redshift, int axis) {
    const float factor = (float)((1.0 + (double)redshift) / (double)hubble);
    for (long i = 0; i < np; i++) {
        float p = pos[3 * i + axis] + vel[3 * i + axis] * factor;
        if (p > box || p < 0.0f) p = fmodf(p + box, box); /* cdivision(True): C fmod  :42-43 */
        pos[3 * i + axis] = p;
    }
}

/* ---------------------------------------------------------------------------------------------
 * Power-spectrum mode loop.  library/Pk_library/Pk_library.pyx
 *   class Pk  :266-425 (loop :314-381)      class XPk :534-798 (loop :628-737)
 * One routine serves both: F fields, X = F(F-1)/2 pairs in order (0,1),(0,2),..,(1,2).. (:718-736).
 * dk[f] points to field f's complex64 half-spectrum (dims,dims,dims/2+1), interleaved re,im,
 * C order, and is deconvolved IN PLACE like the reference (:355, :695-696).
 * Raw sums are returned (units / normalisation are applied by the Python wrapper,
 * oracle/pylians_oracle.py, following :387-421 and :740-796):
 *   k3d[kmax+1], n3d[kmax+1], p3d[(kmax+1)*3*F] (index [k][ell][f]), x3d[(kmax+1)*3*X],
 *   phase[kmax+1] (field 0 only; Pk :361,:380),
 *   k1d[kpar+1], n1d[kpar+1], p1d[(kpar+1)*F], x1d[(kpar+1)*X],
 *   n2d[B2], p2d[B2*F], x2d[B2*X]   with B2=(kmax_par+1)*(kmax_per+1), index (kmax_par+1)*k_per+k_par.
 * `middle = dims/2` is python-2 integer division (:285,:559), kept for odd dims.
 * ------------------------------------------------------------------------------------------- */
static double mas_correction(double x, int p) { /* :86-87 */
    return (x == 0.0) ? 1.0 : pow(x / sin(x), (double)p);
}

/* class XPk_imag (Pk_library.pyx:959-1225) is class XPk with the cross term im_i*re_j - re_i*im_j (:1131-1132) */
static int g_cross_imag = 0;
void orc_set_cross_imag(int on) { g_cross_imag = on; }

void orc_pk_loop(float **dk, int F, int dims, int axis, const int *mas_index, int kmax_par,
                 int kmax_per, int kmax, double *k3d, double *n3d, double *p3d, double *x3d,
                 double *phase, double *k1d, double *n1d, double *p1d, double *x1d, double *n2d,
                 double *p2d, double *x2d) {
    const int middle = dims / 2;
    const int X = F * (F - 1) / 2;
    const long nz = middle + 1;
    const double prefact = M_PI / dims; /* :313 */
    double *cx = (double *)malloc(sizeof(double) * F), *cy = (double *)malloc(sizeof(double) * F),
           *cz = (double *)malloc(sizeof(double) * F);
    double *re = (double *)malloc(sizeof(double) * F), *im = (double *)malloc(sizeof(double) * F);
    (void)kmax; (void)kmax_per;
    for (int kxx = 0; kxx < dims; kxx++) {
        const int kx = (kxx > middle) ? kxx - dims : kxx; /* :315 */
        for (int f = 0; f < F; f++) cx[f] = mas_correction(prefact * kx, mas_index[f]);
        for (int kyy = 0; kyy < dims; kyy++) {
            const int ky = (kyy > middle) ? kyy - dims : kyy; /* :319 */
            for (int f = 0; f < F; f++) cy[f] = mas_correction(prefact * ky, mas_index[f]);
            for (int kzz = 0; kzz < middle + 1; kzz++) {
                const int kz = (kzz > middle) ? kzz - dims : kzz; /* :323 */
                for (int f = 0; f < F; f++) cz[f] = mas_correction(prefact * kz, mas_index[f]);

                /* keep one of each conjugate pair on the self-conjugate planes  :326-330 */
                if (kz == 0 || (kz == middle && dims % 2 == 0)) {
                    if (kx < 0) continue;
                    else if (kx == 0 || (kx == middle && dims % 2 == 0)) {
                        if (ky < 0) continue;
                    }
                }
                const double k = sqrt((double)(kx * kx + ky * ky + kz * kz)); /* :334 */
                const int k_index = (int)k;                                   /* :335 */
                int k_par, k_per;                                             /* :338-343 */
                if (axis == 0)      { k_par = kx; k_per = (int)sqrt((double)(ky * ky + kz * kz)); }
                else if (axis == 1) { k_par = ky; k_per = (int)sqrt((double)(kx * kx + kz * kz)); }
                else                { k_par = kz; k_per = (int)sqrt((double)(kx * kx + ky * ky)); }
                const double mu = (k == 0.0) ? 0.0 : k_par / k; /* :346-347 */
                const double mu2 = mu * mu;
                const double val1 = (3.0 * mu2 - 1.0) / 2.0;                       /* :378, :666 */
                const double val2 = (35.0 * mu2 * mu2 - 30.0 * mu2 + 3.0) / 8.0;   /* :379, :667 */
                if (k_par < 0) k_par = -k_par; /* :351 */
                const int in1d = (k <= middle); /* :364 */
                const long i2 = (long)(kmax_par + 1) * k_per + k_par; /* :371 */

                if (in1d) { k1d[k_par] += k_par; n1d[k_par] += 1.0; } /* :365,:367 */
                n2d[i2] += 1.0;                                       /* :373 */
                k3d[k_index] += k;                                    /* :376 */
                n3d[k_index] += 1.0;                                  /* :381 */

                const long off = 2 * (((long)kxx * dims + kyy) * nz + kzz);
                for (int f = 0; f < F; f++) {
                    /* MAS_factor: double product cast to float; complex64 *= float  :354-355 */
                    const float mf = (float)(cx[f] * cy[f] * cz[f]);
                    float *z = dk[f] + off;
                    z[0] = z[0] * mf;
                    z[1] = z[1] * mf;
                    re[f] = z[0];
                    im[f] = z[1];
                    const double d2 = re[f] * re[f] + im[f] * im[f]; /* :358-360 */
                    if (f == 0 && phase) {
                        const double ph = atan2(re[0], sqrt(d2)); /* :361 (sic: real vs modulus) */
                        phase[k_index] += ph * ph;                /* :380 */
                    }
                    if (in1d) p1d[(long)k_par * F + f] += d2;     /* :366 */
                    p2d[i2 * F + f] += d2;                        /* :372 */
                    p3d[((long)k_index * 3 + 0) * F + f] += d2;   /* :377 */
                    p3d[((long)k_index * 3 + 1) * F + f] += d2 * val1;
                    p3d[((long)k_index * 3 + 2) * F + f] += d2 * val2;
                }
                int ix = 0;
                for (int i = 0; i < F; i++)
                    for (int j = i + 1; j < F; j++) { /* :719-736 */
                        const double dx = g_cross_imag ? im[i] * re[j] - re[i] * im[j] : re[i] * re[j] + im[i] * im[j];
                        if (in1d) x1d[(long)k_par * X + ix] += dx;
                        x2d[i2 * X + ix] += dx;
                        x3d[((long)k_index * 3 + 0) * X + ix] += dx;
                        x3d[((long)k_index * 3 + 1) * X + ix] += dx * val1;
                        x3d[((long)k_index * 3 + 2) * X + ix] += dx * val2;
                        ix++;
                    }
            }
        }
    }
    free(cx); free(cy); free(cz); free(re); free(im);
}
