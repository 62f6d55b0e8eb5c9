// Siblings of Pk/XPk that share the FFT and the mode loop: the 2-D spectra Pk_plane / XPk_plane,
// Pk_theta, correct_MAS and the correlation function Xi of library/Pk_library/Pk_library.pyx.  Same mode set,
// same MAS deconvolution, a different per-mode payload.  One thread per mode (or per cell for Xi); neighbouring
// lanes share their |k| bin, so bins are summed per warp (warp_reduce_by_key) before one red.global per value.
#include <cufft.h>

#include <map>
#include <tuple>

#include "common.cuh"

namespace pylb {

// ---- secondary transforms: cached plans with cuFFT-owned work areas
enum { SK_C2R_3D = 0, SK_R2C_2D = 1, SK_C2R_2D = 2 };
static std::map<std::tuple<int, int, int>, cufftHandle> g_sib_plans;

static int sib_plan(int kind, int dims, cufftHandle *out) {
    int dev = 0;
    PYLB_CHECK(cudaGetDevice(&dev));
    auto key = std::make_tuple(dev, kind, dims);
    auto it = g_sib_plans.find(key);
    if (it != g_sib_plans.end()) { *out = it->second; return 0; }
    cufftHandle h;
    cufftResult r;
    if (kind == SK_C2R_3D) r = cufftPlan3d(&h, dims, dims, dims, CUFFT_C2R);
    else r = cufftPlan2d(&h, dims, dims, kind == SK_R2C_2D ? CUFFT_R2C : CUFFT_C2R);
    if (r != CUFFT_SUCCESS) { set_error("cuFFT plan (kind %d, dims %d) failed: %d", kind, dims, (int)r); return 1; }
    g_sib_plans[key] = h;
    *out = h;
    return 0;
}

// (x/sin x)^p at |k| = 0..middle, Pk_library.pyx:86-87
__global__ void sib_mas_table_kernel(double *tab, int middle, int dims, int p) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > middle) return;
    double v = 1.0;
    if (k != 0 && p != 0) {
        const double x = (M_PI / (double)dims) * (double)k;
        const double q = x / sin(x);
        v = q;
        for (int i = 1; i < p; i++) v *= q;
    }
    tab[k] = v;
}

struct SibGeom {
    int dims, middle, even, kmax1;   // kmax1 = number of bins (kmax + 1)
};

__device__ __forceinline__ bool skip_mode_3d(int kx, int ky, int kz, const SibGeom &g) {   // :326-330
    if (kz == 0 || (kz == g.middle && g.even)) {
        if (kx < 0) return true;
        if ((kx == 0 || (kx == g.middle && g.even)) && ky < 0) return true;
    }
    return false;
}

// ------------------------------------------------------------------------------------------------
// correct_MAS (:1749-1806) and the first loop of Xi (:2063-2083)
//   MODE 0 (correct_MAS): independent modes are multiplied by the MAS factor, the dependent ones on the planes
//     kz = 0 and kz = N/2 are left alone (:1786-1794), and the backward transform then sees planes that are no longer
//     Hermitian.  FFTW/pocketfft c2r take the real part after the complex passes, i.e. they transform the Hermitian
//     average B(k) = (A'(k) + conj(A'(-k)))/2.  cuFFT makes no such promise, so B is written to both members of
//     each pair here and every library gives the same field.
//   MODE 1 (Xi): every stored mode becomes (|M delta_k|^2, 0), formed in fp32 like the reference's `float real, imag`.
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256)
mas_correct_kernel(float2 *__restrict__ dk, SibGeom g, const double *__restrict__ tab) {
    const int nz = g.middle + 1;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)g.dims * g.dims * nz;
    if (idx >= total) return;
    const int kzz = (int)(idx % nz);
    const int iy = (int)((idx / nz) % g.dims), ix = (int)(idx / ((long long)nz * g.dims));
    const int kx = wavenumber(ix, g.dims, g.middle), ky = wavenumber(iy, g.dims, g.middle);
    const float mf = (float)(tab[kx < 0 ? -kx : kx] * tab[ky < 0 ? -ky : ky] * tab[kzz]);
    const float2 z = dk[idx];
    const float re = __fmul_rn(z.x, mf), im = __fmul_rn(z.y, mf);
    if (MODE == 1) {
        dk[idx] = make_float2(re * re + im * im, 0.0f);
        return;
    }
    const bool plane = kzz == 0 || (kzz == g.middle && g.even);
    if (!plane) { dk[idx] = make_float2(re, im); return; }
    if (skip_mode_3d(kx, ky, kzz, g)) return;            // written by its partner
    const int jx = (g.dims - ix) % g.dims, jy = (g.dims - iy) % g.dims;
    if (jx == ix && jy == iy) { dk[idx] = make_float2(re, im); return; }   // self-conjugate
    const long long pidx = ((long long)jx * g.dims + jy) * nz + kzz;
    const float2 b = dk[pidx];                            // untouched by the reference
    const float2 B = make_float2(0.5f * (re + b.x), 0.5f * (im - b.y));
    dk[idx] = B;
    dk[pidx] = make_float2(B.x, -B.y);
}

// ------------------------------------------------------------------------------------------------
// Pk_theta mode loop (:1283-1325): theta(k) = i k.V(k); sums[0..2][bin] = sum |k|, sum |theta|^2, Nmodes
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
theta_bin_kernel(const float2 *__restrict__ vx, const float2 *__restrict__ vy, const float2 *__restrict__ vz, SibGeom g,
                 const double *__restrict__ tab, double *__restrict__ sums) {
    const int nz = g.middle + 1;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)g.dims * g.dims * nz;
    bool valid = idx < total;
    const long long id = valid ? idx : 0;
    const int kz = (int)(id % nz);
    const int iy = (int)((id / nz) % g.dims), ix = (int)(id / ((long long)nz * g.dims));
    const int kx = wavenumber(ix, g.dims, g.middle), ky = wavenumber(iy, g.dims, g.middle);
    if (skip_mode_3d(kx, ky, kz, g)) valid = false;
    const int n = kx * kx + ky * ky + kz * kz;
    const int bin = isqrt_exact(n);
    double v[3] = {0, 0, 0};
    if (valid) {
        const float mf = (float)(tab[kx < 0 ? -kx : kx] * tab[ky < 0 ? -ky : ky] * tab[kz]);
        const float2 a = vx[id], b = vy[id], c = vz[id];
        const float ar = __fmul_rn(a.x, mf), ai = __fmul_rn(a.y, mf), br = __fmul_rn(b.x, mf), bi = __fmul_rn(b.y, mf),
                    cr = __fmul_rn(c.x, mf), ci = __fmul_rn(c.y, mf);
        // int * float is float arithmetic in the reference's C (:1308-1314)
        const float fx = (float)kx, fy = (float)ky, fz = (float)kz;
        const double real = (double)(-(fx * ai + fy * bi + fz * ci));
        const double imag = (double)(fx * ar + fy * br + fz * cr);
        v[0] = sqrt((double)n);
        v[1] = real * real + imag * imag;
        v[2] = 1.0;
    }
    warp_reduce_by_key<3>(bin, valid, v, [&](int b, const double (&s)[3]) {
        red_add(sums + b, s[0]);
        red_add(sums + g.kmax1 + b, s[1]);
        red_add(sums + 2 * g.kmax1 + b, s[2]);
    });
}

// ------------------------------------------------------------------------------------------------
// Pk_plane (:472-502) and XPk_plane (:871-925): 2-D field(s), modes (kx, ky >= 0);
// sums[0..4][bin] = sum |k|, sum |d1|^2, sum |d2|^2, sum Re(d1 conj d2), Nmodes
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
plane_bin_kernel(const float2 *__restrict__ d1, const float2 *__restrict__ d2, SibGeom g, const double *__restrict__ tab1,
                 const double *__restrict__ tab2, double *__restrict__ sums) {
    const int ny = g.middle + 1;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = idx < g.dims * ny;
    const int id = valid ? idx : 0;
    const int ky = id % ny, ix = id / ny;
    const int kx = wavenumber(ix, g.dims, g.middle);
    if ((ky == 0 || (ky == g.middle && g.even)) && kx < 0) valid = false;      // :484-485
    const int n = kx * kx + ky * ky;
    const int bin = isqrt_exact(n);
    double v[5] = {0, 0, 0, 0, 0};
    if (valid) {
        const int ax = kx < 0 ? -kx : kx;
        const float m1 = (float)(tab1[ax] * tab1[ky]);
        const float2 a = d1[id];
        const double ar = (double)__fmul_rn(a.x, m1), ai = (double)__fmul_rn(a.y, m1);
        v[0] = sqrt((double)n);
        v[1] = ar * ar + ai * ai;
        v[4] = 1.0;
        if (d2) {
            const float m2 = (float)(tab2[ax] * tab2[ky]);
            const float2 b = d2[id];
            const double br = (double)__fmul_rn(b.x, m2), bi = (double)__fmul_rn(b.y, m2);
            v[2] = br * br + bi * bi;
            v[3] = ar * br + ai * bi;
        }
    }
    warp_reduce_by_key<5>(bin, valid, v, [&](int b, const double (&s)[5]) {
#pragma unroll
        for (int q = 0; q < 5; q++)
            if (q == 0 || q == 1 || q == 4 || d2) red_add(sums + q * g.kmax1 + b, s[q]);
    });
}

// ------------------------------------------------------------------------------------------------
// Xi real-space loop (:2097-2133): every cell of the inverse transform, signed separations;
// sums[0..4][bin] = sum r, sum xi, sum xi L2(mu), sum xi L4(mu), Nmodes
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
xi_bin_kernel(const float *__restrict__ xi, SibGeom g, int axis, double *__restrict__ sums) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)g.dims * g.dims * g.dims;
    const bool valid = idx < total;
    const long long id = valid ? idx : 0;
    const int iz = (int)(id % g.dims), iy = (int)((id / g.dims) % g.dims), ix = (int)(id / ((long long)g.dims * g.dims));
    const int kx = wavenumber(ix, g.dims, g.middle), ky = wavenumber(iy, g.dims, g.middle), kz = wavenumber(iz, g.dims, g.middle);
    const int n = kx * kx + ky * ky + kz * kz;
    const int bin = isqrt_exact(n);
    double v[5] = {0, 0, 0, 0, 0};
    if (valid) {
        const double k = sqrt((double)n);
        const int k_par = axis == 0 ? kx : (axis == 1 ? ky : kz);
        const double mu = (n == 0) ? 0.0 : (double)k_par / k;
        const double mu2 = mu * mu;
        const double x = (double)xi[id];
        v[0] = k;
        v[1] = x;
        v[2] = x * (3.0 * mu2 - 1.0) / 2.0;
        v[3] = x * (35.0 * mu2 * mu2 - 30.0 * mu2 + 3.0) / 8.0;
        v[4] = 1.0;
    }
    warp_reduce_by_key<5>(bin, valid, v, [&](int b, const double (&s)[5]) {
#pragma unroll
        for (int q = 0; q < 5; q++) red_add(sums + q * g.kmax1 + b, s[q]);
    });
}

static SibGeom sib_geom(int dims, int ndim) {
    SibGeom g;
    g.dims = dims; g.middle = dims / 2; g.even = (dims % 2 == 0);
    g.kmax1 = isqrt_exact(ndim * g.middle * g.middle) + 1;
    return g;
}

__global__ void __launch_bounds__(256) scale_f32_kernel(float *x, int64_t n, float mul) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = __fmul_rn(x[i], mul);
}

static int make_tab(double **tab, int dims, int mas_index, cudaStream_t st) {
    keep_pool_memory();
    const int middle = dims / 2;
    PYLB_REQUIRE(mas_index >= 0 && mas_index <= 4, "MAS index %d out of range", mas_index);
    PYLB_CHECK(cudaMallocAsync(tab, sizeof(double) * (size_t)(middle + 1), st));
    sib_mas_table_kernel<<<(middle + 1 + 127) / 128, 128, 0, st>>>(*tab, middle, dims, mas_index);
    PYLB_LAUNCH_CHECK();
    return 0;
}

}  // namespace pylb

using namespace pylb;

#define PYLB_CUFFT2(expr)                                                                          \
    do {                                                                                           \
        cufftResult _r = (expr);                                                                   \
        if (_r != CUFFT_SUCCESS) {                                                                 \
            pylb::set_error("%s:%d: %s failed: cufftResult %d", __FILE__, __LINE__, #expr, (int)_r); \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

// pyfftw's FFTW.__call__ scales a backward transform by 1/N unless told otherwise (normalise_idft=True, the default
// the reference relies on: Xi divides by dims^3 only once, Pk_library.pyx:2139-2143): one fp32 multiply per element
// by (float)(1/N), as numpy's in-place `output_array *= scaling` does.
static int c2r_normalise(float *out, int64_t n, cudaStream_t st) {
    const float s = (float)(1.0 / (double)n);
    int64_t b = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (b > cap) b = cap;
    scale_f32_kernel<<<(unsigned)b, 256, 0, st>>>(out, n, s);
    PYLB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pylb_fft_c2r(void *in, float *out, int dims, int normalise, void *stream) {
    PYLB_REQUIRE(in && out && dims >= 2 && in != (void *)out, "pylb_fft_c2r: bad arguments");
    cufftHandle h;
    if (sib_plan(SK_C2R_3D, dims, &h)) return 1;
    PYLB_CUFFT2(cufftSetStream(h, (cudaStream_t)stream));
    PYLB_CUFFT2(cufftExecC2R(h, (cufftComplex *)in, out));
    count_launch();
    if (normalise) return c2r_normalise(out, (int64_t)dims * dims * dims, (cudaStream_t)stream);
    return 0;
}

extern "C" int pylb_fft2d_r2c(const float *in, void *out, int dims, void *stream) {
    PYLB_REQUIRE(in && out && dims >= 2 && (const void *)in != out, "pylb_fft2d_r2c: bad arguments");
    cufftHandle h;
    if (sib_plan(SK_R2C_2D, dims, &h)) return 1;
    PYLB_CUFFT2(cufftSetStream(h, (cudaStream_t)stream));
    PYLB_CUFFT2(cufftExecR2C(h, (cufftReal *)in, (cufftComplex *)out));
    count_launch();
    return 0;
}

extern "C" int pylb_fft2d_c2r(void *in, float *out, int dims, int normalise, void *stream) {
    PYLB_REQUIRE(in && out && dims >= 2 && in != (void *)out, "pylb_fft2d_c2r: bad arguments");
    cufftHandle h;
    if (sib_plan(SK_C2R_2D, dims, &h)) return 1;
    PYLB_CUFFT2(cufftSetStream(h, (cudaStream_t)stream));
    PYLB_CUFFT2(cufftExecC2R(h, (cufftComplex *)in, out));
    count_launch();
    if (normalise) return c2r_normalise(out, (int64_t)dims * dims, (cudaStream_t)stream);
    return 0;
}

extern "C" int pylb_mas_correct(void *dk, int dims, int mas_index, int mode, void *stream) {
    PYLB_REQUIRE(dk && dims >= 2 && (mode == 0 || mode == 1), "pylb_mas_correct: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    double *tab = nullptr;
    if (make_tab(&tab, dims, mas_index, st)) return 1;
    const SibGeom g = sib_geom(dims, 3);
    const long long total = (long long)dims * dims * (g.middle + 1);
    const unsigned blocks = (unsigned)((total + 255) / 256);
    if (mode == 0) mas_correct_kernel<0><<<blocks, 256, 0, st>>>((float2 *)dk, g, tab);
    else mas_correct_kernel<1><<<blocks, 256, 0, st>>>((float2 *)dk, g, tab);
    PYLB_LAUNCH_CHECK();
    cudaFreeAsync(tab, st);
    return 0;
}

extern "C" int pylb_theta_bin(const void *vx, const void *vy, const void *vz, int dims, int mas_index, double *sums,
                              void *stream) {
    PYLB_REQUIRE(vx && vy && vz && sums && dims >= 2, "pylb_theta_bin: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    double *tab = nullptr;
    if (make_tab(&tab, dims, mas_index, st)) return 1;
    const SibGeom g = sib_geom(dims, 3);
    PYLB_CHECK(cudaMemsetAsync(sums, 0, sizeof(double) * 3 * (size_t)g.kmax1, st));
    const long long total = (long long)dims * dims * (g.middle + 1);
    theta_bin_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const float2 *)vx, (const float2 *)vy,
                                                                      (const float2 *)vz, g, tab, sums);
    PYLB_LAUNCH_CHECK();
    cudaFreeAsync(tab, st);
    return 0;
}

extern "C" int pylb_plane_bin(const void *d1, const void *d2, int dims, int mas1, int mas2, double *sums, void *stream) {
    PYLB_REQUIRE(d1 && sums && dims >= 2 && dims <= 32768, "pylb_plane_bin: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    double *t1 = nullptr, *t2 = nullptr;
    if (make_tab(&t1, dims, mas1, st)) return 1;
    if (d2 && make_tab(&t2, dims, mas2, st)) return 1;
    const SibGeom g = sib_geom(dims, 2);
    PYLB_CHECK(cudaMemsetAsync(sums, 0, sizeof(double) * 5 * (size_t)g.kmax1, st));
    const int total = dims * (g.middle + 1);
    plane_bin_kernel<<<(total + 255) / 256, 256, 0, st>>>((const float2 *)d1, (const float2 *)d2, g, t1, t2, sums);
    PYLB_LAUNCH_CHECK();
    cudaFreeAsync(t1, st);
    if (t2) cudaFreeAsync(t2, st);
    return 0;
}

extern "C" int pylb_xi_bin(const float *xi, int dims, int axis, double *sums, void *stream) {
    PYLB_REQUIRE(xi && sums && dims >= 2 && axis >= 0 && axis <= 2, "pylb_xi_bin: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const SibGeom g = sib_geom(dims, 3);
    PYLB_CHECK(cudaMemsetAsync(sums, 0, sizeof(double) * 5 * (size_t)g.kmax1, st));
    const long long total = (long long)dims * dims * dims;
    xi_bin_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(xi, g, axis, sums);
    PYLB_LAUNCH_CHECK();
    return 0;
}
