// cuFFT plumbing: cached plans with caller-owned work areas.
//
// Replaces FFT3Dr_f (library/Pk_library/Pk_library.pyx:120-133; pyfftw/FFTW, FFTW_ESTIMATE, forward,
// unnormalised, float32 -> complex64 with the half spectrum on the last axis).  The slab pieces have
// no reference counterpart (the reference is single-process): real x-slabs -> batched 2-D R2C over
// (y,z) -> all-to-all transpose -> strided 1-D C2C along x on [dims][ny_local][dims/2+1].
#include <cufft.h>

#include <map>
#include <tuple>

#include "common.cuh"

namespace pylb {

static const char *cufft_str(cufftResult r) {
    switch (r) {
        case CUFFT_SUCCESS: return "CUFFT_SUCCESS";
        case CUFFT_INVALID_PLAN: return "CUFFT_INVALID_PLAN";
        case CUFFT_ALLOC_FAILED: return "CUFFT_ALLOC_FAILED";
        case CUFFT_INVALID_TYPE: return "CUFFT_INVALID_TYPE";
        case CUFFT_INVALID_VALUE: return "CUFFT_INVALID_VALUE";
        case CUFFT_INTERNAL_ERROR: return "CUFFT_INTERNAL_ERROR";
        case CUFFT_EXEC_FAILED: return "CUFFT_EXEC_FAILED";
        case CUFFT_SETUP_FAILED: return "CUFFT_SETUP_FAILED";
        case CUFFT_INVALID_SIZE: return "CUFFT_INVALID_SIZE";
        case CUFFT_UNALIGNED_DATA: return "CUFFT_UNALIGNED_DATA";
        default: return "CUFFT_<other>";
    }
}

#define PYLB_CUFFT(expr)                                                                          \
    do {                                                                                          \
        cufftResult _r = (expr);                                                                  \
        if (_r != CUFFT_SUCCESS) {                                                                \
            pylb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, pylb::cufft_str(_r)); \
            return 1;                                                                             \
        }                                                                                         \
    } while (0)

enum { KIND_R2C_3D = 0, KIND_SLAB_YZ = 1, KIND_SLAB_X = 2, KIND_R2C_3D_PITCHED = 3 };

struct Plan {
    cufftHandle h = 0;
    size_t work = 0;
};

typedef std::tuple<int, int, int, long long, long long> PlanKey;  // device, kind, dims, nloc | in_pitch, inplace | out_pitch
static std::map<PlanKey, Plan> g_plans;

// KIND_R2C_3D_PITCHED: nloc = input row pitch in floats, inplace = output row pitch in complex elements
static int get_plan(int kind, int dims, long long nloc, long long inplace, Plan **out) {
    int dev = 0;
    PYLB_CHECK(cudaGetDevice(&dev));
    PlanKey key(dev, kind, dims, nloc, inplace);
    auto it = g_plans.find(key);
    if (it != g_plans.end()) {
        *out = &it->second;
        return 0;
    }
    Plan p;
    PYLB_CUFFT(cufftCreate(&p.h));
    PYLB_CUFFT(cufftSetAutoAllocation(p.h, 0));
    const long long N = dims, nz = dims / 2 + 1;
    if (kind == KIND_R2C_3D) {
        long long n[3] = {N, N, N};
        if (inplace) {
            long long inembed[3] = {N, N, 2 * nz}, onembed[3] = {N, N, nz};
            PYLB_CUFFT(cufftMakePlanMany64(p.h, 3, n, inembed, 1, N * N * 2 * nz, onembed, 1, N * N * nz,
                                           CUFFT_R2C, 1, &p.work));
        } else {
            PYLB_CUFFT(cufftMakePlanMany64(p.h, 3, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_R2C, 1, &p.work));
        }
    } else if (kind == KIND_R2C_3D_PITCHED) {
        long long n[3] = {N, N, N};
        long long inembed[3] = {N, N, nloc}, onembed[3] = {N, N, inplace};
        PYLB_CUFFT(cufftMakePlanMany64(p.h, 3, n, inembed, 1, N * N * nloc, onembed, 1, N * N * inplace, CUFFT_R2C, 1, &p.work));
    } else if (kind == KIND_SLAB_YZ) {           // inplace = output row pitch in complex elements (>= nz)
        long long n[2] = {N, N};
        long long inembed[2] = {N, N}, onembed[2] = {N, inplace};
        PYLB_CUFFT(cufftMakePlanMany64(p.h, 2, n, inembed, 1, N * N, onembed, 1, N * inplace, CUFFT_R2C, nloc, &p.work));
    } else {                                     // KIND_SLAB_X: inplace = row pitch; the padding column rides along
        long long n[1] = {N};
        const long long batch = (long long)nloc * inplace;
        long long embed[1] = {N};
        PYLB_CUFFT(cufftMakePlanMany64(p.h, 1, n, embed, batch, 1, embed, batch, 1, CUFFT_C2C, batch, &p.work));
    }
    auto res = g_plans.emplace(key, p);
    *out = &res.first->second;
    return 0;
}

static int exec_setup(Plan *p, void *work, size_t work_bytes, cudaStream_t st) {
    PYLB_REQUIRE(work_bytes >= p->work && (p->work == 0 || work != nullptr),
                 "cuFFT work area too small: need %zu bytes, got %zu", p->work, work_bytes);
    PYLB_CUFFT(cufftSetStream(p->h, st));
    if (p->work) PYLB_CUFFT(cufftSetWorkArea(p->h, work));
    return 0;
}

}  // namespace pylb

using namespace pylb;

extern "C" size_t pylb_fft_r2c_work_bytes(int dims, int inplace) {
    Plan *p = nullptr;
    if (dims < 2 || get_plan(KIND_R2C_3D, dims, 0, inplace ? 1 : 0, &p)) return (size_t)-1;
    return p->work;
}

extern "C" int pylb_fft_r2c(const float *in, void *out, int dims, int inplace, void *work, size_t work_bytes,
                            void *stream) {
    PYLB_REQUIRE(dims >= 2, "pylb_fft_r2c: dims must be >= 2");
    PYLB_REQUIRE(inplace || (const void *)in != out, "pylb_fft_r2c: in == out requires inplace=1 (padded layout)");
    Plan *p = nullptr;
    if (get_plan(KIND_R2C_3D, dims, 0, inplace ? 1 : 0, &p)) return 1;
    if (exec_setup(p, work, work_bytes, (cudaStream_t)stream)) return 1;
    timing_begin(PYLB_T_FFT, (cudaStream_t)stream);
    PYLB_CUFFT(cufftExecR2C(p->h, (cufftReal *)in, (cufftComplex *)out));
    timing_end(PYLB_T_FFT, (cudaStream_t)stream);
    count_launch();
    return 0;
}

extern "C" size_t pylb_fft_r2c_pitched_work_bytes(int dims, int64_t in_pitch, int64_t out_pitch) {
    Plan *p = nullptr;
    if (dims < 2 || in_pitch < dims || out_pitch < dims / 2 + 1 || get_plan(KIND_R2C_3D_PITCHED, dims, in_pitch, out_pitch, &p))
        return (size_t)-1;
    return p->work;
}

extern "C" int pylb_fft_r2c_pitched(const float *in, int64_t in_pitch, void *out, int64_t out_pitch, int dims, void *work,
                                    size_t work_bytes, void *stream) {
    PYLB_REQUIRE(dims >= 2 && in_pitch >= dims && out_pitch >= dims / 2 + 1, "pylb_fft_r2c_pitched: bad shape");
    PYLB_REQUIRE((const void *)in != out || in_pitch == 2 * out_pitch,
                 "pylb_fft_r2c_pitched: in == out requires in_pitch == 2*out_pitch");
    Plan *p = nullptr;
    if (get_plan(KIND_R2C_3D_PITCHED, dims, in_pitch, out_pitch, &p)) return 1;
    if (exec_setup(p, work, work_bytes, (cudaStream_t)stream)) return 1;
    timing_begin(PYLB_T_FFT, (cudaStream_t)stream);
    PYLB_CUFFT(cufftExecR2C(p->h, (cufftReal *)in, (cufftComplex *)out));
    timing_end(PYLB_T_FFT, (cudaStream_t)stream);
    count_launch();
    return 0;
}

extern "C" size_t pylb_fft_slab_yz_work_bytes(int dims, int nx_local, int64_t out_pitch) {
    Plan *p = nullptr;
    if (dims < 2 || nx_local < 1 || out_pitch < dims / 2 + 1 || get_plan(KIND_SLAB_YZ, dims, nx_local, out_pitch, &p)) return (size_t)-1;
    return p->work;
}

extern "C" int pylb_fft_slab_yz(const float *in, void *out, int dims, int nx_local, int64_t out_pitch, void *work,
                                size_t work_bytes, void *stream) {
    PYLB_REQUIRE(dims >= 2 && nx_local >= 1 && out_pitch >= dims / 2 + 1, "pylb_fft_slab_yz: bad shape");
    Plan *p = nullptr;
    if (get_plan(KIND_SLAB_YZ, dims, nx_local, out_pitch, &p)) return 1;
    if (exec_setup(p, work, work_bytes, (cudaStream_t)stream)) return 1;
    timing_begin(PYLB_T_FFT, (cudaStream_t)stream);
    PYLB_CUFFT(cufftExecR2C(p->h, (cufftReal *)in, (cufftComplex *)out));
    timing_end(PYLB_T_FFT, (cudaStream_t)stream);
    count_launch();
    return 0;
}

extern "C" size_t pylb_fft_slab_x_work_bytes(int dims, int ny_local, int64_t pitch) {
    Plan *p = nullptr;
    if (dims < 2 || ny_local < 1 || pitch < dims / 2 + 1 || get_plan(KIND_SLAB_X, dims, ny_local, pitch, &p)) return (size_t)-1;
    return p->work;
}

extern "C" int pylb_fft_slab_x(void *data, int dims, int ny_local, int64_t pitch, void *work, size_t work_bytes, void *stream) {
    PYLB_REQUIRE(dims >= 2 && ny_local >= 1 && pitch >= dims / 2 + 1, "pylb_fft_slab_x: bad shape");
    Plan *p = nullptr;
    if (get_plan(KIND_SLAB_X, dims, ny_local, pitch, &p)) return 1;
    if (exec_setup(p, work, work_bytes, (cudaStream_t)stream)) return 1;
    timing_begin(PYLB_T_FFT, (cudaStream_t)stream);
    PYLB_CUFFT(cufftExecC2C(p->h, (cufftComplex *)data, (cufftComplex *)data, CUFFT_FORWARD));
    timing_end(PYLB_T_FFT, (cudaStream_t)stream);
    count_launch();
    return 0;
}
