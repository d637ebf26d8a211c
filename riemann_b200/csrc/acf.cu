// riemann_b200 -- integrated autocorrelation time of traced chains ON THE DEVICE: the estimator
// `emcee.autocorr.integrated_time` implements and the reference calls at examples/test_randomwalk.py:42 to report
// "steps per independent sample".  emcee is an un-vendored, un-pinned dependency of the reference; its published
// algorithm (Sokal 1989 / Goodman & Weare 2010) is:
//     rho(t) = autocorrelation function, by FFT of the centred chain zero-padded to 2 * next_pow2(n);
//              several chains ("walkers"): each chain's function normalised to rho(0) = 1, then averaged
//     tau(W) = 2 * sum_{t <= W} rho(t) - 1
//     window = first W with W >= c * tau(W) (the last lag if there is none);   result tau(window)
// Input is the thinned device trace exactly as rmn_trace_t.d_theta holds it: x[n][K][nd].  The FFTs are batched cuFFT
// D2Z / Z2D plans over all K * nd series; cuFFT is resolved with dlopen (the copy the host process already loaded,
// else the toolkit's), so the library has no link-time dependency on it and the call fails loudly if it is absent.
// Host restatement of the same algorithm: riemann_b200/diagnostics.py (the test compares the two).
#include "common.cuh"

#include <dlfcn.h>
#include <cstdlib>
#include <vector>

namespace {

typedef int cufftHandle;
typedef int cufftResult;
constexpr int CUFFT_D2Z_ = 0x6a, CUFFT_Z2D_ = 0x6c;

struct CufftApi {
    void* h = nullptr;
    cufftResult (*PlanMany)(cufftHandle*, int, int*, int*, int, int, int*, int, int, int, int) = nullptr;
    cufftResult (*SetStream)(cufftHandle, cudaStream_t) = nullptr;
    cufftResult (*ExecD2Z)(cufftHandle, double*, double2*) = nullptr;
    cufftResult (*ExecZ2D)(cufftHandle, double2*, double*) = nullptr;
    cufftResult (*Destroy)(cufftHandle) = nullptr;
    bool ok = false;
};

CufftApi& fft() {
    static CufftApi a;
    static bool tried = false;
    if (tried) return a;
    tried = true;
    const char* env = getenv("RMN_CUFFT_LIB");
    if (env && *env) a.h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!a.h) a.h = dlopen("libcufft.so.11", RTLD_NOW | RTLD_NOLOAD);
    if (!a.h) a.h = dlopen("libcufft.so.11", RTLD_NOW | RTLD_GLOBAL);
    if (!a.h) a.h = dlopen("/usr/local/cuda/lib64/libcufft.so.11", RTLD_NOW | RTLD_GLOBAL);
    if (!a.h) a.h = dlopen("libcufft.so", RTLD_NOW | RTLD_GLOBAL);
    if (!a.h) return a;
#define RMN_SYM(field, name) *(void**)(&a.field) = dlsym(a.h, name)
    RMN_SYM(PlanMany, "cufftPlanMany");
    RMN_SYM(SetStream, "cufftSetStream");
    RMN_SYM(ExecD2Z, "cufftExecD2Z");
    RMN_SYM(ExecZ2D, "cufftExecZ2D");
    RMN_SYM(Destroy, "cufftDestroy");
#undef RMN_SYM
    a.ok = a.PlanMany && a.SetStream && a.ExecD2Z && a.ExecZ2D && a.Destroy;
    return a;
}

#define RMN_CUFFT(call)                                                                 \
    do {                                                                                \
        const cufftResult r_ = (call);                                                  \
        if (r_ != 0) {                                                                  \
            rmn_set_error("cuFFT error %d at %s:%d", r_, __FILE__, __LINE__);           \
            return RMN_ERR_CUDA;                                                        \
        }                                                                               \
    } while (0)

// series b = k * nd + j of x[n][B]: centred on its own mean, zero-padded to L.  One warp per series for the mean
// (lanes stride over time), then the same warp writes the padded row.
__global__ void __launch_bounds__(256)
acf_center_pad_kernel(const double* __restrict__ x, int64_t n, int64_t B, int64_t L, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t b = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;
    double s = 0.0;
    for (int64_t t = lane; t < n; t += 32) s += x[t * B + b];
    s = group_sum<32>(s);
    const double mean = s / (double)n;
    double* o = out + b * L;
    for (int64_t t = lane; t < L; t += 32) o[t] = (t < n) ? x[t * B + b] - mean : 0.0;
}

__global__ void __launch_bounds__(256)
acf_power_kernel(double2* __restrict__ f, int64_t count) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double2 v = f[i];
    f[i] = make_double2(v.x * v.x + v.y * v.y, 0.0);
}

// rho[j][t] = mean over chains k of acov[k*nd+j][t] / acov[k*nd+j][0]
__global__ void __launch_bounds__(256)
acf_average_kernel(const double* __restrict__ acov, int64_t n, int64_t K, int64_t nd, int64_t L,
                   double* __restrict__ rho) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n * nd) return;
    const int64_t j = i / n, t = i % n;
    double s = 0.0;
    for (int64_t k = 0; k < K; ++k) {
        const double* a = acov + (k * nd + j) * L;
        s += a[t] / a[0];
    }
    rho[j * n + t] = s / (double)K;
}

// one thread per functional: running tau(W) = 2 cumsum(rho) - 1, stop at the first W >= c tau(W)
__global__ void acf_window_kernel(const double* __restrict__ rho, int64_t n, int64_t nd, double c,
                                  double* __restrict__ tau, long long* __restrict__ window) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= nd) return;
    const double* r = rho + j * n;
    double cs = 0.0, tw = 0.0;
    long long w = n - 1;
    bool found = false;
    for (int64_t t = 0; t < n; ++t) {
        cs += r[t];
        tw = 2.0 * cs - 1.0;
        if (!((double)t < c * tw)) { w = t; found = true; break; }
    }
    (void)found;                        // no window: tw already holds tau(n - 1), the last lag
    tau[j] = tw;
    window[j] = w;
}

}  // namespace

int rmn_autocorr_tau_impl(const double* d_x, int64_t n, int64_t K, int64_t nd, double c, double* h_tau,
                          int64_t* h_window, cudaStream_t stream) {
    if (!d_x || !h_tau || n < 2 || K < 1 || nd < 1 || !(c > 0.0)) {
        rmn_set_error("rmn_autocorr_tau: need a trace x[n >= 2][K >= 1][nd >= 1], c > 0 and an output array");
        return RMN_ERR_PARAM;
    }
    if (!fft().ok) {
        rmn_set_error("cuFFT (libcufft.so.11) could not be loaded: rmn_autocorr_tau needs it (set RMN_CUFFT_LIB to its path)");
        return RMN_ERR_UNSUPPORTED;
    }
    int64_t m = 1;
    while (m < n) m <<= 1;
    const int64_t L = 2 * m, B = K * nd, Lc = L / 2 + 1;
    if (L > (int64_t)1 << 30 || B > (int64_t)1 << 30) {
        rmn_set_error("rmn_autocorr_tau: trace too long for one batched plan");
        return RMN_ERR_PARAM;
    }
    double* d_pad = nullptr; double2* d_f = nullptr; double* d_rho = nullptr; double* d_tau = nullptr; long long* d_w = nullptr;
    auto cleanup = [&]() { cudaFree(d_pad); cudaFree(d_f); cudaFree(d_rho); cudaFree(d_tau); cudaFree(d_w); };
    cudaError_t e = cudaMalloc(&d_pad, (size_t)B * L * 8);
    if (e == cudaSuccess) e = cudaMalloc(&d_f, (size_t)B * Lc * 16);
    if (e == cudaSuccess) e = cudaMalloc(&d_rho, (size_t)nd * n * 8);
    if (e == cudaSuccess) e = cudaMalloc(&d_tau, (size_t)nd * 8);
    if (e == cudaSuccess) e = cudaMalloc(&d_w, (size_t)nd * 8);
    if (e != cudaSuccess) {
        cleanup();
        rmn_set_error("rmn_autocorr_tau: scratch allocation failed (%s)", cudaGetErrorString(e));
        return RMN_ERR_CUDA;
    }
    int len = (int)L;
    cufftHandle fwd = 0, inv = 0;
    int rc = RMN_OK;
    auto run = [&]() -> int {
        RMN_CUFFT(fft().PlanMany(&fwd, 1, &len, nullptr, 1, (int)L, nullptr, 1, (int)Lc, CUFFT_D2Z_, (int)B));
        RMN_CUFFT(fft().PlanMany(&inv, 1, &len, nullptr, 1, (int)Lc, nullptr, 1, (int)L, CUFFT_Z2D_, (int)B));
        RMN_CUFFT(fft().SetStream(fwd, stream));
        RMN_CUFFT(fft().SetStream(inv, stream));
        acf_center_pad_kernel<<<(unsigned)((B * 32 + 255) / 256), 256, 0, stream>>>(d_x, n, B, L, d_pad);
        RMN_KERNEL_CHECK();
        RMN_CUFFT(fft().ExecD2Z(fwd, d_pad, d_f));
        acf_power_kernel<<<(unsigned)((B * Lc + 255) / 256), 256, 0, stream>>>(d_f, B * Lc);
        RMN_KERNEL_CHECK();
        RMN_CUFFT(fft().ExecZ2D(inv, d_f, d_pad));
        acf_average_kernel<<<(unsigned)((n * nd + 255) / 256), 256, 0, stream>>>(d_pad, n, K, nd, L, d_rho);
        RMN_KERNEL_CHECK();
        acf_window_kernel<<<(unsigned)((nd + 63) / 64), 64, 0, stream>>>(d_rho, n, nd, c, d_tau, d_w);
        RMN_KERNEL_CHECK();
        RMN_CUDA(cudaMemcpyAsync(h_tau, d_tau, (size_t)nd * 8, cudaMemcpyDeviceToHost, stream));
        if (h_window) {
            static_assert(sizeof(long long) == sizeof(int64_t), "window words");
            RMN_CUDA(cudaMemcpyAsync(h_window, d_w, (size_t)nd * 8, cudaMemcpyDeviceToHost, stream));
        }
        RMN_CUDA(cudaStreamSynchronize(stream));
        return RMN_OK;
    };
    rc = run();
    if (fwd) fft().Destroy(fwd);
    if (inv) fft().Destroy(inv);
    cleanup();
    return rc;
}
