// riemann_b200 -- small generic kernels: fills, diagnostics reduction, RNG test entry points.
#include "common.cuh"
#include <map>
#include <mutex>
#include <utility>

namespace {

__global__ void fill_f64_kernel(double* p, int64_t n, double v) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        p[i] = v;
}
__global__ void fill_i64_kernel(long long* p, int64_t n, long long v) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        p[i] = v;
}

__device__ __forceinline__ double block_sum_256(double v, double* sh) {
    v = group_sum<32>(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double r = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
    if (w == 0) r = group_sum<32>(r);
    return r;   // valid in warp 0
}

// S1/S2 laid out [nd][K] (chain fastest).  Block j < nd reduces functional j over
// chains; block nd reduces the counters.  Plain stores, no atomics.
__global__ void __launch_bounds__(256)
reduce_diag_kernel(int64_t K, int nd, int64_t nsamples, int64_t nsteps, const double* __restrict__ S1,
                   const double* __restrict__ S2, const long long* __restrict__ acc,
                   const long long* __restrict__ ovf, double* __restrict__ block) {
    __shared__ double sh[8];
    const int j = blockIdx.x;
    if (j < nd) {
        double sm = 0, sm2 = 0, sv = 0;
        const double inv = nsamples > 0 ? 1.0 / (double)nsamples : 0.0;
        for (int64_t c = threadIdx.x; c < K; c += blockDim.x) {
            const double m = S1[(int64_t)j * K + c] * inv;
            const double v = S2[(int64_t)j * K + c] * inv - m * m;
            sm += m; sm2 += m * m; sv += v;
        }
        sm = block_sum_256(sm, sh);
        sm2 = block_sum_256(sm2, sh);
        sv = block_sum_256(sv, sh);
        if (threadIdx.x == 0) {
            block[RMN_DIAG_HDR + j] = sm;
            block[RMN_DIAG_HDR + nd + j] = sm2;
            block[RMN_DIAG_HDR + 2 * nd + j] = sv;
        }
    } else {
        double a = 0, o = 0;
        for (int64_t c = threadIdx.x; c < K; c += blockDim.x) {
            a += (double)acc[c];
            if (ovf) o += (double)ovf[c];
        }
        a = block_sum_256(a, sh);
        o = block_sum_256(o, sh);
        if (threadIdx.x == 0) {
            block[0] = (double)K;
            block[1] = (double)nsamples;
            block[2] = a;
            block[3] = o;
            block[4] = (double)nsteps;
            block[5] = 0.0;
        }
    }
}

__global__ void copy_adapt_kernel(int64_t K, const double* sc_in, const int64_t* ns_in, const int64_t* na_in,
                                  double* sc, long long* ns, long long* na) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= K) return;
    if (sc_in) sc[c] = sc_in[c];
    if (ns_in) ns[c] = ns_in[c];
    if (na_in) na[c] = na_in[c];
}

__global__ void philox_raw_kernel(int64_t n, const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 r = philox4x32_10(make_uint4(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3]),
                                  make_uint2(key[2 * i], key[2 * i + 1]));
    out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
}

__global__ void rng_draws_kernel(uint64_t seed, int64_t chain0, int64_t step, int64_t n, int nn,
                                 double* normals, double* uniform) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    RngKey rk(seed, (uint64_t)(chain0 + i));
    for (int b = 0; 4 * b < nn; ++b) {
        double v[4];
        normal4(rk.block((uint64_t)step, (uint32_t)b), v);
        for (int q = 0; q < 4 && 4 * b + q < nn; ++q) normals[i * nn + 4 * b + q] = v[q];
    }
    uniform[i] = u01(rk.block((uint64_t)step, RMN_BLOCK_ACCEPT).x);
}

__global__ void chain_moments_kernel(int64_t n, double inv, const double* __restrict__ S1, const double* __restrict__ S2,
                                     double* __restrict__ mean, double* __restrict__ var) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double m = S1[i] * inv;
    if (mean) mean[i] = m;
    if (var) var[i] = S2[i] * inv - m * m;
}

}  // namespace

cudaError_t rmn_raise_dyn_smem(const void* kernel, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> limit;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    size_t& cur = limit[std::make_pair(dev, kernel)];
    if (bytes <= cur) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) cur = bytes;
    return e;
}

int rmn_chain_moments(int64_t K, int nd, int64_t nsamples, const double* S1, const double* S2, double* d_mean,
                      double* d_var, cudaStream_t st) {
    const int64_t n = K * (int64_t)nd;
    if (n <= 0) return RMN_OK;
    RMN_REQUIRE(nsamples > 0, "rmn_sampler_chain_moments: no samples accumulated since the last reset");
    chain_moments_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, 1.0 / (double)nsamples, S1, S2, d_mean, d_var);
    RMN_KERNEL_CHECK();
    return RMN_OK;
}

int rmn_fill_f64(double* p, int64_t n, double v, cudaStream_t st) {
    if (n <= 0) return RMN_OK;
    int grid = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
    fill_f64_kernel<<<grid, 256, 0, st>>>(p, n, v);
    RMN_KERNEL_CHECK();
    return RMN_OK;
}
int rmn_fill_i64(long long* p, int64_t n, long long v, cudaStream_t st) {
    if (n <= 0) return RMN_OK;
    int grid = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
    fill_i64_kernel<<<grid, 256, 0, st>>>(p, n, v);
    RMN_KERNEL_CHECK();
    return RMN_OK;
}

int rmn_copy_adapt(int64_t K, const double* sc_in, const int64_t* ns_in, const int64_t* na_in, double* sc,
                   long long* ns, long long* na, cudaStream_t st) {
    copy_adapt_kernel<<<(unsigned)((K + 127) / 128), 128, 0, st>>>(K, sc_in, ns_in, na_in, sc, ns, na);
    RMN_KERNEL_CHECK();
    return RMN_OK;
}

int rmn_reduce_diag_block(int64_t K, int nd, int64_t nsamples, int64_t nsteps, const double* S1, const double* S2,
                          const long long* acc, const long long* ovf, double* d_block,
                          cudaStream_t st) {
    reduce_diag_kernel<<<nd + 1, 256, 0, st>>>(K, nd, nsamples, nsteps, S1, S2, acc, ovf, d_block);
    RMN_KERNEL_CHECK();
    return RMN_OK;
}

extern "C" int rmn_philox_raw(int64_t n, const uint32_t* d_ctr, const uint32_t* d_key,
                              uint32_t* d_out, void* stream) {
    RMN_REQUIRE(n >= 0 && d_ctr && d_key && d_out, "rmn_philox_raw: null argument");
    if (n == 0) return RMN_OK;
    philox_raw_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(n, d_ctr, d_key, d_out);
    RMN_KERNEL_CHECK();
    return RMN_OK;
}

extern "C" int rmn_rng_draws(uint64_t seed, int64_t chain0, int64_t step, int64_t n, int nn,
                             double* d_normals, double* d_uniform, void* stream) {
    RMN_REQUIRE(n >= 0 && nn >= 0 && d_normals && d_uniform, "rmn_rng_draws: bad argument");
    if (n == 0) return RMN_OK;
    rng_draws_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        seed, chain0, step, n, nn, d_normals, d_uniform);
    RMN_KERNEL_CHECK();
    return RMN_OK;
}
