#include "common.cuh"
SamplerImpl* make_dense_gauss_sampler(rmn_sampler* s) { rmn_set_error("dense sampler not built yet"); return nullptr; }

namespace {
// one warp per point: v = P y, quad = y.v, grad = -v      (gaussian.py:49-58 via P = C^-1)
__global__ void __launch_bounds__(128)
gauss_point_kernel(int d, const double* __restrict__ mu, const double* __restrict__ prec, double c1,
                   double c2, int which, int64_t n, const double* __restrict__ theta,
                   double* __restrict__ out, double* __restrict__ grad) {
    const int64_t pt = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (pt >= n) return;
    const double* th = theta + pt * d;
    double quad = 0.0;
    for (int i = 0; i < d; ++i) {
        double s = 0.0;
        for (int j = lane; j < d; j += 32) s += prec[(size_t)i * d + j] * (th[j] - mu[j]);
        s = group_sum<32>(s);
        if (grad && lane == 0) grad[pt * d + i] = -s;
        quad += s * (th[i] - mu[i]);
    }
    if (out && lane == 0) {
        const double ll = -0.5 * ((quad + c1) + c2);
        out[pt] = (which == 2) ? 0.0 : ((which == 1) ? ll : combine_logpost(0.0, ll));
    }
}
}  // namespace

int gauss_pointwise(rmn_model* m, int which, int64_t n, const double* d_theta, double* d_out,
                    double* d_grad, cudaStream_t st) {
    if (n <= 0) return RMN_OK;
    gauss_point_kernel<<<(unsigned)((n * 32 + 127) / 128), 128, 0, st>>>(
        m->d, m->d_mu, m->d_prec, m->d * log(2.0 * M_PI), m->logdetC, which, n, d_theta, d_out, d_grad);
    RMN_KERNEL_CHECK();
    return RMN_OK;
}
