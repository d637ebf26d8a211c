// riemann_b200 -- definitions of the changepoint kernels (changepoint.cu).
#pragma once
#include "common.cuh"

namespace cp {

constexpr int LANES = RMN_CP_LANES;
constexpr int NQ = 6;   // prediction query points tracked by the diagnostics
#ifndef RMN_CP_DIAG_EVERY
#define RMN_CP_DIAG_EVERY 16  // diagnostics functionals are accumulated every 16th MH step (tau of every tracked
                              // functional is > 100 steps on this posterior; 4 -> 16 is +13 % throughput, gpurun r2e)
#endif

struct CPParams {
    int M, P2, alpha_is_one, XP;  // XP = 2*P2: padded length of the x table (device data = xpad[XP] | cy[M+1] | cyy[M+1])
    double xmin, xmax, alpha, beta, cv, logL, ycenter, Mlog2pi, sqrtM;
    double tab1[LANES + 1];     // k log(lam) - gammaln(k) - lam        changepoint.py:134
    double tab2[LANES + 1];     // gammaln(2k+1)                        changepoint.py:143
    double sx[LANES + 1];       // sqrt(0.01 (xmax-xmin)/(k+1))         test_changepoint.py:36
    double sv, ss;              // sqrt(0.01 hscale^2/M), sqrt(0.01 hscale)   :37-38
    double p1, p2, p3;          // sequential selection thresholds      :28-30
    uint32_t t1, t2, t3, tpad;  // the same thresholds on the raw 32-bit Philox words
    double xq[NQ];
};

struct CPState {
    int32_t* k;          // [K]
    double* cpx;         // [K][LANES]
    double* cpv;         // [K][LANES]
    double* sig;         // [K]
    double* lp;          // [K]
    long long* dacc;     // [K]
    long long* dovf;     // [K]
    double* S1;          // [NDIAG][K]
    double* S2;          // [NDIAG][K]
};

__device__ __forceinline__ unsigned group_ballot(bool pred) {
    const unsigned full = __ballot_sync(0xffffffffu, pred);
    return (full >> (threadIdx.x & 16)) & 0xffffu;
}

// #{i : x_i <= c}  (upper bound, np.searchsorted(x, c, 'right')), branch-free, always in [0, M].
// The x table is padded with NaN up to XP = 2*P2 entries (P2 = largest power of two <= M): a NaN
// never compares <= c, so the search needs no bound check and stops at M by itself; a NaN / +inf
// query gives 0 / M like the unpadded search.  LOGP2 >= 0: P2 known at compile time, fully unrolled
// (one LDS + one compare + one predicated add per level); LOGP2 < 0: run-time P2.
template <int LOGP2>
__device__ __forceinline__ int upper_bound(const double* __restrict__ xs, int P2, double c) {
    const char* base = (const char*)xs;
    unsigned off = 0;                               // byte offset of the running position
    if (LOGP2 >= 0) {
#pragma unroll
        for (int b = LOGP2; b >= 0; --b)
            if (*(const double*)(base + off + ((1 << b) - 1) * 8) <= c) off += 8u << b;
    } else {
        for (int step = P2; step > 0; step >>= 1)
            if (*(const double*)(base + off + (step - 1) * 8) <= c) off += 8u * step;
    }
    return (int)(off >> 3);
}

// R independent searches, levels interleaved (the R dependent LDS chains overlap)
template <int LOGP2, int R>
__device__ __forceinline__ void upper_bound_rows(const double* __restrict__ xs, int P2, const double (&c)[R],
                                                 int (&pos)[R]) {
    const char* base = (const char*)xs;
    unsigned off[R];
#pragma unroll
    for (int r = 0; r < R; ++r) off[r] = 0;
    if (LOGP2 >= 0) {
#pragma unroll
        for (int b = LOGP2; b >= 0; --b) {
            double v[R];
#pragma unroll
            for (int r = 0; r < R; ++r) v[r] = *(const double*)(base + off[r] + ((1 << b) - 1) * 8);
#pragma unroll
            for (int r = 0; r < R; ++r) if (v[r] <= c[r]) off[r] += 8u << b;
        }
    } else {
        for (int step = P2; step > 0; step >>= 1) {
            double v[R];
#pragma unroll
            for (int r = 0; r < R; ++r) v[r] = *(const double*)(base + off[r] + (step - 1) * 8);
#pragma unroll
            for (int r = 0; r < R; ++r) if (v[r] <= c[r]) off[r] += 8u * step;
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) pos[r] = (int)(off[r] >> 3);
}

// uniform on (0,1) from 32 random bits with 2 fp64 instructions: [1,2) mantissa trick + 2^-33
__device__ __forceinline__ double u01_fast(uint32_t x) {
    return __hiloint2double(0x3ff00000 | (x >> 12), x << 20) - (1.0 - 1.1641532182693481e-10);
}


}  // namespace cp
