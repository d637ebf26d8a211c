// riemann_b200 -- Bayesian logistic regression: MALA (config 4) and simplified manifold
// MALA with the Fisher metric (config 5).  Neither model nor metric proposal exists in the
// reference; they follow its protocols -- Model.log_posterior (riemann/models/model.py:43-55),
// Proposal.propose -> (theta', log q(theta'|theta)/q(theta|theta')) (riemann/proposals/
// proposal.py:10-17), MALA == VanillaHMC(eps, 1, grad) (riemann/proposals/hamiltonian.py:76-91),
// Sampler.sample (riemann/samplers/sampler.py:72-90) -- and are checked against the fp64 numpy
// restatement in oracle/riemann_port.py.
//
//   log L(theta) = sum_i y_i z_i - softplus(z_i),  z = X theta       prior N(0, pv I)
//   grad         = X^T (y - sigmoid(z)) - theta / pv
//   metric G     = X^T diag(p (1-p)) X + I / pv
//
// lg_eval_kernel  ("flash-attention shaped"): a CTA owns 64 chains and a range of data rows.
//   Per 64-row tile of X (cp.async double-buffered in shared memory, read ONCE for all 64
//   chains):  Z^T = Theta X^T  on the fp64 tensor cores (DMMA m8n8k4)  ->  pointwise
//   sigmoid / softplus in registers  ->  G += R^T X  on the tensor cores again.  The C
//   fragment of the first product is fed straight into the A fragment of the second (the
//   contraction index is permuted instead of transposing through shared memory).  Z never
//   touches memory; the log-likelihood is accumulated in fp64.
// lg_metric_kernel: per chain  G = X^T diag(w) X  as a 64x64xN DMMA product, two warps per
//   chain, w recomputed in-kernel from z.
// mm_geometry (inside the finish/propose kernel): per-chain Cholesky, log-determinant and
//   triangular solves in shared memory, one warp per chain.
//
// Precision mode RMN_PREC_TF32_METRIC (mMALA only): the Fisher metric is a dense contraction over the
// data rows, G_c[a][b] = sum_i w_ic x_ia x_ib, i.e. ONE GEMM  Gp[K][d(d+1)/2] = W[K][N] . KR[d(d+1)/2][N]^T
// with the Khatri-Rao table KR[(a,b)][i] = x_ia x_ib built once per sampler and the weights
// W[c][i] = p(1-p) written by lg_eval_kernel (which has z in registers anyway).  It runs on the tcgen05
// tensor cores in single-pass TF32 (tc_gemm.cu, PASSES = 1).  The metric only shapes the proposal: the
// SAME deterministic function theta -> G~(theta) enters the forward and the reverse proposal density,
// and the log-posterior / gradient stay fp64, so the chain still targets the exact posterior.
//
// Precision mode RMN_PREC_TF32X3 (MALA and mMALA): the likelihood sweep itself on the tcgen05 tensor cores, fp32-accurate
// (3xTF32), as ONE fused kernel (logistic_fused.cu): Z = Theta' X^T into tensor memory, sigmoid / softplus straight out
// of it, the log-likelihood summed in fp64, R = y - p written back into tensor memory as the A operand of G += R X; Z and
// R never touch HBM.  The log-likelihood is a deterministic function of theta within the budget of
// riemann_b200/budgets.py; the accept test, prior, proposal arithmetic and Cholesky stay fp64.
#include <algorithm>
#include <cuda_bf16.h>
#include "common.cuh"
#include "tc_gemm.cuh"
#include "logistic_math.cuh"
#include "logistic_fused.cuh"
#include <stdlib.h>
#include <stdio.h>

namespace tc {
int launch_plain_tf32(const GemmMaps& maps, int64_t M, int N, int Kdim, float* C, int ldc, cudaStream_t st);
int launch_plain_bf16(const GemmMaps& maps, int64_t M, int N, int Kdim, float* C, int ldc, cudaStream_t st);
int splits_used(int Kdim, int ksplit, bool bf16);
int launch_single_splitk(const GemmMaps& maps, int64_t M, int N, int Kdim, float* C, int ldc, int ksplit,
                         int64_t split_stride, bool bf16, cudaStream_t st);
}

namespace {

__device__ double g_lg_tab[lgmath::TAB_DOUBLES];     // sigmoid/softplus tables (logistic_math.cuh)
static int lg_tables_ready() {
    static int done[64] = {0};                        // per device
    int dev = 0;
    RMN_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && done[dev]) return RMN_OK;
    double h[lgmath::TAB_DOUBLES];
    lgmath::fill_tables(h);
    RMN_CUDA(cudaMemcpyToSymbol(g_lg_tab, h, sizeof(h)));
    if (dev < 64) done[dev] = 1;
    return RMN_OK;
}

constexpr int BC = 64;        // chains per CTA (eval)
constexpr int BI = 64;        // data rows per tile
constexpr int EVAL_THREADS = 512;
constexpr int ND_MAX = 8;
constexpr int MM_DMAX = 64;   // mMALA: d <= 64 (shared-memory Cholesky, 64x64 accumulators)

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int n = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(n));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int n = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(n));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// row permutation inside an 8-row group that makes BOTH products' shared-memory fragment
// loads bank-conflict free with a row stride = 4 (mod 16) doubles
__device__ __forceinline__ int perm8(int g) { return (g < 4) ? g : (g ^ 1); }   // 0 1 2 3 5 4 7 6

struct LogisticState {
    int64_t K, N; int d, dp, ldt, nsplit; int64_t rows_per_split;
    int eval_tab_off;   // offset (doubles) of the sigmoid/softplus tables in lg_eval_kernel's shared memory
    double pv;
    const double* X; const double* y;
    double* Th;      // [2][K][dp]  two state slots per chain
    double* Gr;      // [2][K][dp]  grad log posterior
    double* Xi;      // [K][dp]
    double* llpart;  // [nsplit][K]
    double* gpart;   // [nsplit][K][dp]
    double* lp; double* k0; double* epsrow; int* cur;
    double* scale; long long* nsamp; long long* nacc; long long* dacc;
    double* S1; double* S2;
    // random walk / pCN (randomwalk.py:12-26, 78-100): L = chol(C) and, for pCN, its inverse, d x d row-major
    const double* PL; const double* PLinv;     // RW / pCN: L and L^-1 of the proposal covariance; HMC with a mass matrix: chM, chM^-1
    const double* Minv;                        // HMC with a mass matrix: M^-1 (NULL otherwise)
    double rho, rho_c;
    // parallel tempering (ptsampler.py:11-127): nt = 0 off; rows l*nt .. l*nt + nt - 1 form ladder l; lp = lprior + beta * ll
    int nt; double pswap;
    const double* betas;                       // [nt]
    double* llc; double* lpriorc;              // [K] log-likelihood / log-prior of the current state
    unsigned char* role;                       // [K] role in the pending step: 0 within-chain, 1 swap initiator, 2 partner
    // mMALA
    double* Gm;      // [K][d][d]   metric of the pending proposal (likelihood part)
    double* Lc;      // [2][K][d][d] Cholesky factors (current / proposal slot follows cur)
    double* logdet;  // [2][K]
    double* nat;     // [2][K][dp]   G^-1 grad
    // mMALA, RMN_PREC_TF32_METRIC (all null / 0 otherwise)
    float* W;        // [K][Npad]    p(1-p) of the pending proposal, written by lg_eval_kernel
    float* KR;       // [NP][Npad]   x_ia x_ib for a >= b, pair index a(a+1)/2 + b
    float* Gp;       // [gsplit][K][NP]  packed lower triangle of the metric (likelihood part), split-K partials
    int64_t Npad; int NP;
    int gsplit;      // partial products of the metric GEMM (1 unless the chain shard leaves most SMs without a tile)
    int kr_bf16;     // W and KR are bf16 arrays (tf32x3 mode: the metric GEMM runs as kind::f16)
};

// likelihood part of the metric entry (a, b) of chain r, from whichever representation is live
__device__ __forceinline__ double metric_entry(const LogisticState& st, int64_t r, int a, int b) {
    if (st.Gp) {
        const int hi = a > b ? a : b, lo = a > b ? b : a;
        return (double)st.Gp[r * st.NP + hi * (hi + 1) / 2 + lo];
    }
    return st.Gm[r * (int64_t)st.d * st.d + a * st.d + b];
}

// Lw[a][b] = likelihood part of chain r's metric + the prior precision on the diagonal (one warp).  The split-K partials
// of the metric GEMM are summed here, in a fixed order; the unsplit case keeps its own loop (its gathers stay
// independent of each other, so the compiler overlaps them).
__device__ __forceinline__ void fill_metric(const LogisticState& st, int64_t r, double* Lw, int d, int ldl, double pvinv, int lane) {
    if (st.Gp && st.gsplit > 1) {
        const int64_t stride = st.K * (int64_t)st.NP;
        for (int q = lane; q < d * d; q += 32) {
            const int a = q / d, b = q % d;
            const int hi = a > b ? a : b, lo = a > b ? b : a;
            const float* g = st.Gp + r * st.NP + hi * (hi + 1) / 2 + lo;
            float part[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) part[k] = (k < st.gsplit) ? g[k * stride] : 0.0f;       // gsplit <= 8 (constructor)
            double v = (double)part[0];
#pragma unroll
            for (int k = 1; k < 8; ++k) v += (double)part[k];
            Lw[a * ldl + b] = v + ((a == b) ? pvinv : 0.0);
        }
        return;
    }
    for (int q = lane; q < d * d; q += 32) {
        const int a = q / d, b = q % d;
        Lw[a * ldl + b] = metric_entry(st, r, a, b) + ((a == b) ? pvinv : 0.0);
    }
}

// ---------------------------------------------------------------------------------------
// eval: llpart[split][c], gpart[split][c][:] for the chains' PROPOSAL slots (slot = cur^1),
// or for slot `fixed_slot` when >= 0 (set_state / pointwise).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EVAL_THREADS, 1)
lg_eval_kernel(LogisticState st, int fixed_slot) {
    // 16 warps: warp = (rh, cw); cw = chain group (8 chains), rh = which 32-row half of each
    // 64-row X tile the warp contracts over.  Both halves accumulate G for the same chains and
    // are summed through shared memory once, at the end of the kernel.
    extern __shared__ __align__(16) double sm[];
    const int ldt = st.ldt, dp = st.dp, d = st.d;
    double* Ts = sm;                              // [BC][ldt]
    double* Xs = Ts + BC * ldt;                   // [2][BI][ldt]
    double* ys = Xs + 2 * BI * ldt;               // [2][BI]
    double* tab = sm + st.eval_tab_off;           // sigmoid/softplus tables, behind everything else
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int cw = warp & 7, rh = warp >> 3;
    for (int q = tid; q < lgmath::TAB_DOUBLES; q += EVAL_THREADS) tab[q] = g_lg_tab[q];
    const int64_t K = st.K;
    const int64_t c0 = (int64_t)blockIdx.x * BC;
    const int split = blockIdx.y;
    const int64_t r_begin = (int64_t)split * st.rows_per_split;
    const int64_t r_end = min(st.N, r_begin + st.rows_per_split);
    const int ntiles = (r_end > r_begin) ? (int)((r_end - r_begin + BI - 1) / BI) : 0;

    // stage Theta (zero padded) and clear the pad columns of the X buffers
    for (int q = tid; q < BC * ldt; q += EVAL_THREADS) {
        const int c = q / ldt, k = q % ldt;
        const int64_t cc = c0 + c;
        double v = 0.0;
        if (cc < K && k < d) {
            const int slot = (fixed_slot >= 0) ? fixed_slot : (st.cur[cc] ^ 1);
            v = st.Th[((int64_t)slot * K + cc) * dp + k];
        }
        Ts[q] = v;
    }
    for (int q = tid; q < 2 * BI * ldt; q += EVAL_THREADS)
        if (q % ldt >= d) Xs[q] = 0.0;

    const bool even = (d & 1) == 0;
    const int cpr = even ? d / 2 : d;             // chunks per row
    auto load_tile = [&](int tile, int buf) {
        const int64_t r0 = r_begin + (int64_t)tile * BI;
        double* xs = Xs + buf * BI * ldt;
        // 8 threads per row walk its chunks: no integer division in the copy loop
        const int row = tid >> 3, sub = tid & 7;
        const int64_t r = r0 + row;
        const bool ok = r < r_end;
        const double* src = st.X + (ok ? r : 0) * d;
        double* dst = xs + row * ldt;
        if (even) { for (int ch = sub; ch < cpr; ch += 8) cp_async16(dst + 2 * ch, src + 2 * ch, ok); }
        else      { for (int ch = sub; ch < cpr; ch += 8) cp_async8(dst + ch, src + ch, ok); }
        if (tid < BI) {
            const int64_t ry = r0 + tid;
            cp_async8(ys + buf * BI + tid, st.y + (ry < r_end ? ry : 0), ry < r_end);
        }
    };

    double accg[16][2];                           // G[c = 8*cw + g][k = 8*jn + 2t (+1)], jn < dp/8 <= 16
#pragma unroll
    for (int j = 0; j < 16; ++j) { accg[j][0] = 0.0; accg[j][1] = 0.0; }
    double ll = 0.0;
    const int nkn = dp / 8;

    if (ntiles > 0) load_tile(0, 0);
    cp_async_commit();
    for (int tile = 0; tile < ntiles; ++tile) {
        const int buf = tile & 1;
        if (tile + 1 < ntiles) load_tile(tile + 1, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const double* xs = Xs + (buf * BI + rh * 32) * ldt;       // this warp's 32 rows
        const double* yb = ys + buf * BI + rh * 32;
        const int64_t r0 = r_begin + (int64_t)tile * BI + rh * 32;

        // ---- product 1: Z^T[c][i] = sum_k Theta[c][k] X[i][k]
        double z[4][2];
#pragma unroll
        for (int j = 0; j < 4; ++j) { z[j][0] = 0.0; z[j][1] = 0.0; }
        const double* ta = Ts + (cw * 8 + g) * ldt + t;
        const double* xb = xs + perm8(g) * ldt + t;
        for (int k4 = 0; k4 < dp; k4 += 4) {
            const double a = ta[k4];
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma884(z[j][0], z[j][1], a, xb[j * 8 * ldt + k4]);
        }
        // ---- pointwise: p = sigmoid(z), r = y - p, ll += y z - softplus(z)
        //      table-driven fp64 exp / log / reciprocal sharing their range reduction (logistic_math.cuh)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int i = j * 8 + perm8(2 * t + e);
                const bool ok = (r0 + i) < r_end;
                const double yv = yb[i];
                const double zz = z[j][e];
                double p, sp, pq;
                lgmath::sigmoid_softplus(zz, tab, p, sp, pq);
                ll += ok ? (yv * zz - sp) : 0.0;
                z[j][e] = ok ? (yv - p) : 0.0;
                if (st.W && ok && c0 + cw * 8 + g < K)
                    st.W[(c0 + cw * 8 + g) * st.Npad + (r0 + i)] = (float)pq;                    // p(1-p)
            }
        }
        // ---- product 2: G[c][k] += sum_i R[c][i] X[i][k]   (contraction rows permuted, see perm8)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double* xr = xs + (j * 8 + perm8(2 * t + e)) * ldt + g;
                const double a = z[j][e];
#pragma unroll
                for (int jn = 0; jn < 16; ++jn)
                    if (jn < nkn) dmma884(accg[jn][0], accg[jn][1], a, xr[jn * 8]);
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();
    __syncthreads();

    // ---- combine the two row halves through shared memory (reuse the X buffers)
    ll += __shfl_xor_sync(0xffffffffu, ll, 1);
    ll += __shfl_xor_sync(0xffffffffu, ll, 2);
    double* red = Xs;                             // [8 warps][32 lanes][33]
    if (rh == 1) {
        double* my = red + ((size_t)cw * 32 + lane) * 33;
#pragma unroll
        for (int jn = 0; jn < 16; ++jn) { my[2 * jn] = accg[jn][0]; my[2 * jn + 1] = accg[jn][1]; }
        my[32] = ll;
    }
    __syncthreads();
    if (rh == 0) {
        const double* ot = red + ((size_t)cw * 32 + lane) * 33;
        const int64_t c = c0 + cw * 8 + g;
        if (c < K) {
            if (t == 0) st.llpart[(int64_t)split * K + c] = ll + ot[32];
            double* gp = st.gpart + ((int64_t)split * K + c) * dp;
#pragma unroll
            for (int jn = 0; jn < 16; ++jn)
                if (jn < nkn)
                    *reinterpret_cast<double2*>(gp + jn * 8 + 2 * t) =
                        make_double2(accg[jn][0] + ot[2 * jn], accg[jn][1] + ot[2 * jn + 1]);
        }
    }
}

// ---------------------------------------------------------------------------------------
// metric: Gm[c] = sum_i w_ic x_i x_i^T (likelihood part) for the PROPOSAL slot (or fixed).
// CTA = 8 warps = 4 chains x 2 warps.  G is symmetric: only the 36 8x8 tiles on or below the block
// diagonal are accumulated (18 per warp: warp 0 owns row-blocks {0,3,4,7}, warp 1 {1,2,5,6}, so both
// issue the same number of DMMAs) and the strict upper triangle is mirrored at the store.
// Each CTA loops over ALL rows (X is L2-resident at config 5: 51 MB).
// ---------------------------------------------------------------------------------------
template <int H> struct MetricTiles {
    // row-blocks of warp H and the offset of each row-block's tiles in the accumulator array
    static constexpr int rb(int q) { return H == 0 ? (q == 0 ? 0 : q == 1 ? 3 : q == 2 ? 4 : 7)
                                                   : (q == 0 ? 1 : q == 1 ? 2 : q == 2 ? 5 : 6); }
    static constexpr int off(int q) { return H == 0 ? (q == 0 ? 0 : q == 1 ? 1 : q == 2 ? 5 : 10)
                                                    : (q == 0 ? 0 : q == 1 ? 2 : q == 2 ? 5 : 11); }
};

template <int H>
__device__ __forceinline__ void metric_slab(double (&acc)[18][2], const double* __restrict__ xr, double w, int g) {
    double b[8];
#pragma unroll
    for (int jb = 0; jb < 8; ++jb) b[jb] = xr[jb * 8 + g];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int rb = MetricTiles<H>::rb(q);
        const double a = xr[rb * 8 + g] * w;
#pragma unroll
        for (int cb = 0; cb < 8; ++cb)
            if (cb <= rb) dmma884(acc[MetricTiles<H>::off(q) + cb][0], acc[MetricTiles<H>::off(q) + cb][1], a, b[cb]);
    }
}

template <int H>
__device__ __forceinline__ void metric_store(const double (&acc)[18][2], double* __restrict__ G, int d, int g, int t) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int rb = MetricTiles<H>::rb(q);
        const int a = rb * 8 + g;
#pragma unroll
        for (int cb = 0; cb < 8; ++cb) {
            if (cb <= rb) {
                const int b = cb * 8 + 2 * t;
                const double v0 = acc[MetricTiles<H>::off(q) + cb][0], v1 = acc[MetricTiles<H>::off(q) + cb][1];
                if (a < d && b < d) { G[a * d + b] = v0; if (cb < rb) G[b * d + a] = v0; }
                if (a < d && b + 1 < d) { G[a * d + b + 1] = v1; if (cb < rb) G[(b + 1) * d + a] = v1; }
            }
        }
    }
}

__global__ void __launch_bounds__(256, 2)
lg_metric_kernel(LogisticState st, int fixed_slot) {
    extern __shared__ __align__(16) double sm[];
    const int ldt = st.ldt, dp = st.dp, d = st.d;
    double* Ts = sm;                              // [4][ldt]
    double* Xs = Ts + 4 * ldt;                    // [2][BI][ldt]
    double* ws = Xs + 2 * BI * ldt;               // [4][BI]  weights of the current tile
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int ch = warp >> 1, half = warp & 1;
    const int64_t K = st.K;
    const int64_t c = (int64_t)blockIdx.x * 4 + ch;
    const int ntiles = (int)((st.N + BI - 1) / BI);

    for (int q = tid; q < 4 * ldt; q += 256) {
        const int cc = q / ldt, k = q % ldt;
        const int64_t cg = (int64_t)blockIdx.x * 4 + cc;
        double v = 0.0;
        if (cg < K && k < d) {
            const int slot = (fixed_slot >= 0) ? fixed_slot : (st.cur[cg] ^ 1);
            v = st.Th[((int64_t)slot * K + cg) * dp + k];
        }
        Ts[q] = v;
    }
    for (int q = tid; q < 2 * BI * ldt; q += 256)
        if (q % ldt >= d) Xs[q] = 0.0;
    const bool even = (d & 1) == 0;
    const int cpr = even ? d / 2 : d;
    auto load_tile = [&](int tile, int buf) {
        const int64_t r0 = (int64_t)tile * BI;
        double* xs = Xs + buf * BI * ldt;
        for (int q = tid; q < BI * cpr; q += 256) {
            const int row = q / cpr, chk = q % cpr;
            const int64_t r = r0 + row;
            const bool ok = r < st.N;
            const double* src = st.X + (ok ? r : 0) * d + (even ? 2 * chk : chk);
            if (even) cp_async16(xs + row * ldt + 2 * chk, src, ok);
            else cp_async8(xs + row * ldt + chk, src, ok);
        }
    };

    double acc[18][2];                            // this warp's 18 lower-triangle tiles (MetricTiles)
#pragma unroll
    for (int i = 0; i < 18; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }

    load_tile(0, 0);
    cp_async_commit();
    for (int tile = 0; tile < ntiles; ++tile) {
        const int buf = tile & 1;
        if (tile + 1 < ntiles) load_tile(tile + 1, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const double* xs = Xs + buf * BI * ldt;
        // weights: thread q handles (chain q/64, row q%64): z = x_i . theta, w = p (1 - p)
        {
            const int cc = tid >> 6, i = tid & 63;
            const double* xr = xs + i * ldt;
            const double* th = Ts + cc * ldt;
            double zz = 0.0;
            for (int k = 0; k < d; ++k) zz += xr[k] * th[k];
            const double ex = exp(-fabs(zz));
            const double inv = 1.0 / (1.0 + ex);
            const bool ok = ((int64_t)tile * BI + i) < st.N;
            ws[cc * BI + i] = ok ? (ex * inv) * inv : 0.0;       // p(1-p) = e/(1+e)^2
        }
        __syncthreads();
        const double* wc = ws + ch * BI;
#pragma unroll 2
        for (int i4 = 0; i4 < BI; i4 += 4) {
            // A[a][i] = x_ia w_i  (thread: a = 8*rb + g, i = i4 + t);  B[i][b] = x_ib (i = i4 + t, b = 8*cb + g)
            const double* xr = xs + (i4 + t) * ldt;
            const double w = wc[i4 + t];
            if (half == 0) metric_slab<0>(acc, xr, w, g);
            else metric_slab<1>(acc, xr, w, g);
        }
        __syncthreads();
    }
    cp_async_wait<0>();
    if (c < K) {
        double* G = st.Gm + c * (int64_t)d * d;
        if (half == 0) metric_store<0>(acc, G, d, g, t);
        else metric_store<1>(acc, G, d, g, t);
    }
}

// Khatri-Rao table for the tensor-core metric: KR[a(a+1)/2 + b][i] = x_ia x_ib (fp32), a >= b.
// A block stages 32 data rows in shared memory; each warp then writes one 128-byte line per pair.
__global__ void __launch_bounds__(256)
lg_build_kr_kernel(LogisticState st) {
    extern __shared__ __align__(16) double sm[];                  // [32][d + 1]
    const int d = st.d, ld = d + 1;
    const int64_t i0 = (int64_t)blockIdx.x * 32;
    for (int q = threadIdx.x; q < 32 * d; q += blockDim.x) {
        const int r = q / d, k = q % d;
        sm[r * ld + k] = (i0 + r < st.N) ? st.X[(i0 + r) * d + k] : 0.0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int npair = d * (d + 1) / 2;
    if (i0 + lane >= st.Npad) return;
    const double* xr = sm + lane * ld;
    int a = 0, base = 0;                                           // base = a(a+1)/2
    for (int pr = warp; pr < npair; pr += nw) {
        while (pr >= base + a + 1) { base += a + 1; ++a; }
        const int b = pr - base;
        const float v = (float)(xr[a] * xr[b]);
        if (st.kr_bf16) reinterpret_cast<__nv_bfloat16*>(st.KR)[(int64_t)pr * st.Npad + i0 + lane] = __float2bfloat16_rn(v);
        else st.KR[(int64_t)pr * st.Npad + i0 + lane] = v;
    }
}

// Row-sharded data mode: fold the row splits of this rank into slot 0 (same fixed order the finish kernels use), so
// ONE contiguous block per quantity goes through the all-reduce over ranks; afterwards slot 0 holds the sum over all
// rows of all ranks and the finish kernels run with nsplit = 1.
__global__ void __launch_bounds__(256)
lg_fold_splits_kernel(LogisticState st) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t n = st.K * st.dp;
    if (i < n) {
        double g = 0.0;
        for (int s = 0; s < st.nsplit; ++s) g += st.gpart[(int64_t)s * n + i];
        st.gpart[i] = g;
    }
    if (i < st.K) {
        double ll = 0.0;
        for (int s = 0; s < st.nsplit; ++s) ll += st.llpart[(int64_t)s * st.K + i];
        st.llpart[i] = ll;
    }
}

// Interior leapfrog step (hamiltonian.py:33-37, Nsteps > 1) for the chains' proposal slots: with the gradient of
// the log-posterior at the trajectory point just swept, g = sum_splits gpart - theta / pv,
//   p <- p + eps g,   theta <- theta + eps p      (in place; p lives in Xi)
__global__ void __launch_bounds__(256)
lg_leapfrog_mid_kernel(LogisticState st) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= st.K * st.dp) return;
    const int64_t r = i / st.dp;
    const int j = (int)(i % st.dp);
    if (j >= st.d) return;
    const double eps = st.epsrow[r];
    double* th = st.Th + ((int64_t)(st.cur[r] ^ 1) * st.K + r) * st.dp + j;
    double g = 0.0;
    for (int s = 0; s < st.nsplit; ++s) g += st.gpart[((int64_t)s * st.K + r) * st.dp + j];
    g -= *th / st.pv;
    const double p = st.Xi[i] + eps * g;
    st.Xi[i] = p;
    *th = *th + eps * p;
}

// The same step with a mass matrix: p <- p + eps g, theta <- theta + eps solve(M, p); one warp per chain.
__global__ void __launch_bounds__(128)
lg_leapfrog_mid_mass_kernel(LogisticState st) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (r >= st.K) return;
    const int d = st.d, dp = st.dp;
    const double eps = st.epsrow[r];
    double* th = st.Th + ((int64_t)(st.cur[r] ^ 1) * st.K + r) * dp;
    double* p = st.Xi + r * dp;
    for (int j = lane; j < d; j += 32) {
        double g = 0.0;
        for (int s = 0; s < st.nsplit; ++s) g += st.gpart[((int64_t)s * st.K + r) * dp + j];
        g -= th[j] / st.pv;
        p[j] = p[j] + eps * g;
    }
    __syncwarp();
    for (int j = lane; j < d; j += 32) {
        double sacc = 0.0;
        for (int i = 0; i < d; ++i) sacc += st.Minv[(size_t)j * d + i] * p[i];
        th[j] = th[j] + eps * sacc;
    }
}

// ---------------------------------------------------------------------------------------
// finish / propose, one warp per chain.
// ---------------------------------------------------------------------------------------
enum { LG_HMC = 0, LG_RW = 1, LG_PCN = 2 };   // the non-mMALA proposals of lg_finish_propose_kernel<false>
struct LgStep {
    int mmala, adapt, finish, propose, diag;
    int record;             // trace / diagnostics of the state after the finished step (= finish, except under tempering)
    const double* inj_usel; // tempering: injected selection uniforms of the step being proposed
    int pkind;
    double target, eps0;
    uint64_t seed; int64_t chain_offset, step_fin, step_prop;
    const double* inj_xi; const double* inj_u;
    int64_t trace_slot;
    double* tr_theta; double* tr_logpost; double* tr_prop_lp; uint8_t* tr_acc;
    double* tr_lqr; double* tr_prop_theta;
};

// In-place Cholesky of the d x d matrix A (row-major, leading dimension ld) held in shared memory,
// executed by one warp.  Returns log det = 2 sum log L_jj; NaN if not positive definite.
// ld = d + 1 (LDL below): the column sweeps read A[i * ld + k] with the LANE in i; with ld = d = 64 every lane hit the
// same bank (a 32-way conflict on every load of the O(d^3) loop), with an odd ld the lanes spread over all banks.
__device__ __forceinline__ int LDL(int d) { return d + 1; }
__device__ double warp_cholesky(double* A, int d, int ld, int lane) {
    double logdet = 0.0;
    for (int j = 0; j < d; ++j) {
        double s = 0.0;
        for (int k = lane; k < j; k += 32) s += A[j * ld + k] * A[j * ld + k];
        s = group_sum<32>(s);
        const double djj = A[j * ld + j] - s;
        const double ljj = sqrt(djj);
        logdet += 2.0 * log(ljj);
        __syncwarp();
        if (lane == 0) A[j * ld + j] = ljj;
        for (int i = j + 1 + lane; i < d; i += 32) {
            double v = A[i * ld + j];
            for (int k = 0; k < j; ++k) v -= A[i * ld + k] * A[j * ld + k];
            A[i * ld + j] = v / ljj;
        }
        __syncwarp();
    }
    return logdet;
}
// x <- L^-1 x (forward) ; x in shared memory
__device__ void warp_trsv_lower(const double* L, double* x, int d, int ld, int lane) {
    for (int j = 0; j < d; ++j) {
        __syncwarp();
        const double xj = x[j] / L[j * ld + j];
        __syncwarp();
        if (lane == 0) x[j] = xj;
        for (int i = j + 1 + lane; i < d; i += 32) x[i] -= L[i * ld + j] * xj;
    }
    __syncwarp();
}
// x <- L^-T x (backward)
__device__ void warp_trsv_lower_t(const double* L, double* x, int d, int ld, int lane) {
    for (int j = d - 1; j >= 0; --j) {
        __syncwarp();
        const double xj = x[j] / L[j * ld + j];
        __syncwarp();
        if (lane == 0) x[j] = xj;
        for (int i = lane; i < j; i += 32) x[i] -= L[j * ld + i] * xj;
    }
    __syncwarp();
}

template <bool MMALA>
__global__ void __launch_bounds__(128)
lg_finish_propose_kernel(LogisticState st, LgStep sp) {
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (r >= st.K) return;
    const int dp = st.dp, d = st.d;
    const int64_t K = st.K;
    const int ldl = LDL(d);
    double* Lw = sm + (size_t)wib * (MMALA ? (d * ldl + 2 * dp) : 0);  // [d][ldl] + two vectors
    double* v1 = Lw + d * ldl;
    double* v2 = v1 + dp;
    int c = st.cur[r];
    double lp = st.lp[r];
    const RngKey rk(sp.seed, (uint64_t)(sp.chain_offset + r));
    const double pvinv = 1.0 / st.pv;

    const bool pt = !MMALA && st.nt > 0;
    const int ti = pt ? (int)((sp.chain_offset + r) % st.nt) : 0;
    const double beta = pt ? st.betas[ti] : 1.0;
    const int role = (pt && sp.finish) ? st.role[r] : 0;
    if (sp.finish && role == 1) {
        // ---- swap proposal with the next rung (ptsampler.py:113-125); this warp moves BOTH rows.  TemperedModel scales
        //      the likelihood only (:33-37): log-posterior of model i at theta_j = lprior_j + beta_i ll_j
        const int64_t rj = r + 1;
        const double beta_j = st.betas[ti + 1];
        const double ll_i = st.llc[r], ll_j = st.llc[rj], pr_i = st.lpriorc[r], pr_j = st.lpriorc[rj], lp_j = st.lp[rj];
        const double lp_ij = combine_logpost(pr_j, ll_j * beta);
        const double lp_ji = combine_logpost(pr_i, ll_i * beta_j);
        const double x = exp((lp_ji + lp_ij) - (lp + lp_j));
        const double mhr = (x < 1.0) ? x : 1.0;                        // Python min(1, x): nan -> 1
        const double u = sp.inj_u ? sp.inj_u[r] : u01(rk.block((uint64_t)sp.step_fin, RMN_BLOCK_ACCEPT).x);
        const bool sw = u < mhr;                                       // :121
        if (sp.tr_acc && lane == 0) { sp.tr_acc[r] = sw ? 1 : 0; sp.tr_acc[rj] = sw ? 1 : 0; }
        if (sp.tr_prop_lp && lane == 0) { sp.tr_prop_lp[r] = lp_ij; sp.tr_prop_lp[rj] = lp_ji; }
        if (sw) {
            const int cj = st.cur[rj];
            double* ti_ = st.Th + ((int64_t)c * K + r) * dp;  double* tj_ = st.Th + ((int64_t)cj * K + rj) * dp;
            double* gi_ = st.Gr + ((int64_t)c * K + r) * dp;  double* gj_ = st.Gr + ((int64_t)cj * K + rj) * dp;
            for (int j = lane; j < dp; j += 32) {
                const double a = ti_[j], b = tj_[j]; ti_[j] = b; tj_[j] = a;
                const double e = gi_[j], f = gj_[j]; gi_[j] = f; gj_[j] = e;
            }
            if (lane == 0) {
                st.llc[r] = ll_j; st.llc[rj] = ll_i; st.lpriorc[r] = pr_j; st.lpriorc[rj] = pr_i;
                st.lp[r] = lp_ij; st.lp[rj] = lp_ji;
            }
        }
    } else if (sp.finish && role == 0) {
        const int pslot = c ^ 1;
        const double* thp = st.Th + ((int64_t)pslot * K + r) * dp;
        const double* thc = st.Th + ((int64_t)c * K + r) * dp;
        double* grp = st.Gr + ((int64_t)pslot * K + r) * dp;
        const double* grc = st.Gr + ((int64_t)c * K + r) * dp;
        const double* xi = st.Xi + r * dp;
        const double eps = st.epsrow[r];
        double ll = 0.0;
        for (int s = 0; s < st.nsplit; ++s) ll += st.llpart[(int64_t)s * K + r];   // fixed order
        double tt = 0.0, k1 = 0.0;
        for (int j = lane; j < dp; j += 32) {
            double gsum = 0.0;
            for (int s = 0; s < st.nsplit; ++s) gsum += st.gpart[((int64_t)s * K + r) * dp + j];
            const double th = thp[j];
            const double gp = (j < d) ? gsum - th * pvinv : 0.0;
            grp[j] = gp;
            tt += th * th;
            if (!MMALA && sp.pkind == LG_HMC) {
                const double p1 = xi[j] + 0.5 * eps * gp;                // final half step (hamiltonian.py:40); Xi holds p
                if (st.Minv) st.Xi[r * dp + j] = (j < d) ? p1 : 0.0;     // staged for solve(chM, p) below
                else k1 += p1 * p1;
            }
        }
        if (!MMALA && sp.pkind == LG_HMC && st.Minv) {
            // kinetic energy in the mass metric: |solve(chM, p')|^2 (hamiltonian.py:86-89); PLinv = chM^-1, lower triangular
            const double* v = st.Xi + r * dp;
            __syncwarp();
            for (int j = lane; j < d; j += 32) {
                double sacc = 0.0;
                for (int i = 0; i <= j; ++i) sacc += st.PLinv[(size_t)j * d + i] * v[i];
                k1 += sacc * sacc;
            }
            __syncwarp();
        }
        if (!MMALA && sp.pkind == LG_PCN) {
            // u_rev = (rho_c L)^-1 (theta - rho theta')  (randomwalk.py:95,98); |u_fwd|^2 = |xi|^2 = k0
            double* v = st.Xi + r * dp;                                  // the noise is no longer needed: staging
            __syncwarp();
            for (int j = lane; j < d; j += 32) v[j] = thc[j] - st.rho * thp[j];
            __syncwarp();
            for (int j = lane; j < d; j += 32) {
                double sacc = 0.0;
                for (int i = 0; i <= j; ++i) sacc += st.PLinv[(size_t)j * d + i] * v[i];
                sacc /= st.rho_c;
                k1 += sacc * sacc;
            }
        }
        tt = group_sum<32>(tt);
        k1 = group_sum<32>(k1);
        const double lprior = -0.5 * tt * pvinv - 0.5 * (double)d * log(2.0 * M_PI * st.pv);
        const double lpn = combine_logpost(lprior, pt ? ll * beta : ll);    // TemperedModel (ptsampler.py:33-34)
        double lqr;
        if (!MMALA) {
            if (sp.pkind == LG_HMC) lqr = 0.5 * (k1 - st.k0[r]);          // hamiltonian.py:89
            else if (sp.pkind == LG_PCN) lqr = -0.5 * (st.k0[r] - k1);    // randomwalk.py:99-100
            else lqr = 0.0;                                               // randomwalk.py:26
        } else {
            // geometry of the proposal: L' = chol(G'), logdet', nat' = G'^-1 grad'
            fill_metric(st, r, Lw, d, ldl, pvinv, lane);
            __syncwarp();
            const double ld = warp_cholesky(Lw, d, ldl, lane);
            for (int j = lane; j < d; j += 32) v1[j] = grp[j];
            __syncwarp();
            warp_trsv_lower(Lw, v1, d, ldl, lane);
            warp_trsv_lower_t(Lw, v1, d, ldl, lane);                      // v1 = nat'
            // reverse residual  r = L'^T (theta - mean'),  mean' = theta' + eps^2/2 nat'
            for (int j = lane; j < d; j += 32) v2[j] = thc[j] - (thp[j] + 0.5 * eps * eps * v1[j]);
            __syncwarp();
            double rr = 0.0;
            for (int j = lane; j < d; j += 32) {
                double s = 0.0;
                for (int i = j; i < d; ++i) s += Lw[i * ldl + j] * v2[i];  // (L^T v)_j
                rr += s * s;
            }
            rr = group_sum<32>(rr);
            const double ldc = st.logdet[(int64_t)c * K + r];
            const double cst = 0.5 * (double)d * log(2.0 * M_PI * eps * eps);
            const double lq_fwd = 0.5 * ldc - cst - 0.5 * st.k0[r];      // |L^T(theta'-mean)|^2 = eps^2 |xi|^2
            const double lq_rev = 0.5 * ld - cst - 0.5 * rr / (eps * eps);
            lqr = lq_fwd - lq_rev;
            // stash the proposal's geometry in its slot
            double* Lp = st.Lc + ((int64_t)pslot * K + r) * d * d;
            for (int q = lane; q < d * d; q += 32) Lp[q] = Lw[(q / d) * ldl + q % d];
            double* np_ = st.nat + ((int64_t)pslot * K + r) * dp;
            for (int j = lane; j < d; j += 32) np_[j] = v1[j];
            if (lane == 0) st.logdet[(int64_t)pslot * K + r] = ld;
        }
        const double u = sp.inj_u ? sp.inj_u[r] : u01(rk.block((uint64_t)sp.step_fin, RMN_BLOCK_ACCEPT).x);
        const bool acc = mh_accept(lpn, lp, lqr, u);
        if (sp.tr_prop_theta)
            for (int j = lane; j < d; j += 32) sp.tr_prop_theta[r * d + j] = thp[j];
        if (sp.tr_lqr && lane == 0) sp.tr_lqr[r] = lqr;
        if (acc) { c ^= 1; lp = lpn; }
        if (lane == 0) {
            if (acc) { st.cur[r] = c; st.lp[r] = lp; if (pt) { st.llc[r] = ll; st.lpriorc[r] = lprior; } }
            st.dacc[r] += acc ? 1 : 0;
            if (sp.adapt) {
                AdaptState ad{st.scale[r], st.nsamp[r], st.nacc[r]};
                ad.update(acc, sp.target);
                st.scale[r] = ad.scale; st.nsamp[r] = ad.nsamples; st.nacc[r] = ad.naccepts;
            }
            if (sp.tr_prop_lp) sp.tr_prop_lp[r] = lpn;
            if (sp.tr_acc) sp.tr_acc[r] = acc ? 1 : 0;
            if (!pt && sp.trace_slot >= 0 && sp.tr_logpost) sp.tr_logpost[sp.trace_slot * K + r] = lp;
        }
        __syncwarp();
    }
    if (!sp.propose && !sp.record) return;           // tempering: the finish-only launch ends here
    if (pt && sp.record && lane == 0 && sp.trace_slot >= 0 && sp.tr_logpost) sp.tr_logpost[sp.trace_slot * K + r] = lp;
    // tempering: roles of the step being proposed (ptsampler.py:102-112), sequential along the ladder
    int role_next = 0;
    if (pt && sp.propose) {
        const int64_t r0 = r - ti;
        double usel = 1.0;
        if (lane < st.nt) {
            usel = sp.inj_usel ? sp.inj_usel[r0 + lane]
                               : u01(RngKey(sp.seed, (uint64_t)(sp.chain_offset + r0 + lane)).block((uint64_t)sp.step_prop, RMN_BLOCK_AUX).x);
        }
        const unsigned wmask = __ballot_sync(0xffffffffu, lane < st.nt - 1 && !(usel > st.pswap));
        unsigned init = 0;
        for (int b = 0; b < st.nt - 1; ++b)
            if (((wmask >> b) & 1u) && !(b > 0 && ((init >> (b - 1)) & 1u))) init |= 1u << b;
        role_next = ((init >> ti) & 1u) ? 1 : ((ti > 0 && ((init >> (ti - 1)) & 1u)) ? 2 : 0);
        if (lane == 0) st.role[r] = (unsigned char)role_next;
    }
    const bool do_propose = sp.propose && role_next == 0;    // swap rows draw no proposal

    const double* th = st.Th + ((int64_t)c * K + r) * dp;
    const double* gr = st.Gr + ((int64_t)c * K + r) * dp;
    double* thn = st.Th + ((int64_t)(c ^ 1) * K + r) * dp;
    double* xo = st.Xi + r * dp;
    const double scale = sp.adapt ? st.scale[r] : 1.0;
    const double eps = scale * sp.eps0;                                  // RW: eps0 = 1, so eps is the scale (randomwalk.py:26)
    const bool want_trace = sp.record && sp.trace_slot >= 0 && sp.tr_theta;
    double k0 = 0.0, rowsum = 0.0;

    if (MMALA && sp.propose) {
        const double* Lcur = st.Lc + ((int64_t)c * K + r) * d * d;
        for (int q = lane; q < d * d; q += 32) Lw[(q / d) * ldl + q % d] = Lcur[q];
    }
    for (int j4 = lane * 4; j4 < dp; j4 += 128) {
        double xi[4];
        if (do_propose) {
            if (sp.inj_xi) {
#pragma unroll
                for (int q = 0; q < 4; ++q) xi[q] = (j4 + q < d) ? sp.inj_xi[r * d + j4 + q] : 0.0;
            } else {
                normal4(rk.block((uint64_t)sp.step_prop, (uint32_t)(j4 >> 2)), xi);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (j4 + q >= d) xi[q] = 0.0;
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = j4 + q;
            const double tv = th[j];
            rowsum += tv;
            if (want_trace && j < d) sp.tr_theta[(sp.trace_slot * K + r) * d + j] = tv;
            if (!do_propose) continue;
            k0 += xi[q] * xi[q];
            if (!MMALA && (sp.pkind != LG_HMC || st.Minv)) {
                xo[j] = xi[q];     // staged for the L xi (chM xi) product below
            } else if (!MMALA) {
                const double ph = xi[q] + 0.5 * eps * gr[j];             // hamiltonian.py:27
                thn[j] = tv + eps * ph;                                  // :30
                xo[j] = ph;        // the momentum rides in Xi through the trajectory (Nsteps >= 1)
            } else {
                xo[j] = xi[q];
                v1[j] = xi[q];
            }
        }
    }
    if (!MMALA && do_propose && sp.pkind != LG_HMC) {
        // theta' = theta + scale L xi (randomwalk.py:25-26) or rho theta + rho_c L xi (:93), L lower triangular
        __syncwarp();
        const double a0 = (sp.pkind == LG_PCN) ? st.rho : 1.0, a1 = (sp.pkind == LG_PCN) ? st.rho_c : eps;
        for (int j = lane; j < dp; j += 32) {
            double sacc = 0.0;
            if (j < d)
                for (int i = 0; i <= j; ++i) sacc += st.PL[(size_t)j * d + i] * xo[i];
            thn[j] = (j < d) ? a0 * th[j] + a1 * sacc : 0.0;
        }
    }
    if (!MMALA && do_propose && sp.pkind == LG_HMC && st.Minv) {
        // VanillaHMC with a mass matrix (hamiltonian.py:79-88, leapfrog :26-30):
        //   p0 = chM xi,  k0 = |solve(chM, p0)|^2,  p_half = p0 + eps/2 grad,  theta' = theta + eps solve(M, p_half)
        // staged through the chain's own rows: xi in Xi, p0 in the proposal slot, then p_half in Xi
        __syncwarp();
        for (int j = lane; j < dp; j += 32) {
            double sacc = 0.0;
            if (j < d)
                for (int i = 0; i <= j; ++i) sacc += st.PL[(size_t)j * d + i] * xo[i];
            thn[j] = sacc;
        }
        __syncwarp();
        k0 = 0.0;
        for (int j = lane; j < dp; j += 32) {
            double sacc = 0.0;
            if (j < d)
                for (int i = 0; i <= j; ++i) sacc += st.PLinv[(size_t)j * d + i] * thn[i];
            k0 += sacc * sacc;
            xo[j] = (j < d) ? thn[j] + 0.5 * eps * gr[j] : 0.0;          // the momentum rides in Xi through the trajectory
        }
        __syncwarp();
        for (int j = lane; j < dp; j += 32) {
            double sacc = 0.0;
            if (j < d)
                for (int i = 0; i < d; ++i) sacc += st.Minv[(size_t)j * d + i] * xo[i];
            thn[j] = (j < d) ? th[j] + eps * sacc : 0.0;
        }
    }
    if (MMALA && sp.propose) {
        __syncwarp();
        warp_trsv_lower_t(Lw, v1, d, ldl, lane);                         // v1 = L^-T xi
        const double* nc = st.nat + ((int64_t)c * K + r) * dp;
        for (int j = lane; j < dp; j += 32)
            thn[j] = (j < d) ? (th[j] + 0.5 * eps * eps * nc[j]) + eps * v1[j] : 0.0;
    }
    if (do_propose) {
        k0 = group_sum<32>(k0);
        if (lane == 0) { st.k0[r] = k0; st.epsrow[r] = eps; }
    }
    if (sp.diag) {
        rowsum = group_sum<32>(rowsum);
        const int nd = min(d, ND_MAX - 1) + 1;
        if (lane < nd) {
            const double f = (lane == nd - 1) ? rowsum / (double)d : th[lane];
            st.S1[(int64_t)lane * K + r] += f;
            st.S2[(int64_t)lane * K + r] += f * f;
        }
    }
}

// set_state helpers: stage into slot 1, evaluate with fixed_slot = 1, then adopt
__global__ void lg_set_kernel(LogisticState st, const double* __restrict__ theta) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= st.K * st.dp) return;
    const int64_t r = i / st.dp;
    const int j = (int)(i % st.dp);
    st.Th[((int64_t)1 * st.K + r) * st.dp + j] = (j < st.d) ? theta[r * st.d + j] : 0.0;
    st.Xi[i] = 0.0;
    if (j == 0) { st.cur[r] = 0; st.k0[r] = 0.0; st.epsrow[r] = 0.0; }
}
template <bool MMALA>
__global__ void __launch_bounds__(128)
lg_adopt_kernel(LogisticState st) {
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (r >= st.K) return;
    const int dp = st.dp, d = st.d;
    const int64_t K = st.K;
    const double pvinv = 1.0 / st.pv;
    const double* th = st.Th + ((int64_t)1 * K + r) * dp;
    double* gr = st.Gr + ((int64_t)1 * K + r) * dp;
    double ll = 0.0;
    for (int s = 0; s < st.nsplit; ++s) ll += st.llpart[(int64_t)s * K + r];
    double tt = 0.0;
    for (int j = lane; j < dp; j += 32) {
        double gsum = 0.0;
        for (int s = 0; s < st.nsplit; ++s) gsum += st.gpart[((int64_t)s * K + r) * dp + j];
        gr[j] = (j < d) ? gsum - th[j] * pvinv : 0.0;
        tt += th[j] * th[j];
    }
    tt = group_sum<32>(tt);
    if (MMALA) {
        const int ldl = LDL(d);
        double* Lw = sm + (size_t)wib * (d * ldl + dp);
        double* v1 = Lw + d * ldl;
        fill_metric(st, r, Lw, d, ldl, pvinv, lane);
        __syncwarp();
        const double ld = warp_cholesky(Lw, d, ldl, lane);
        for (int j = lane; j < d; j += 32) v1[j] = gr[j];
        __syncwarp();
        warp_trsv_lower(Lw, v1, d, ldl, lane);
        warp_trsv_lower_t(Lw, v1, d, ldl, lane);
        double* Lp = st.Lc + ((int64_t)1 * K + r) * d * d;
        for (int q = lane; q < d * d; q += 32) Lp[q] = Lw[(q / d) * ldl + q % d];
        double* np_ = st.nat + ((int64_t)1 * K + r) * dp;
        for (int j = lane; j < d; j += 32) np_[j] = v1[j];
        if (lane == 0) st.logdet[(int64_t)1 * K + r] = ld;
    }
    if (lane == 0) {
        const double lprior = -0.5 * tt * pvinv - 0.5 * (double)d * log(2.0 * M_PI * st.pv);
        if (!MMALA && st.nt > 0) {
            st.llc[r] = ll; st.lpriorc[r] = lprior;
            st.lp[r] = combine_logpost(lprior, ll * st.betas[r % st.nt]);     // chain_offset is a multiple of nt
        } else {
            st.lp[r] = combine_logpost(lprior, ll);
        }
        st.cur[r] = 1;
    }
}
__global__ void lg_get_kernel(LogisticState st, double* theta, double* lp) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= st.K * st.d) return;
    const int64_t r = i / st.d;
    const int j = (int)(i % st.d);
    if (theta) theta[i] = st.Th[((int64_t)st.cur[r] * st.K + r) * st.dp + j];
    if (lp && j == 0) lp[r] = st.lp[r];
}
__global__ void lg_get_adapt_kernel(LogisticState st, double* scale, int64_t* ns, int64_t* na) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= st.K) return;
    if (scale) scale[c] = st.scale[c];
    if (ns) ns[c] = st.nsamp[c];
    if (na) na[c] = st.nacc[c];
}
// pointwise outputs from a staged evaluation (slot 1)
__global__ void lg_point_out_kernel(LogisticState st, int which, double* out, double* grad, double* metric) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (r >= st.K) return;
    const int dp = st.dp, d = st.d;
    const double pvinv = 1.0 / st.pv;
    const double* th = st.Th + ((int64_t)1 * st.K + r) * dp;
    double ll = 0.0;
    for (int s = 0; s < st.nsplit; ++s) ll += st.llpart[(int64_t)s * st.K + r];
    double tt = 0.0;
    for (int j = lane; j < d; j += 32) {
        double gsum = 0.0;
        for (int s = 0; s < st.nsplit; ++s) gsum += st.gpart[((int64_t)s * st.K + r) * dp + j];
        if (grad) grad[r * d + j] = gsum - th[j] * pvinv;
        tt += th[j] * th[j];
    }
    tt = group_sum<32>(tt);
    const double lprior = -0.5 * tt * pvinv - 0.5 * (double)d * log(2.0 * M_PI * st.pv);
    if (out && lane == 0) out[r] = (which == 1) ? ll : ((which == 2) ? lprior : combine_logpost(lprior, ll));
    if (metric) {
        const double* Gm = st.Gm + r * (int64_t)d * d;
        for (int q = lane; q < d * d; q += 32) metric[r * (int64_t)d * d + q] = Gm[q] + ((q / d == q % d) ? pvinv : 0.0);
    }
}

// dynamic shared memory of lg_eval_kernel: Theta + two X tiles + y, and at least the
// [8][32][33] combine buffer that overlays the X tiles at the end
static size_t eval_tab_offset(int ldt) {
    const size_t tiles = (size_t)2 * BI * ldt + 2 * BI;
    const size_t red = (size_t)8 * 32 * 33;
    return ((size_t)BC * ldt + (tiles > red ? tiles : red) + 1) & ~(size_t)1;     // 16-byte aligned
}
static size_t eval_smem_bytes(int ldt) { return (eval_tab_offset(ldt) + lgmath::TAB_DOUBLES) * 8; }

static int choose_nsplit(int64_t K, int64_t N) {
    const int64_t nblocks = (K + BC - 1) / BC;
    int ns = (int)(148 / nblocks);
    if (ns < 1) ns = 1;
    const int64_t max_by_rows = (N + 4 * BI - 1) / (4 * BI);
    if (ns > max_by_rows) ns = (int)max_by_rows;
    if (ns > 64) ns = 64;
    return ns;
}

static void fill_geometry(LogisticState& st, const rmn_model* m, int64_t K) {
    st.K = K; st.N = m->N; st.d = m->d;
    st.dp = (m->d + 15) / 16 * 16;
    st.ldt = st.dp + 4;
    st.nsplit = choose_nsplit(K, m->N);
    const int64_t per = (m->N + st.nsplit - 1) / st.nsplit;
    st.rows_per_split = (per + BI - 1) / BI * BI;
    st.pv = m->prior_var;
    st.X = m->d_X; st.y = m->d_y;
    st.eval_tab_off = (int)eval_tab_offset(st.ldt);
}

struct LogisticSampler : SamplerImpl {
    rmn_sampler* s;
    LogisticState st{};
    bool mmala;
    bool tf32m;                 // tcgen05 metric GEMM instead of lg_metric_kernel (TF32_METRIC and TF32X3)
    bool bf16m = false;         // ... with bf16 operands (TF32X3)
    bool tcx3;                  // RMN_PREC_TF32X3: the likelihood sweep on tcgen05 (fused kernel, logistic_fused.cu)
    tc::GemmMaps maps;          // metric GEMM
    // RMN_PREC_TF32X3: the fused sweep (logistic_fused.cu) covers every d this family supports (d <= 128)
    bool fused = false;
    int pkind = LG_HMC;
    double* d_PL = nullptr; double* d_PLinv = nullptr; double* d_Minv = nullptr;
    lgf::Geometry fg{};
    lgf::Maps fmaps{};
    float* fXh = nullptr; float* fXl = nullptr; uint32_t* fys = nullptr; double* fllp = nullptr; float* fgp = nullptr;
    double* d_betas = nullptr;
    int set_tempering(int nt, const double* betas, double pswap) override {
        const rmn_proposal* pr = s->prop;
        RMN_REQUIRE(nt >= 2 && nt <= 32 && betas, "set_tempering: need 2 <= nt <= 32 temperatures");
        RMN_REQUIRE(pswap > 0.0 && pswap < 1.0, "Pswap must be a number between 0 and 1");
        RMN_REQUIRE(s->K % nt == 0, "set_tempering: the number of chains (%lld) must be a multiple of nt = %d", (long long)s->K, nt);
        RMN_REQUIRE(s->chain_offset % nt == 0, "set_tempering: chain_offset must be a multiple of nt");
        RMN_REQUIRE(!mmala && pkind != LG_HMC, "parallel tempering on the logistic model: RW or pCN proposals (the reference shares one "
                                               "proposal between the temperatures and does not temper the HMC gradient, ptsampler.py:81)");
        RMN_REQUIRE(!pr->adapt, "parallel tempering supports non-adaptive proposals only (ptsampler.py:81)");
        for (int i = 0; i < nt; ++i) RMN_REQUIRE(betas[i] >= 0.0 && betas[i] <= 1.0, "beta = %g must be a number between 0 and 1", betas[i]);
        if (!d_betas) RMN_CUDA(cudaMalloc(&d_betas, 32 * 8));
        RMN_CUDA(cudaMemcpy(d_betas, betas, (size_t)nt * 8, cudaMemcpyHostToDevice));
        if (!st.llc) {
            RMN_CUDA(cudaMalloc(&st.llc, (size_t)st.K * 8));
            RMN_CUDA(cudaMalloc(&st.lpriorc, (size_t)st.K * 8));
            RMN_CUDA(cudaMalloc(&st.role, (size_t)st.K));
            RMN_CUDA(cudaMemset(st.role, 0, (size_t)st.K));
        }
        st.betas = d_betas; st.nt = nt; st.pswap = pswap;
        return RMN_OK;
    }
    RowComm rowc;               // row-sharded data mode: the ranks that hold the other slices of X
    ~LogisticSampler() override {
        rmn_rowcomm_destroy(&rowc); cudaFree(d_PL); cudaFree(d_PLinv); cudaFree(d_Minv);
        cudaFree(d_betas); cudaFree(st.llc); cudaFree(st.lpriorc); cudaFree(st.role);
    }
    int set_row_comm(const void* id, size_t nbytes, int rank, int world) override {
        if (tcx3 || tf32m) {
            rmn_set_error("row-sharded data mode runs in f64 precision");
            return RMN_ERR_UNSUPPORTED;
        }
        return rmn_rowcomm_init(&rowc, id, nbytes, rank, world);
    }
    // the state the finish / adopt / leapfrog kernels see: after the exchange the sums live in slot 0
    LogisticState fst() const { LogisticState f = st; if (rowc.comm) f.nsplit = 1; return f; }
    int exchange(cudaStream_t stream) {
        if (!rowc.comm) return RMN_OK;
        const int64_t n = st.K * st.dp;
        lg_fold_splits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(st);
        RMN_KERNEL_CHECK(); launches++;
        double* bufs[3] = {st.llpart, st.gpart, st.Gm};
        const size_t counts[3] = {(size_t)st.K, (size_t)n, mmala ? (size_t)st.K * st.d * st.d : 0};
        return rmn_rowcomm_allreduce_f64(&rowc, bufs, counts, 3, stream);
    }
    explicit LogisticSampler(rmn_sampler* s_) : s(s_) {
        fill_geometry(st, s->model, s->K);
        mmala = (s->prop->kind == RMN_PROP_MMALA);
        pkind = (s->prop->kind == RMN_PROP_RW) ? LG_RW : ((s->prop->kind == RMN_PROP_PCN) ? LG_PCN : LG_HMC);
        tcx3 = s->precision == RMN_PREC_TF32X3;
        tf32m = mmala && (s->precision == RMN_PREC_TF32_METRIC || tcx3);
        // tf32x3: the fused sweep hands p(1-p) over as bf16 and the metric GEMM contracts bf16 operands (k-blocks of 64
        // rows); RMN_MMALA_BF16=0 keeps the single-pass TF32 product of the tf32-metric mode
        bf16m = mmala && tcx3;
        if (const char* e = getenv("RMN_MMALA_BF16")) bf16m = bf16m && !(e[0] == '0');
        st.kr_bf16 = bf16m ? 1 : 0;
        if (tf32m || tcx3) st.Npad = bf16m ? (st.N + 63) / 64 * 64 : (st.N + 31) / 32 * 32;
        if (tf32m) st.NP = (st.d * (st.d + 1) / 2 + 3) / 4 * 4;
        st.gsplit = 1;
        if (tf32m) {
            // a small chain shard (BASELINE's 4,096 chains over 8 GPUs = 512 each: 4 x 9 tiles of 128 x 256) leaves most
            // SMs without a tile while the contraction is N rows deep: split it so that about one wave of tiles exists
            int dev = 0, sms = 148;
            if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) {
                sms = 148;
                cudaGetLastError();
            }
            const int64_t tiles = ((st.K + tc::TM - 1) / tc::TM) * ((st.NP + tc::TN - 1) / tc::TN);
            int want = (int)std::min<int64_t>(8, sms / std::max<int64_t>(tiles, 1));
            if (const char* e = getenv("RMN_MMALA_KSPLIT")) want = std::min(atoi(e), 8);       // fill_metric sums at most 8
            if (want > 1) st.gsplit = tc::splits_used((int)st.Npad, want, bf16m);
        }
        if (tcx3) {
            fused = lgf::supported(st.d);
            if (fused) lgf::make_geometry(&fg, st.N, st.d, st.K);
        }
    }
    size_t f_bytes(int which) const {      // fused sweep: 0 Xh (= Xl), 1 label masks, 2 llp, 3 gp
        switch (which) {
            case 0: return align256((size_t)st.N * fg.ldx * 4);
            case 1: return align256((size_t)fg.nys * 4);
            case 2: return align256((size_t)lgf::LLP_PER_SPLIT * fg.ns * st.K * 8);
            default: return align256((size_t)fg.ns * st.K * fg.dp32 * 4);
        }
    }
    size_t kr_bytes() const { return align256((size_t)st.NP * st.Npad * (bf16m ? 2 : 4)); }
    size_t w_bytes() const { return align256((size_t)st.K * st.Npad * (bf16m ? 2 : 4)); }
    size_t gp_bytes() const { return align256((size_t)st.gsplit * st.K * st.NP * 4); }
    size_t rowb() const { return align256((size_t)st.K * st.dp * 8); }
    size_t eval_smem() const { return eval_smem_bytes(st.ldt); }
    size_t metric_smem() const { return ((size_t)4 * st.ldt + 2 * BI * st.ldt + 4 * BI) * 8; }
    // mMALA: one warp per chain with its own [d][d + 1] matrix in shared memory (34 KB at d = 64): blocks of TWO warps, so
    // that three blocks (six warps) fit an SM instead of one block of four
    int fp_threads() const { return mmala ? 64 : 128; }
    unsigned fp_grid() const { return (unsigned)((st.K * 32 + fp_threads() - 1) / fp_threads()); }
    size_t fp_smem() const { return mmala ? (size_t)(fp_threads() / 32) * (st.d * (st.d + 1) + 2 * st.dp) * 8 : 0; }   // LDL(d) = d + 1
    size_t workspace_bytes() const override {
        const size_t K = (size_t)st.K;
        size_t n = 5 * rowb() + align256((size_t)st.nsplit * K * 8) +
                   align256((size_t)st.nsplit * K * st.dp * 8) + 7 * align256(K * 8) + align256(K * 4) +
                   2 * align256(ND_MAX * K * 8) + 256;
        if (mmala) n += 3 * align256(K * st.d * st.d * 8) + 2 * align256(K * 8) + 2 * rowb();
        if (tf32m) n += kr_bytes() + w_bytes() + gp_bytes();
        if (fused) n += 2 * f_bytes(0) + f_bytes(1) + f_bytes(2) + f_bytes(3);
        return n;
    }
    int bind(void* ws) override {
        const size_t K = (size_t)st.K;
        char* p = (char*)ws;
        st.Th = (double*)p; p += 2 * rowb();
        st.Gr = (double*)p; p += 2 * rowb();
        st.Xi = (double*)p; p += rowb();
        st.llpart = (double*)p; p += align256((size_t)st.nsplit * K * 8);
        st.gpart = (double*)p; p += align256((size_t)st.nsplit * K * st.dp * 8);
        st.lp = (double*)p; p += align256(K * 8);
        st.k0 = (double*)p; p += align256(K * 8);
        st.epsrow = (double*)p; p += align256(K * 8);
        st.scale = (double*)p; p += align256(K * 8);
        st.nsamp = (long long*)p; p += align256(K * 8);
        st.nacc = (long long*)p; p += align256(K * 8);
        st.dacc = (long long*)p; p += align256(K * 8);
        st.cur = (int*)p; p += align256(K * 4);
        st.S1 = (double*)p; p += align256(ND_MAX * K * 8);
        st.S2 = (double*)p; p += align256(ND_MAX * K * 8);
        if (mmala) {
            st.Gm = (double*)p; p += align256(K * st.d * st.d * 8);
            st.Lc = (double*)p; p += 2 * align256(K * st.d * st.d * 8);
            st.logdet = (double*)p; p += 2 * align256(K * 8);
            st.nat = (double*)p; p += 2 * rowb();
        }
        if (tf32m) {
            st.KR = (float*)p; p += kr_bytes();
            st.W = (float*)p; p += w_bytes();
            st.Gp = (float*)p; p += gp_bytes();
        }
        if (fused) {
            fXh = (float*)p; p += f_bytes(0); fXl = (float*)p; p += f_bytes(0);
            fys = (uint32_t*)p; p += f_bytes(1);
            fllp = (double*)p; p += f_bytes(2);
            fgp = (float*)p; p += f_bytes(3);
        }
        RMN_CUDA(cudaMemset(ws, 0, workspace_bytes()));
        if (pkind != LG_HMC) {
            const rmn_proposal* pr = s->prop;
            const size_t nb = (size_t)st.d * st.d * 8;
            RMN_CUDA(cudaMalloc(&d_PL, nb));
            RMN_CUDA(cudaMemcpy(d_PL, pr->h_L.data(), nb, cudaMemcpyHostToDevice));
            st.PL = d_PL;
            if (pkind == LG_PCN) {
                RMN_CUDA(cudaMalloc(&d_PLinv, nb));
                RMN_CUDA(cudaMemcpy(d_PLinv, pr->h_Linv.data(), nb, cudaMemcpyHostToDevice));
                st.PLinv = d_PLinv;
                st.rho = pr->rho; st.rho_c = sqrt(1.0 - pr->rho * pr->rho);     // randomwalk.py:85
            }
        }
        if (pkind == LG_HMC && s->prop->has_mass) {
            const rmn_proposal* pr = s->prop;
            const size_t nb = (size_t)st.d * st.d * 8;
            RMN_CUDA(cudaMalloc(&d_PL, nb));
            RMN_CUDA(cudaMemcpy(d_PL, pr->h_chM.data(), nb, cudaMemcpyHostToDevice));
            RMN_CUDA(cudaMalloc(&d_PLinv, nb));
            RMN_CUDA(cudaMemcpy(d_PLinv, pr->h_chMinv.data(), nb, cudaMemcpyHostToDevice));
            RMN_CUDA(cudaMalloc(&d_Minv, nb));
            RMN_CUDA(cudaMemcpy(d_Minv, pr->h_Minv.data(), nb, cudaMemcpyHostToDevice));
            st.PL = d_PL; st.PLinv = d_PLinv; st.Minv = d_Minv;
        }
        if (int rc = lg_tables_ready()) return rc;
        if (fused) {
            if (int rc = lgf::prep_x(st.N, st.d, fg, st.X, st.y, fXh, fXl, fys, 0)) return rc;
            if (int rc = lgf::make_maps(&fmaps, fg, st.N, fXh, fXl)) return rc;
        }
        if (tf32m) {
            const size_t ksm = (size_t)32 * (st.d + 1) * 8;
            lg_build_kr_kernel<<<(unsigned)(st.Npad / 32), 256, ksm>>>(st);
            RMN_KERNEL_CHECK();
            if (bf16m) {
                if (int rc = tc::make_tmap_2d_bf16(&maps.ah, st.W, (uint64_t)st.K, (uint64_t)st.Npad, (uint64_t)st.Npad, tc::TM)) return rc;
                if (int rc = tc::make_tmap_2d_bf16(&maps.bh, st.KR, (uint64_t)st.NP, (uint64_t)st.Npad, (uint64_t)st.Npad, tc::TN)) return rc;
            } else {
                if (int rc = tc::make_tmap_2d(&maps.ah, st.W, (uint64_t)st.K, (uint64_t)st.Npad, (uint64_t)st.Npad, tc::TM)) return rc;
                if (int rc = tc::make_tmap_2d(&maps.bh, st.KR, (uint64_t)st.NP, (uint64_t)st.Npad, (uint64_t)st.Npad, tc::TN)) return rc;
            }
            maps.al = maps.ah; maps.bl = maps.bh;
        }
        RMN_RAISE_SMEM(lg_eval_kernel, (int)eval_smem());
        if (mmala) {
            RMN_RAISE_SMEM(lg_metric_kernel, (int)metric_smem());
            RMN_RAISE_SMEM(lg_finish_propose_kernel<true>, (int)fp_smem());
            RMN_RAISE_SMEM(lg_adopt_kernel<true>, (int)fp_smem());
        }
        int rc = rmn_fill_f64(st.scale, st.K, 1.0, 0);
        if (rc) return rc;
        RMN_CUDA(cudaDeviceSynchronize());
        return RMN_OK;
    }
    // Lc / logdet slots must be exactly K*d*d / K apart
    unsigned row_grid() const { return (unsigned)((st.K * 32 + 127) / 128); }

    // RMN_PREC_TF32X3: logits GEMM -> pointwise -> split-K gradient GEMM -> partial sums, per chain block
    int eval_fused(int fixed_slot, cudaStream_t stream) {
        lgf::SweepArgs a{};
        a.ys = fys; a.Th = st.Th; a.cur = st.cur; a.fixed_slot = fixed_slot;
        a.K = st.K; a.N = st.N; a.d = st.d; a.dp = st.dp;
        a.llp = fllp; a.gp = fgp; a.W = st.W; a.ldw = st.Npad; a.w_bf16 = bf16m ? 1 : 0;
        static long long* d_tl = nullptr;                          // RMN_LGF_TIMELINE=1: clock64 stamps of CTA 0 (debug aid)
        if (const char* e = getenv("RMN_LGF_TIMELINE")) {
            if (e[0] == '1' && !d_tl) { cudaMalloc(&d_tl, 256 * 8 * 8); }
            if (e[0] == '1' && d_tl) { cudaMemsetAsync(d_tl, 0, 256 * 8 * 8, stream); a.dbg = d_tl; }
        }
        if (!tf32m) ktimer.begin("lg_fused_sweep_kernel", stream);     // mMALA: the metric GEMM is the dominant kernel
        if (int rc = lgf::sweep(fmaps, fg, a, stream)) return rc;
        if (!tf32m) ktimer.end(stream);
        if (int rc = lgf::reduce(fg, st.K, st.dp, fllp, fgp, st.llpart, st.gpart, stream)) return rc;
        if (a.dbg) {
            if (const char* f = getenv("RMN_LGF_TIMELINE_FILE")) {
                std::vector<long long> h(256 * 8);
                cudaStreamSynchronize(stream);
                cudaMemcpy(h.data(), a.dbg, h.size() * 8, cudaMemcpyDeviceToHost);
                if (FILE* fp = fopen(f, "wb")) { fwrite(h.data(), 8, h.size(), fp); fclose(fp); }
            }
        }
        launches += 2;
        return RMN_OK;
    }
    // Gp = W KR^T on tcgen05 (bf16 operands in tf32x3 mode, single-pass TF32 in tf32-metric mode), split-K partials
    // when the shard's tile count is small (metric_entry sums them)
    int metric_gemm(cudaStream_t stream) {
        if (st.gsplit > 1)
            return tc::launch_single_splitk(maps, st.K, st.NP, (int)st.Npad, st.Gp, st.NP, st.gsplit,
                                            (int64_t)st.K * st.NP, bf16m, stream);
        return bf16m ? tc::launch_plain_bf16(maps, st.K, st.NP, (int)st.Npad, st.Gp, st.NP, stream)
                     : tc::launch_plain_tf32(maps, st.K, st.NP, (int)st.Npad, st.Gp, st.NP, stream);
    }
    int eval(int fixed_slot, cudaStream_t stream) {
        if (tcx3) {
            if (int rc = eval_fused(fixed_slot, stream)) return rc;
            if (tf32m) {
                ktimer.begin(bf16m ? "tf32x3_gemm_kernel<1,bf16>" : "tf32x3_gemm_kernel<1>", stream);
                if (int rc = metric_gemm(stream)) return rc;
                ktimer.end(stream);
                launches++;
            }
            return RMN_OK;
        }
        dim3 grid((unsigned)((st.K + BC - 1) / BC), st.nsplit);
        const bool time_eval = !mmala || tf32m;          // the dominant kernel of this configuration
        if (time_eval) ktimer.begin("lg_eval_kernel", stream);
        lg_eval_kernel<<<grid, EVAL_THREADS, eval_smem(), stream>>>(st, fixed_slot);
        if (time_eval) ktimer.end(stream);
        RMN_KERNEL_CHECK(); launches++;
        if (tf32m) {
            if (int rc = metric_gemm(stream)) return rc;
            launches++;
        } else if (mmala) {
            ktimer.begin("lg_metric_kernel", stream);
            lg_metric_kernel<<<(unsigned)((st.K + 3) / 4), 256, metric_smem(), stream>>>(st, fixed_slot);
            ktimer.end(stream);
            RMN_KERNEL_CHECK(); launches++;
        }
        return exchange(stream);
    }
    int set_state(const double* d_theta, cudaStream_t stream) override {
        const int64_t n = st.K * st.dp;
        lg_set_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(st, d_theta);
        RMN_KERNEL_CHECK(); launches++;
        if (int rc = eval(1, stream)) return rc;
        if (mmala) lg_adopt_kernel<true><<<fp_grid(), fp_threads(), fp_smem(), stream>>>(fst());
        else lg_adopt_kernel<false><<<row_grid(), 128, 0, stream>>>(fst());
        RMN_KERNEL_CHECK(); launches++;
        return RMN_OK;
    }
    int get_state(double* d_theta, double* d_lp, cudaStream_t stream) override {
        const int64_t n = st.K * st.d;
        lg_get_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(st, d_theta, d_lp);
        RMN_KERNEL_CHECK(); launches++;
        return RMN_OK;
    }
    int run(int64_t T, const rmn_inject_t* inj, const rmn_trace_t* tr, cudaStream_t stream) override {
        const rmn_proposal* pr = s->prop;
        if (inj) RMN_REQUIRE(inj->d_xi && inj->d_u, "injected run needs d_xi and d_u");
        if (inj && st.nt > 0) RMN_REQUIRE(inj->d_usel, "injected tempered run needs d_usel (selection uniforms)");
        rmn_trace_t t0{};
        if (tr) t0 = *tr;
        if (t0.thin <= 0) t0.thin = 1;
        LgStep sp{};
        sp.mmala = mmala; sp.adapt = pr->adapt; sp.target = pr->target; sp.eps0 = (pkind == LG_HMC) ? pr->eps : 1.0;
        sp.pkind = pkind;
        sp.seed = s->seed; sp.chain_offset = s->chain_offset;
        const int64_t K = st.K;
        const int d = st.d;
        for (int64_t t = 0; t <= T; ++t) {
            sp.finish = (t > 0); sp.propose = (t < T); sp.diag = (t > 0); sp.record = sp.finish;
            sp.step_fin = step0 + t - 1; sp.step_prop = step0 + t;
            sp.inj_u = (inj && t > 0) ? inj->d_u + (t - 1) * K : nullptr;
            sp.inj_xi = (inj && t < T) ? inj->d_xi + t * K * d : nullptr;
            sp.inj_usel = (inj && st.nt > 0 && t < T) ? inj->d_usel + t * K : nullptr;
            sp.trace_slot = -1;
            sp.tr_theta = t0.d_theta; sp.tr_logpost = t0.d_logpost;
            sp.tr_prop_lp = (t0.d_prop_logpost && t > 0) ? t0.d_prop_logpost + (t - 1) * K : nullptr;
            sp.tr_acc = (t0.d_accepted && t > 0) ? t0.d_accepted + (t - 1) * K : nullptr;
            sp.tr_lqr = (t0.d_logqratio && t > 0) ? t0.d_logqratio + (t - 1) * K : nullptr;
            sp.tr_prop_theta = (t0.d_prop_theta && t > 0) ? t0.d_prop_theta + (t - 1) * K * d : nullptr;
            if (t > 0 && (t0.d_theta || t0.d_logpost)) {
                const int64_t i = t;
                if (i >= t0.first && (i - t0.first) % t0.thin == 0) sp.trace_slot = (i - t0.first) / t0.thin;
            }
            if (st.nt > 0) {
                // tempering: a swap moves TWO rows -- finish in one launch (the initiator's warp exchanges both rows),
                // record and propose in a second one
                if (sp.finish) {
                    LgStep fin = sp; fin.propose = 0; fin.record = 0; fin.diag = 0;
                    lg_finish_propose_kernel<false><<<row_grid(), 128, 0, stream>>>(fst(), fin);
                    RMN_KERNEL_CHECK(); launches++;
                }
                LgStep pro = sp; pro.finish = 0;
                pro.tr_prop_lp = nullptr; pro.tr_acc = nullptr; pro.tr_lqr = nullptr; pro.tr_prop_theta = nullptr;
                lg_finish_propose_kernel<false><<<row_grid(), 128, 0, stream>>>(fst(), pro);
            } else if (mmala) lg_finish_propose_kernel<true><<<fp_grid(), fp_threads(), fp_smem(), stream>>>(fst(), sp);
            else lg_finish_propose_kernel<false><<<row_grid(), 128, 0, stream>>>(fst(), sp);
            RMN_KERNEL_CHECK(); launches++;
            if (t == T) break;
            if (int rc = eval(-1, stream)) return rc;
            if (!mmala && pkind == LG_HMC) {
                // Nsteps - 1 interior leapfrog steps: each one a full likelihood sweep at the new trajectory point
                for (int l = 1; l < pr->nsteps; ++l) {
                    const int64_t n = st.K * st.dp;
                    if (st.Minv) lg_leapfrog_mid_mass_kernel<<<row_grid(), 128, 0, stream>>>(fst());
                    else lg_leapfrog_mid_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(fst());
                    RMN_KERNEL_CHECK(); launches++;
                    if (int rc = eval(-1, stream)) return rc;
                }
            }
        }
        step0 += T; diag_steps += T;
        return RMN_OK;
    }
    int get_adapt(double* sc, int64_t* ns, int64_t* na, cudaStream_t stream) override {
        lg_get_adapt_kernel<<<(unsigned)((st.K + 127) / 128), 128, 0, stream>>>(st, sc, ns, na);
        RMN_KERNEL_CHECK(); launches++;
        return RMN_OK;
    }
    int set_adapt(const double* sc, const int64_t* ns, const int64_t* na, cudaStream_t stream) override {
        launches++;
        return rmn_copy_adapt(st.K, sc, ns, na, st.scale, st.nsamp, st.nacc, stream);
    }
    int diag_dim() const override { return (st.d < ND_MAX - 1 ? st.d : ND_MAX - 1) + 1; }
    int reset_diag(cudaStream_t stream) override {
        RMN_CUDA(cudaMemsetAsync(st.S1, 0, (size_t)ND_MAX * st.K * 8, stream));
        RMN_CUDA(cudaMemsetAsync(st.S2, 0, (size_t)ND_MAX * st.K * 8, stream));
        RMN_CUDA(cudaMemsetAsync(st.dacc, 0, (size_t)st.K * 8, stream));
        diag_steps = 0;
        return RMN_OK;
    }
    int chain_sums(const double** S1, const double** S2, int64_t* n) override {
        *S1 = st.S1; *S2 = st.S2; *n = diag_steps;
        return RMN_OK;
    }
    int reduce_diag(double* d_block, cudaStream_t stream) override {
        launches++;
        return rmn_reduce_diag_block(st.K, diag_dim(), diag_steps, diag_steps, st.S1, st.S2, st.dacc, nullptr, d_block, stream);
    }
};

}  // namespace

SamplerImpl* make_logistic_sampler(rmn_sampler* s) {
    const rmn_proposal* p = s->prop;
    const int d = s->model->d;
    if (d > 128) {
        rmn_set_error("logistic kernels support d <= 128 (got %d)", d);
        return nullptr;
    }
    if (p->kind == RMN_PROP_MMALA) {
        if (d > MM_DMAX) {
            rmn_set_error("mMALA supports d <= %d (shared-memory Cholesky); got %d", MM_DMAX, d);
            return nullptr;
        }
        return new LogisticSampler(s);
    }
    if (p->kind == RMN_PROP_HMC && !p->acov) return new LogisticSampler(s);           // VanillaHMC / AdaptScaleHMC, with or without M
    if (p->kind == RMN_PROP_RW && !p->acov) return new LogisticSampler(s);            // MetropolisRandomWalk / AdaptScaleRandomWalk
    if (p->kind == RMN_PROP_PCN && !p->adapt) return new LogisticSampler(s);          // pCN (AdaptScalepCN: small-d family only)
    rmn_set_error("logistic model: device kernels exist for MetropolisRandomWalk / AdaptScaleRandomWalk, pCN, VanillaHMC / "
                  "AdaptScaleHMC (Nsteps = 1 is MALA; fixed mass matrix or none) and SimplifiedMMALA");
    return nullptr;
}

// Pointwise evaluation of n arbitrary points: stage them as a temporary chain block and run
// the same kernels the sampler uses (parity harness entry; allocates scratch, synchronous).
int logistic_pointwise(rmn_model* m, int which, int64_t n, const double* d_theta, double* d_out,
                       double* d_grad, double* d_metric, cudaStream_t stream) {
    if (n <= 0) return RMN_OK;
    if (m->d > 128) { rmn_set_error("logistic kernels support d <= 128"); return RMN_ERR_UNSUPPORTED; }
    if (d_metric && m->d > MM_DMAX) { rmn_set_error("metric supports d <= %d", MM_DMAX); return RMN_ERR_UNSUPPORTED; }
    LogisticState st{};
    fill_geometry(st, m, n);
    const size_t rowb = align256((size_t)n * st.dp * 8);
    const size_t sz_ll = align256((size_t)st.nsplit * n * 8);
    const size_t sz_g = align256((size_t)st.nsplit * n * st.dp * 8);
    const size_t sz_m = d_metric ? align256((size_t)n * st.d * st.d * 8) : 0;
    const size_t sz_s = align256((size_t)n * 8);
    const size_t total = 3 * rowb + sz_ll + sz_g + sz_m + 3 * sz_s;
    char* buf = nullptr;
    RMN_CUDA(cudaMalloc(&buf, total));
    RMN_CUDA(cudaMemsetAsync(buf, 0, total, stream));
    char* p = buf;
    // the two Th slots must be exactly n*dp doubles apart: carve them from one 2*rowb block
    st.Th = (double*)p; p += 2 * rowb;
    st.Xi = (double*)p; p += rowb;
    st.llpart = (double*)p; p += sz_ll;
    st.gpart = (double*)p; p += sz_g;
    if (d_metric) { st.Gm = (double*)p; p += sz_m; }
    st.cur = (int*)p; p += sz_s;
    st.k0 = (double*)p; p += sz_s;
    st.epsrow = (double*)p; p += sz_s;
    double* scratch = nullptr;
    if (int rc = lg_tables_ready()) return rc;
    const size_t esm = eval_smem_bytes(st.ldt);
    const size_t msm = ((size_t)4 * st.ldt + 2 * BI * st.ldt + 4 * BI) * 8;
    RMN_RAISE_SMEM(lg_eval_kernel, (int)esm);
    const int64_t ne = n * st.dp;
    lg_set_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, stream>>>(st, d_theta);
    dim3 grid((unsigned)((n + BC - 1) / BC), st.nsplit);
    lg_eval_kernel<<<grid, EVAL_THREADS, esm, stream>>>(st, 1);
    if (d_metric) {
        RMN_RAISE_SMEM(lg_metric_kernel, (int)msm);
        lg_metric_kernel<<<(unsigned)((n + 3) / 4), 256, msm, stream>>>(st, 1);
    }
    lg_point_out_kernel<<<(unsigned)((n * 32 + 127) / 128), 128, 0, stream>>>(st, which, d_out, d_grad, d_metric);
    cudaError_t e = cudaGetLastError();
    cudaError_t e2 = cudaStreamSynchronize(stream);
    cudaFree(buf);
    cudaFree(scratch);
    if (e != cudaSuccess || e2 != cudaSuccess) {
        rmn_set_error("logistic_pointwise: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
        return RMN_ERR_CUDA;
    }
    return RMN_OK;
}

// ---------------------------------------------------------------------------------------
// validation entry for logistic_math.cuh: p = sigmoid(z), sp = softplus(z), pq = p(1-p)
// ---------------------------------------------------------------------------------------
namespace {
__global__ void lg_math_kernel(int64_t n, const double* __restrict__ z, double* p, double* sp, double* pq) {
    __shared__ double tab[lgmath::TAB_DOUBLES];
    for (int q = threadIdx.x; q < lgmath::TAB_DOUBLES; q += blockDim.x) tab[q] = g_lg_tab[q];
    __syncthreads();
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a, b, c;
    lgmath::sigmoid_softplus(z[i], tab, a, b, c);
    p[i] = a; sp[i] = b; pq[i] = c;
}
}  // namespace

extern "C" int rmn_logistic_math(int64_t n, const double* d_z, double* d_p, double* d_sp, double* d_pq, void* stream) {
    RMN_REQUIRE(n >= 0 && d_z && d_p && d_sp && d_pq, "rmn_logistic_math: bad argument");
    if (n == 0) return RMN_OK;
    if (int rc = lg_tables_ready()) return rc;
    lg_math_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, d_z, d_p, d_sp, d_pq);
    RMN_KERNEL_CHECK();
    return RMN_OK;
}
