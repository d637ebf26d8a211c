// riemann_b200 -- fully fused T-step MH kernel for small-d Gaussian targets.
//
// One thread owns one chain; theta, log-posterior and the AdaptScale state stay in
// registers for all T iterations of a launch.  Replaces, per iteration,
//   Sampler.sample                       riemann/samplers/sampler.py:72-90
//   MetropolisRandomWalk.propose         riemann/proposals/randomwalk.py:21-26
//   VanillaHMC.propose + leapfrog        riemann/proposals/hamiltonian.py:13-52,76-91
//   pCN.propose                          riemann/proposals/randomwalk.py:88-100
//   AdaptScaleProposal.adapt             riemann/proposals/adaptive.py:26-35
//   MultiGaussianDist.log_likelihood / grad_log_likelihood   riemann/models/gaussian.py:49-58
//   Model.log_posterior                  riemann/models/model.py:43-55
//   AdaptCovProposal.adapt (Haario)      riemann/proposals/adaptive.py:38-103 (per chain; ACOV instantiation)
//   PTSampler.sample / TemperedModel     riemann/samplers/ptsampler.py:11-38, 92-127  (set_tempering:
//                                        ladders of nt chains on consecutive threads, swaps by shuffle)
// Arithmetic is fp64 throughout (the reference is fp64 numpy).
// Internal state layout: theta[D][K] (chain fastest => coalesced loads/stores).
#include "common.cuh"

namespace {

template <int D>
struct SGParams {
    static constexpr int TRI = D * (D + 1) / 2;
    double mu[D];
    double linv[TRI];       // L^{-1}, C = L L^T, packed lower row-major
    double c1, c2;          // d*log(2 pi), logdetC   (gaussian.py:52)
    int kind, adapt, nsteps, has_mass;
    double target;
    double lprop[TRI];      // chol(C0) for RW / pCN
    double lpinv[TRI];      // its inverse (pCN)
    double rho, rho_c;
    double eps0;
    double chM[TRI];
    double Minv[D * D];
    double chMinv[TRI];
    // parallel tempering (ptsampler.py): nt = 0 off; a ladder occupies `stride` (power of two >= nt) threads
    int nt, stride;
    double pswap;
    double betas[32];
    // covariance adaptation (adaptive.py:38-103): acov = 0 off
    int acov, marginalize, smooth;
    double t_adapt, creg, dpow04, dpow02;   // creg = mean(diag(C0)); d**0.4, d**0.2 (:101-102)
    double c0[D * D];
    const uint32_t* acmask;                 // bit n: does the reference adapt at state count n (see haario_schedule)
    long long acmask_n;
};

struct SGState {
    double* theta;          // [D][K]
    double* lp;             // [K]   (tempered log-posterior when a ladder is set)
    double* ll;             // [K]   untempered log-likelihood (parallel tempering)
    double* scale;          // [K]
    long long* nsamp;       // [K]
    long long* nacc;        // [K]
    long long* dacc;        // [K] accepts since diagnostics reset
    double* S1;             // [D][K]
    double* S2;             // [D][K]
    // Haario state per chain (ACOV): n, sum x, sum x x^T (full), current proposal Cholesky factor (packed lower)
    double* acS;            // [K]
    double* acSX;           // [D][K]
    double* acSX2;          // [D*D][K]
    double* acL;            // [TRI][K]
    double* rho;            // [K]   AdaptScalepCN: the chain's current rho (randomwalk.py:118)
};

template <int D>
__device__ __forceinline__ void tri_mv(const double* __restrict__ L, const double* x, double* y) {
    // y = L x, L packed lower
#pragma unroll
    for (int i = 0; i < D; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j <= i; ++j) s += L[i * (i + 1) / 2 + j] * x[j];
        y[i] = s;
    }
}

template <int D>
__device__ __forceinline__ void tri_tmv(const double* __restrict__ L, const double* x, double* y) {
    // y = L^T x
#pragma unroll
    for (int j = 0; j < D; ++j) {
        double s = 0.0;
#pragma unroll
        for (int i = j; i < D; ++i) s += L[i * (i + 1) / 2 + j] * x[i];
        y[j] = s;
    }
}

template <int D>
__device__ __forceinline__ double gauss_loglik(const SGParams<D>& P, const double* th) {
    double y[D], u[D];
#pragma unroll
    for (int i = 0; i < D; ++i) y[i] = th[i] - P.mu[i];
    tri_mv<D>(P.linv, y, u);
    double q = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) q += u[i] * u[i];
    return -0.5 * ((q + P.c1) + P.c2);        // gaussian.py:52
}
template <int D>
__device__ __forceinline__ double gauss_logpost(const SGParams<D>& P, const double* th) {
    return combine_logpost(0.0, gauss_loglik<D>(P, th));          // log_prior == 0.0 (gaussian.py:46-47)
}
// thread -> chain mapping: plain = one chain per thread; tempered = ladder l on threads l*stride .. l*stride+nt-1
template <int D>
__device__ __forceinline__ int64_t chain_of_thread(const SGParams<D>& P, int64_t g, int64_t K, int& ti, bool& active) {
    if (P.nt == 0) { ti = 0; active = g < K; return active ? g : 0; }
    const int64_t ladder = g / P.stride;
    ti = (int)(g % P.stride);
    active = ti < P.nt && (ladder + 1) * P.nt <= K;
    if (ti >= P.nt) ti = P.nt - 1;
    return active ? ladder * P.nt + ti : 0;
}

template <int D>
__device__ __forceinline__ void gauss_grad(const SGParams<D>& P, const double* th, double* g) {
    double y[D], u[D];
#pragma unroll
    for (int i = 0; i < D; ++i) y[i] = th[i] - P.mu[i];
    tri_mv<D>(P.linv, y, u);
    tri_tmv<D>(P.linv, u, g);
#pragma unroll
    for (int i = 0; i < D; ++i) g[i] = -g[i];
}

// y = L^-1 x (forward substitution), L packed lower
template <int D>
__device__ __forceinline__ void tri_solve(const double* L, const double* x, double* y) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
        double s = x[i];
#pragma unroll
        for (int j = 0; j < i; ++j) s -= L[i * (i + 1) / 2 + j] * y[j];
        y[i] = s / L[i * (i + 1) / 2 + i];
    }
}
// y = L^-T x (backward substitution), L packed lower
template <int D>
__device__ __forceinline__ void tri_solve_t(const double* L, const double* x, double* y) {
#pragma unroll
    for (int i = D - 1; i >= 0; --i) {
        double s = x[i];
#pragma unroll
        for (int j = i + 1; j < D; ++j) s -= L[j * (j + 1) / 2 + i] * y[j];
        y[i] = s / L[i * (i + 1) / 2 + i];
    }
}

template <int D>
__device__ __forceinline__ void velocity(const SGParams<D>& P, const double* p, double* v) {
    if (P.has_mass) {
#pragma unroll
        for (int i = 0; i < D; ++i) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < D; ++j) s += P.Minv[i * D + j] * p[j];
            v[i] = s;
        }
    } else {
#pragma unroll
        for (int i = 0; i < D; ++i) v[i] = p[i];
    }
}

// AdaptCovProposal.adapt for one chain: accumulate the state, and when n is a perfect square > 2 rebuild
// C from the sample covariance and refactor it (adaptive.py:70-102).  Arithmetic in the reference's order
// with explicit roundings (no FMA contraction), so the proposal matrix tracks numpy's to the last bits.
template <int D>
__device__ __forceinline__ void haario_adapt(const SGParams<D>& P, const double* x, double& n, double* sx, double* sx2,
                                             double* L) {
    n += 1.0;
#pragma unroll
    for (int i = 0; i < D; ++i) sx[i] = __dadd_rn(sx[i], x[i]);
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) sx2[i * D + j] = __dadd_rn(sx2[i * D + j], __dmul_rn(x[i], x[j]));
    // :82-83  `np.sqrt(n)**2 == n and n > 2`: true for the perfect squares the comment there intends and, through
    // rounding, for about half of all other n; the host tabulates it with the reference's own expression
    bool hit;
    if (n < (double)P.acmask_n) {
        const unsigned ni = (unsigned)n;
        hit = (P.acmask[ni >> 5] >> (ni & 31u)) & 1u;
    } else {
        const double r = sqrt(n);
        hit = __dmul_rn(r, r) == n;
    }
    if (!(hit && n > 2.0)) return;
    double Cm[D * D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const double cs = __ddiv_rn(__dsub_rn(sx2[i * D + j], __ddiv_rn(__dmul_rn(sx[i], sx[j]), n)), n - 1.0);   // :84
            double c;
            if (P.smooth) c = __ddiv_rn(__dadd_rn(__dmul_rn(n, cs), __dmul_rn(P.t_adapt, P.c0[i * D + j])), __dadd_rn(n, P.t_adapt));   // :87
            else c = __dadd_rn(cs, (i == j) ? __dmul_rn(1e-12, P.creg) : 0.0);   // :92-93 (n >= t_adapt guaranteed by the host)
            if (P.marginalize && i != j) c = 0.0;                                // :96-97
            Cm[i * D + j] = __ddiv_rn(c, P.dpow04);                              // :101
        }
    // L = chol(C) / d**0.2   (:102); lower, packed
#pragma unroll
    for (int j = 0; j < D; ++j) {
        double s = Cm[j * D + j];
#pragma unroll
        for (int k = 0; k < j; ++k) s -= L[j * (j + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
        const double ljj = sqrt(s);
        L[j * (j + 1) / 2 + j] = ljj;
#pragma unroll
        for (int i = j + 1; i < D; ++i) {
            double v = Cm[i * D + j];
#pragma unroll
            for (int k = 0; k < j; ++k) v -= L[i * (i + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
            L[i * (i + 1) / 2 + j] = v / ljj;
        }
    }
#pragma unroll
    for (int q = 0; q < SGParams<D>::TRI; ++q) L[q] = L[q] / P.dpow02;
}

// RWONLY: instantiation for the plain random walk (config 1's throughput case): pCN / HMC / adaptation code is
// compiled out, which keeps it at the register count the one-proposal kernel had.
template <int D, bool INJ, bool ACOV, bool PT, bool RWONLY = false>
__global__ void __launch_bounds__(128, (D <= 2 && !ACOV) ? 4 : 1)   // d <= 2: 128 registers = 16 warps/SM (config 1)
small_gauss_kernel(const SGParams<D>* __restrict__ gparams, SGState st, int64_t K, int64_t T,
                   int64_t step0, uint64_t seed, int64_t chain_offset, const double* __restrict__ inj_xi,
                   const double* __restrict__ inj_u, const double* __restrict__ inj_usel, rmn_trace_t tr) {
    __shared__ SGParams<D> P;
    {
        const int nw = sizeof(SGParams<D>) / 4;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(gparams);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&P);
        for (int i = threadIdx.x; i < nw; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    int ti;
    bool active;
    const int64_t c = chain_of_thread<D>(P, blockIdx.x * (int64_t)blockDim.x + threadIdx.x, K, ti, active);
    constexpr bool pt = PT;                       // compile-time: the plain kernel carries no ladder logic
    const int kind = RWONLY ? (int)RMN_PROP_RW : P.kind;
    if (!pt && !active) return;                 // tempered warps keep their idle threads for the shuffles
    const double beta = pt ? P.betas[ti] : 1.0;

    double th[D];
#pragma unroll
    for (int i = 0; i < D; ++i) th[i] = st.theta[(int64_t)i * K + c];
    double lp = st.lp[c];
    double ll = pt ? st.ll[c] : 0.0;
    AdaptState ad{st.scale[c], st.nsamp[c], st.nacc[c]};
    double acn = 0.0, acsx[ACOV ? D : 1], acsx2[ACOV ? D * D : 1], acl[ACOV ? SGParams<D>::TRI : 1];
    if (ACOV) {
        acn = st.acS[c];
#pragma unroll
        for (int i = 0; i < D; ++i) acsx[ACOV ? i : 0] = st.acSX[(int64_t)i * K + c];
#pragma unroll
        for (int i = 0; i < D * D; ++i) acsx2[ACOV ? i : 0] = st.acSX2[(int64_t)i * K + c];
#pragma unroll
        for (int i = 0; i < SGParams<D>::TRI; ++i) acl[ACOV ? i : 0] = st.acL[(int64_t)i * K + c];
    }
    long long dacc = st.dacc[c];
    // AdaptScalepCN (randomwalk.py:103-119): rho is re-derived from the previous rho at every proposal,
    // rho_c keeps its initial value -- as written in the reference
    const bool pcn_adapt = !RWONLY && (P.kind == RMN_PROP_PCN) && P.adapt;
    double rho = pcn_adapt ? st.rho[c] : P.rho;
    double s1[D], s2[D];
#pragma unroll
    for (int i = 0; i < D; ++i) { s1[i] = 0.0; s2[i] = 0.0; }
    const RngKey rk(seed, (uint64_t)(chain_offset + c));
    const TraceSel ts{tr.first, tr.thin > 0 ? tr.thin : 1};

    for (int64_t t = 0; t < T; ++t) {
        const uint64_t step = (uint64_t)(step0 + t);
        double xi[D], uacc, usel = 1.0;
        if (INJ) {
#pragma unroll
            for (int i = 0; i < D; ++i) xi[i] = inj_xi[(t * K + c) * D + i];
            uacc = inj_u[t * K + c];
            if (pt) usel = inj_usel[t * K + c];
        } else {
            if (pt) usel = u01(rk.block(step, RMN_BLOCK_AUX).x);
#pragma unroll
            for (int b = 0; 4 * b < D; ++b) {
                double v[4];
                normal4(rk.block(step, (uint32_t)b), v);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (4 * b + q < D) xi[4 * b + q] = v[q];
            }
            uacc = u01(rk.block(step, RMN_BLOCK_ACCEPT).x);
        }

        // ---- parallel tempering roles (ptsampler.py:102-112): chain i initiates a swap with i+1 iff it was not
        //      itself swapped by i-1, u <= Pswap and it is not the last rung; sequential along the ladder
        bool is_init = false, is_part = false;
        if (pt) {
            const bool want = active && ti < P.nt - 1 && !(usel > P.pswap);
            const unsigned base = (threadIdx.x & 31u) & ~(unsigned)(P.stride - 1);
            const unsigned wmask = (__ballot_sync(0xffffffffu, want) >> base) & (P.stride == 32 ? 0xffffffffu : ((1u << P.stride) - 1u));
            unsigned init = 0;
            for (int b = 0; b < P.nt - 1; ++b)
                if (((wmask >> b) & 1u) && !(b > 0 && ((init >> (b - 1)) & 1u))) init |= 1u << b;
            is_init = active && ((init >> ti) & 1u);
            is_part = active && ti > 0 && ((init >> (ti - 1)) & 1u);
        }
        const bool regular = active && !is_init && !is_part;

        double q[D], lqr = 0.0;
        if (kind == RMN_PROP_RW) {
            double lx[D];
            if (ACOV) tri_mv<D>(acl, xi, lx);            // this chain's adapted factor
            else tri_mv<D>(P.lprop, xi, lx);
#pragma unroll
            for (int i = 0; i < D; ++i) q[i] = th[i] + ad.scale * lx[i];
        } else if (kind == RMN_PROP_PCN) {
            double lx[D], df[D], dr[D], uf[D], ur[D];
            if (pcn_adapt) rho = tanh(rho / ad.scale);                 // randomwalk.py:118
            tri_mv<D>(P.lprop, xi, lx);
#pragma unroll
            for (int i = 0; i < D; ++i) q[i] = rho * th[i] + P.rho_c * lx[i];
#pragma unroll
            for (int i = 0; i < D; ++i) { df[i] = q[i] - rho * th[i]; dr[i] = th[i] - rho * q[i]; }
            tri_mv<D>(P.lpinv, df, uf);
            tri_mv<D>(P.lpinv, dr, ur);
            double a = 0.0, b = 0.0;
#pragma unroll
            for (int i = 0; i < D; ++i) { a += uf[i] * uf[i]; b += ur[i] * ur[i]; }
            lqr = -0.5 * (a - b) / (P.rho_c * P.rho_c);
        } else {   // HMC / MALA
            const double eps = P.adapt ? ad.scale * P.eps0 : P.eps0;
            double p0[D], p[D], g[D], v[D];
            // AdaptCovHMC (hamiltonian.py:106-119): M = C, chM = L of THIS chain.  L = chol(C) / d**0.2 after the first
            // adaptation (n >= 4) and chol(C0) before it, so C^-1 p = L^-T L^-1 p / cfac with cfac = d**0.4 resp. 1.
            const double cfac = (ACOV && acn >= 4.0) ? P.dpow04 : 1.0;
            auto vel = [&](const double* pp, double* vv) {
                if (ACOV) {
                    double y1[D];
                    tri_solve<D>(acl, pp, y1);
                    tri_solve_t<D>(acl, y1, vv);
#pragma unroll
                    for (int i = 0; i < D; ++i) vv[i] = vv[i] / cfac;
                } else {
                    velocity<D>(P, pp, vv);
                }
            };
            if (ACOV) tri_mv<D>(acl, xi, p0);
            else if (P.has_mass) tri_mv<D>(P.chM, xi, p0);
            else {
#pragma unroll
                for (int i = 0; i < D; ++i) p0[i] = xi[i];
            }
            gauss_grad<D>(P, th, g);
#pragma unroll
            for (int i = 0; i < D; ++i) p[i] = p0[i] + 0.5 * eps * g[i];
            vel(p, v);
#pragma unroll
            for (int i = 0; i < D; ++i) q[i] = th[i] + eps * v[i];
            for (int s = 1; s < P.nsteps; ++s) {
                gauss_grad<D>(P, q, g);
#pragma unroll
                for (int i = 0; i < D; ++i) p[i] = p[i] + eps * g[i];
                vel(p, v);
#pragma unroll
                for (int i = 0; i < D; ++i) q[i] = q[i] + eps * v[i];
            }
            gauss_grad<D>(P, q, g);
#pragma unroll
            for (int i = 0; i < D; ++i) p[i] = p[i] + 0.5 * eps * g[i];
            double k0 = 0.0, k1 = 0.0;
            if (ACOV) {
                double w0[D], w1[D];
                tri_solve<D>(acl, p0, w0);                      // solve(chM, p), hamiltonian.py:86-87
                tri_solve<D>(acl, p, w1);
#pragma unroll
                for (int i = 0; i < D; ++i) { k0 += w0[i] * w0[i]; k1 += w1[i] * w1[i]; }
            } else if (P.has_mass) {
                double w0[D], w1[D];
                tri_mv<D>(P.chMinv, p0, w0);
                tri_mv<D>(P.chMinv, p, w1);
#pragma unroll
                for (int i = 0; i < D; ++i) { k0 += w0[i] * w0[i]; k1 += w1[i] * w1[i]; }
            } else {
#pragma unroll
                for (int i = 0; i < D; ++i) { k0 += p0[i] * p0[i]; k1 += p[i] * p[i]; }
            }
            lqr = 0.5 * (k1 - k0);
        }

        const double llq = gauss_loglik<D>(P, q);
        double lpq = combine_logpost(0.0, pt ? llq * beta : llq);       // TemperedModel: logL * beta (ptsampler.py:33-34)
        bool acc = regular && mh_accept(lpq, lp, lqr, uacc);
        bool moved = false;
        if (acc) {
#pragma unroll
            for (int i = 0; i < D; ++i) { moved |= (q[i] != th[i]); th[i] = q[i]; }
            lp = lpq; ll = llq;
        }
        if (pt) {
            // ---- swap proposals (ptsampler.py:113-125): the initiator i sees its lower neighbour j = i+1
            const int W = P.stride;
            double thn[D], thp[D];
#pragma unroll
            for (int i = 0; i < D; ++i) {
                thn[i] = __shfl_down_sync(0xffffffffu, th[i], 1, W);
                thp[i] = __shfl_up_sync(0xffffffffu, th[i], 1, W);
            }
            const double lln = __shfl_down_sync(0xffffffffu, ll, 1, W), lpn = __shfl_down_sync(0xffffffffu, lp, 1, W);
            const double llpv = __shfl_up_sync(0xffffffffu, ll, 1, W);
            const double beta_n = P.betas[min(ti + 1, P.nt - 1)];
            const double lp_ij = combine_logpost(0.0, lln * beta);      // model_i at theta_j
            const double lp_ji = combine_logpost(0.0, ll * beta_n);     // model_j at theta_i
            const double x = exp((lp_ji + lp_ij) - (lp + lpn));
            const double mhr = (x < 1.0) ? x : 1.0;                     // Python min(1, x): nan -> 1
            const bool sw = is_init && (uacc < mhr);                    // :121
            const bool sw_prev = __shfl_up_sync(0xffffffffu, (int)sw, 1, W) != 0;
            if (is_init) {
                lpq = lp_ij; acc = sw;
#pragma unroll
                for (int i = 0; i < D; ++i) q[i] = thn[i];
                if (sw) {
#pragma unroll
                    for (int i = 0; i < D; ++i) th[i] = thn[i];
                    ll = lln; lp = lp_ij;
                }
            } else if (is_part) {
                lpq = combine_logpost(0.0, llpv * beta); acc = sw_prev;
#pragma unroll
                for (int i = 0; i < D; ++i) q[i] = thp[i];
                if (sw_prev) {
#pragma unroll
                    for (int i = 0; i < D; ++i) th[i] = thp[i];
                    ll = llpv; lp = lpq;
                }
            }
        }
        if (!active) continue;
        if (tr.d_prop_theta) {
#pragma unroll
            for (int i = 0; i < D; ++i) tr.d_prop_theta[(t * K + c) * D + i] = q[i];
        }
        if (P.adapt) ad.update(moved, P.target);
        if (ACOV) haario_adapt<D>(P, th, acn, acsx, acsx2, acl);          // adapt(theta) after every step (sampler.py:88)
        dacc += (acc && regular) ? 1 : 0;
#pragma unroll
        for (int i = 0; i < D; ++i) { s1[i] += th[i]; s2[i] += th[i] * th[i]; }

        if (tr.d_prop_logpost) tr.d_prop_logpost[t * K + c] = lpq;
        if (tr.d_accepted) tr.d_accepted[t * K + c] = acc ? 1 : 0;
        if (tr.d_logqratio) tr.d_logqratio[t * K + c] = lqr;
        if (tr.d_theta || tr.d_logpost) {
            const long long r = ts.slot(t + 1);
            if (r >= 0) {
                if (tr.d_theta) {
#pragma unroll
                    for (int i = 0; i < D; ++i) tr.d_theta[(r * K + c) * D + i] = th[i];
                }
                if (tr.d_logpost) tr.d_logpost[r * K + c] = lp;
            }
        }
    }

    if (!active) return;
#pragma unroll
    for (int i = 0; i < D; ++i) {
        st.theta[(int64_t)i * K + c] = th[i];
        st.S1[(int64_t)i * K + c] += s1[i];
        st.S2[(int64_t)i * K + c] += s2[i];
    }
    st.lp[c] = lp;
    if (pcn_adapt) st.rho[c] = rho;
    if (ACOV) {
        st.acS[c] = acn;
#pragma unroll
        for (int i = 0; i < D; ++i) st.acSX[(int64_t)i * K + c] = acsx[ACOV ? i : 0];
#pragma unroll
        for (int i = 0; i < D * D; ++i) st.acSX2[(int64_t)i * K + c] = acsx2[ACOV ? i : 0];
#pragma unroll
        for (int i = 0; i < SGParams<D>::TRI; ++i) st.acL[(int64_t)i * K + c] = acl[ACOV ? i : 0];
    }
    if (pt) st.ll[c] = ll;
    st.scale[c] = ad.scale;
    st.nsamp[c] = ad.nsamples;
    st.nacc[c] = ad.naccepts;
    st.dacc[c] = dacc;
}

template <int D>
__global__ void sg_set_state_kernel(const SGParams<D>* __restrict__ gp, SGState st, int64_t K,
                                    const double* __restrict__ theta_in) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= K) return;
    double th[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
        th[i] = theta_in[c * D + i];
        st.theta[(int64_t)i * K + c] = th[i];
    }
    const double ll = gauss_loglik<D>(*gp, th);
    st.ll[c] = ll;
    st.lp[c] = combine_logpost(0.0, gp->nt > 0 ? ll * gp->betas[c % gp->nt] : ll);
}

template <int D>
__global__ void sg_get_state_kernel(SGState st, int64_t K, double* theta_out, double* lp_out) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= K) return;
    if (theta_out) {
#pragma unroll
        for (int i = 0; i < D; ++i) theta_out[c * D + i] = st.theta[(int64_t)i * K + c];
    }
    if (lp_out) lp_out[c] = st.lp[c];
}

template <int D>
__global__ void sg_get_acl_kernel(SGState st, int64_t K, double* L_out) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= K) return;
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j)
            L_out[(c * D + i) * D + j] = (j <= i) ? st.acL[(int64_t)(i * (i + 1) / 2 + j) * K + c] : 0.0;
}

__global__ void sg_get_adapt_kernel(SGState st, int64_t K, double* scale, int64_t* ns, int64_t* na) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= K) return;
    if (scale) scale[c] = st.scale[c];
    if (ns) ns[c] = st.nsamp[c];
    if (na) na[c] = st.nacc[c];
}

// The reference recomputes the covariance when `np.sqrt(n)**2 == n` (adaptive.py:82-83).  For numpy scalars
// `x**2` is libm's pow(x, 2.0), which is not always the correctly rounded product (n = 238, 952, ... differ), so the
// schedule is tabulated here on the host with the same libm call instead of being re-derived on the device.
static std::vector<uint32_t> haario_schedule(long long nmax) {
    std::vector<uint32_t> m((size_t)((nmax + 31) / 32), 0u);
    double (*volatile pw)(double, double) = pow;           // volatile: keep the compiler from folding pow(x, 2) to x*x
    for (long long n = 0; n < nmax; ++n) {
        const double x = (double)n;
        if (pw(sqrt(x), 2.0) == x) m[(size_t)(n >> 5)] |= 1u << (n & 31);
    }
    return m;
}

static void pack_lower(const std::vector<double>& full, int d, double* out) {
    for (int i = 0; i < d; ++i)
        for (int j = 0; j <= i; ++j) out[i * (i + 1) / 2 + j] = full[(size_t)i * d + j];
}

template <int D>
struct SmallGaussSampler : SamplerImpl {
    rmn_sampler* s;
    SGState st{};
    SGParams<D>* d_params = nullptr;
    SGParams<D> hparams{};
    uint32_t* d_acmask = nullptr;
    explicit SmallGaussSampler(rmn_sampler* s_) : s(s_) {}
    ~SmallGaussSampler() override { if (d_params) cudaFree(d_params); if (d_acmask) cudaFree(d_acmask); }

    size_t workspace_bytes() const override {
        const size_t K = (size_t)s->K;
        size_t n = align256(D * K * 8) * 3 + align256(K * 8) * 7 + 256;
        if (s->prop->acov) n += align256(K * 8) + align256(D * K * 8) + align256(D * D * K * 8) + align256(SGParams<D>::TRI * K * 8);
        return n;
    }
    int bind(void* ws) override {
        const size_t K = (size_t)s->K;
        char* p = (char*)ws;
        st.theta = (double*)p; p += align256(D * K * 8);
        st.S1 = (double*)p; p += align256(D * K * 8);
        st.S2 = (double*)p; p += align256(D * K * 8);
        st.lp = (double*)p; p += align256(K * 8);
        st.ll = (double*)p; p += align256(K * 8);
        st.scale = (double*)p; p += align256(K * 8);
        st.nsamp = (long long*)p; p += align256(K * 8);
        st.nacc = (long long*)p; p += align256(K * 8);
        st.dacc = (long long*)p; p += align256(K * 8);
        st.rho = (double*)p; p += align256(K * 8);
        if (s->prop->acov) {
            st.acS = (double*)p; p += align256(K * 8);
            st.acSX = (double*)p; p += align256(D * K * 8);
            st.acSX2 = (double*)p; p += align256(D * D * K * 8);
            st.acL = (double*)p; p += align256(SGParams<D>::TRI * K * 8);
        }

        SGParams<D> h{};
        const rmn_model* m = s->model;
        const rmn_proposal* pr = s->prop;
        for (int i = 0; i < D; ++i) h.mu[i] = m->h_mu[i];
        for (int i = 0; i < SGParams<D>::TRI; ++i) h.linv[i] = m->h_linv[i];
        h.c1 = D * log(2.0 * M_PI);
        h.c2 = m->logdetC;
        h.kind = pr->kind; h.adapt = pr->adapt; h.target = pr->target;
        h.nsteps = pr->nsteps; h.has_mass = pr->has_mass ? 1 : 0;
        h.rho = pr->rho; h.rho_c = sqrt(1.0 - pr->rho * pr->rho);
        h.eps0 = pr->eps;
        if (!pr->h_L.empty()) pack_lower(pr->h_L, D, h.lprop);
        if (!pr->h_Linv.empty()) pack_lower(pr->h_Linv, D, h.lpinv);
        if (pr->has_mass) {
            pack_lower(pr->h_chM, D, h.chM);
            pack_lower(pr->h_chMinv, D, h.chMinv);
            for (int i = 0; i < D * D; ++i) h.Minv[i] = pr->h_Minv[i];
        }
        if (pr->acov) {
            h.acov = 1; h.marginalize = pr->ac_marginalize; h.smooth = pr->ac_smooth; h.t_adapt = pr->ac_t_adapt;
            double tr = 0.0;
            for (int i = 0; i < D; ++i) tr += pr->h_C0[(size_t)i * D + i];
            h.creg = tr / D;
            h.dpow04 = pow((double)D, 0.4); h.dpow02 = pow((double)D, 0.2);
            for (int i = 0; i < D * D; ++i) h.c0[i] = pr->h_C0[i];
            const long long nmax = 1ll << 22;                  // 4M states tabulated; beyond: correctly rounded r*r
            const std::vector<uint32_t> mask = haario_schedule(nmax);
            RMN_CUDA(cudaMalloc(&d_acmask, mask.size() * 4));
            RMN_CUDA(cudaMemcpy(d_acmask, mask.data(), mask.size() * 4, cudaMemcpyHostToDevice));
            h.acmask = d_acmask; h.acmask_n = nmax;
        }
        hparams = h;
        RMN_CUDA(cudaMalloc(&d_params, sizeof(h)));
        RMN_CUDA(cudaMemcpy(d_params, &h, sizeof(h), cudaMemcpyHostToDevice));
        RMN_CUDA(cudaMemset(ws, 0, workspace_bytes()));
        if (int rc3 = rmn_fill_f64(st.rho, s->K, pr->rho, 0)) return rc3;
        if (pr->acov) {                                   // every chain starts from chol(C0) (adaptive.py:61)
            for (int q = 0; q < SGParams<D>::TRI; ++q) {
                int rc2 = rmn_fill_f64(st.acL + (size_t)q * K, s->K, h.lprop[q], 0);
                if (rc2) return rc2;
            }
        }
        int rc = rmn_fill_f64(st.scale, s->K, 1.0, 0);
        if (rc) return rc;
        RMN_CUDA(cudaDeviceSynchronize());
        return RMN_OK;
    }
    unsigned grid() const { return (unsigned)((s->K + 127) / 128); }
    // tempered runs: one ladder per `stride` threads
    unsigned run_grid() const {
        if (hparams.nt == 0) return grid();
        const int64_t threads = (s->K / hparams.nt) * hparams.stride;
        return (unsigned)((threads + 127) / 128);
    }
    // PTSampler (ptsampler.py:41-81): chains c = l*nt + i form ladder l with beta_i; call before set_state
    int set_tempering(int nt, const double* betas, double pswap) override {
        RMN_REQUIRE(nt >= 2 && nt <= 32 && betas, "set_tempering: need 2 <= nt <= 32 temperatures");
        RMN_REQUIRE(pswap > 0.0 && pswap < 1.0, "Pswap must be a number between 0 and 1");
        RMN_REQUIRE(s->K % nt == 0, "set_tempering: the number of chains (%lld) must be a multiple of nt = %d", (long long)s->K, nt);
        RMN_REQUIRE(s->chain_offset % nt == 0, "set_tempering: chain_offset must be a multiple of nt");
        RMN_REQUIRE(hparams.acov == 0, "parallel tempering supports non-adaptive proposals only");
        RMN_REQUIRE(hparams.adapt == 0, "parallel tempering supports non-adaptive proposals only (the reference shares ONE "
                                        "proposal object between all temperatures, ptsampler.py:81)");
        RMN_REQUIRE(hparams.kind != RMN_PROP_HMC, "parallel tempering: the HMC gradient is not tempered in the reference; use RW or pCN");
        for (int i = 0; i < nt; ++i) {
            RMN_REQUIRE(betas[i] >= 0.0 && betas[i] <= 1.0, "beta = %g must be a number between 0 and 1", betas[i]);
            hparams.betas[i] = betas[i];
        }
        int stride = 1;
        while (stride < nt) stride *= 2;
        hparams.nt = nt; hparams.stride = stride; hparams.pswap = pswap;
        RMN_CUDA(cudaMemcpy(d_params, &hparams, sizeof(hparams), cudaMemcpyHostToDevice));
        return RMN_OK;
    }

    int set_state(const double* d_theta, cudaStream_t stream) override {
        sg_set_state_kernel<D><<<grid(), 128, 0, stream>>>(d_params, st, s->K, d_theta);
        RMN_KERNEL_CHECK(); launches++;
        return RMN_OK;
    }
    int get_state(double* d_theta, double* d_lp, cudaStream_t stream) override {
        sg_get_state_kernel<D><<<grid(), 128, 0, stream>>>(st, s->K, d_theta, d_lp);
        RMN_KERNEL_CHECK(); launches++;
        return RMN_OK;
    }
    int run(int64_t T, const rmn_inject_t* inj, const rmn_trace_t* tr, cudaStream_t stream) override {
        rmn_trace_t t0{};
        if (tr) t0 = *tr;
        if (t0.thin <= 0) t0.thin = 1;
        if (inj) RMN_REQUIRE(inj->d_xi && inj->d_u, "injected run needs d_xi and d_u");
        if (inj && hparams.nt > 0) RMN_REQUIRE(inj->d_usel, "injected tempered run needs d_usel (selection uniforms)");
        ktimer.begin("small_gauss_kernel", stream);
        const bool ac = hparams.acov != 0, ptm = hparams.nt > 0;
#define RMN_SG_LAUNCH(INJ_, AC_, PT_, XI_, U_, US_)                                                      \
        small_gauss_kernel<D, INJ_, AC_, PT_><<<run_grid(), 128, 0, stream>>>(                             \
            d_params, st, s->K, T, step0, s->seed, s->chain_offset, XI_, U_, US_, t0)
        if (inj) {
            if (ptm) RMN_SG_LAUNCH(true, false, true, inj->d_xi, inj->d_u, inj->d_usel);
            else if (ac) RMN_SG_LAUNCH(true, true, false, inj->d_xi, inj->d_u, nullptr);
            else RMN_SG_LAUNCH(true, false, false, inj->d_xi, inj->d_u, nullptr);
        } else {
            if (ptm) RMN_SG_LAUNCH(false, false, true, nullptr, nullptr, nullptr);
            else if (ac) RMN_SG_LAUNCH(false, true, false, nullptr, nullptr, nullptr);
            else if (hparams.kind == RMN_PROP_RW)
                small_gauss_kernel<D, false, false, false, true><<<run_grid(), 128, 0, stream>>>(
                    d_params, st, s->K, T, step0, s->seed, s->chain_offset, nullptr, nullptr, nullptr, t0);
            else RMN_SG_LAUNCH(false, false, false, nullptr, nullptr, nullptr);
        }
#undef RMN_SG_LAUNCH
        ktimer.end(stream);
        RMN_KERNEL_CHECK(); launches++;
        step0 += T; diag_steps += T;
        return RMN_OK;
    }
    int get_adapt(double* sc, int64_t* ns, int64_t* na, cudaStream_t stream) override {
        sg_get_adapt_kernel<<<grid(), 128, 0, stream>>>(st, s->K, sc, ns, na);
        RMN_KERNEL_CHECK(); launches++;
        return RMN_OK;
    }
    int set_adapt(const double* sc, const int64_t* ns, const int64_t* na, cudaStream_t stream) override {
        launches++;
        return rmn_copy_adapt(s->K, sc, ns, na, st.scale, st.nsamp, st.nacc, stream);
    }
    // the adapted proposal factor of every chain, full lower-triangular [K][D][D] (AdaptCovProposal.L)
    int get_adaptcov(double* d_L, cudaStream_t stream) override {
        RMN_REQUIRE(hparams.acov, "this proposal does not adapt its covariance");
        sg_get_acl_kernel<D><<<grid(), 128, 0, stream>>>(st, s->K, d_L);
        RMN_KERNEL_CHECK(); launches++;
        return RMN_OK;
    }
    int diag_dim() const override { return D; }
    int reset_diag(cudaStream_t stream) override {
        RMN_CUDA(cudaMemsetAsync(st.S1, 0, (size_t)D * s->K * 8, stream));
        RMN_CUDA(cudaMemsetAsync(st.S2, 0, (size_t)D * s->K * 8, stream));
        RMN_CUDA(cudaMemsetAsync(st.dacc, 0, (size_t)s->K * 8, stream));
        diag_steps = 0;
        return RMN_OK;
    }
    int chain_sums(const double** S1, const double** S2, int64_t* n) override {
        *S1 = st.S1; *S2 = st.S2; *n = diag_steps;
        return RMN_OK;
    }
    int reduce_diag(double* d_block, cudaStream_t stream) override {
        launches++;
        return rmn_reduce_diag_block(s->K, D, diag_steps, diag_steps, st.S1, st.S2, st.dacc, nullptr, d_block, stream);
    }
};

}  // namespace

SamplerImpl* make_small_gauss_sampler(rmn_sampler* s) {
    switch (s->model->d) {
        case 1: return new SmallGaussSampler<1>(s);
        case 2: return new SmallGaussSampler<2>(s);
        case 3: return new SmallGaussSampler<3>(s);
        case 4: return new SmallGaussSampler<4>(s);
        case 5: return new SmallGaussSampler<5>(s);
        case 6: return new SmallGaussSampler<6>(s);
        case 7: return new SmallGaussSampler<7>(s);
        case 8: return new SmallGaussSampler<8>(s);
        default: return nullptr;
    }
}
