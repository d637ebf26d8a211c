// riemann_b200 -- fused T-step MH kernel for the changepoint regression model.
//
// RMN_CP_LANES (=16) lanes cooperate on one chain: lane j holds cpx[j] and cpv[j] in
// registers, so the variable-dimension state (k <= LANES-1 changepoints) needs no
// dynamic indexing; insert/delete are warp shuffles.  Per iteration the kernel replaces
//   Sampler.sample                                riemann/samplers/sampler.py:72-90
//   ChangepointRegression1DProp.propose           examples/test_changepoint.py:44-73
//   MetropolisRandomWalk.propose (3 block moves)  riemann/proposals/randomwalk.py:21-26
//   add_changepoint / subtract_changepoint        riemann/models/changepoint.py:193-240
//   ChangepointRegression1D.log_likelihood/log_prior/predict   changepoint.py:106-160,181
//   Model.log_posterior                           riemann/models/model.py:43-55
//
// Likelihood: x is sorted, so the prediction is constant on runs of data points; with
// prefix sums cy[i] = sum_{m<i} (y_m - c), cyy[i] = sum_{m<i} (y_m - c)^2 the residual
// sum of squares of segment j is  n_j (v_j-c)^2 - 2 (v_j-c) S1_j + S2_j,  where the run
// boundaries b_j = #{x_i <= cpx_j} come from a branch-free binary search (equivalent to
// np.searchsorted(cpx, x), changepoint.py:181).  O(k log M) instead of O(M) per
// evaluation; all arithmetic fp64 like the reference.  x/cy/cyy are staged in shared
// memory once per block and reused by all chains and all T iterations.
//
// Reference behaviours reproduced: k = number of STEPS in
// the three k-terms of the prior; constraints enforced only through nan/inf -> -inf;
// a fresh uniform per block-selection elif; trans-dimensional moves use log|J| alone
// as logqratio.
#include "changepoint.cuh"

#ifdef RMN_WITH_TPC   /* experiments/changepoint_tpc.cu linked in (not part of the product build) */
#include <stdlib.h>
#include <string.h>
namespace cp {
size_t tpc_smem_bytes(int M);
void tpc_launch(bool inj, const CPParams& P, const double* gdata, const CPState& st, int64_t K, int64_t T,
                int64_t step0, uint64_t seed, int64_t chain_offset, const double* tape, const rmn_trace_t& tr,
                cudaStream_t stream);
}
#endif

namespace {
using namespace cp;

// ---------------------------------------------------------------------------------------
// Log-posterior of one state held across the 16 lanes of a group, in two stages so that the
// kernel can batch ALL of a step's fp64 logarithms into one (rarely two) `log` calls:
//   stage 1 (cp_segments): binary search, per-lane segment residual, gap and height term;
//   the caller takes log(gap) per lane together with the step's scalar logs
//   (log sigma^2, log 1/sigma^2, log u, log|J|) computed on otherwise idle lanes;
//   stage 2 (cp_combine): 16-lane reductions and the reference's term-by-term sums.
// Plain IEEE arithmetic: log of a negative gap/height gives nan, log(0) gives -inf, exactly
// like numpy; the reference's nan -> -inf and inf/nan -> -inf rules are applied at the end.
// ---------------------------------------------------------------------------------------
struct CPSeg { double ss, gap, vt; };

__device__ __forceinline__ CPSeg cp_segments(const CPParams& P, const double* __restrict__ cy,
                                             const double* __restrict__ cyy, int lane, int k,
                                             double cx, double cv_, int bu) {
    // bu = #{x_i <= cpx[lane]} for lane < k, M otherwise (upper end of this lane's run of data)
    int bl = __shfl_up_sync(0xffffffffu, bu, 1, LANES);
    if (lane == 0) bl = 0;
    double prev = __shfl_up_sync(0xffffffffu, cx, 1, LANES);
    if (lane == 0) prev = P.xmin;
    const double hi = (lane < k) ? cx : P.xmax;
    CPSeg o;
    o.ss = 0.0; o.vt = 0.0; o.gap = 1.0;                       // log(1) = 0 on inactive lanes
    if (lane <= k) {
        const double n = (double)(bu - bl);
        const double s1 = cy[bu] - cy[bl];
        const double s2 = cyy[bu] - cyy[bl];
        const double vc = cv_ - P.ycenter;
        o.ss = n * vc * vc - 2.0 * vc * s1 + s2;
        o.gap = hi - prev;                                     // changepoint.py:142-143
        o.vt = -P.beta * cv_ + P.cv;                           // changepoint.py:18-19,136
        if (!P.alpha_is_one) o.vt += (P.alpha - 1.0) * log(cv_);
        else if (!(cv_ > 0.0)) o.vt = NAN;                     // 0 * log(v<=0) is nan in numpy
    }
    return o;
}

// lg = per-lane log(gap) (0 on inactive lanes); log_s2 = log(sigma^2); lsig = log(1/sigma^2)
__device__ __forceinline__ double cp_combine(const CPParams& P, int k, double sig, CPSeg sg, double lg,
                                             double log_s2, double lsig, int which) {
    const double ss = group_sum<LANES>(sg.ss);
    lg = group_sum<LANES>(lg);
    const double vt = group_sum<LANES>(sg.vt);
    const int ks = k + 1;                                      // number of steps
    const double s2v = sig * sig;
    double logl = -0.5 * ((ss / s2v + (double)P.M * log_s2) + P.Mlog2pi);   // :118-120
    if (isnan(logl)) logl = -INFINITY;                         // :124-125
    if (sig < 0.0) lsig = NAN;                                 // :146-147
    const double lps = (P.tab2[ks] + lg) - (double)ks * P.logL;
    double logp = ((P.tab1[ks] + vt) + lps) + lsig;            // :156
    if (isnan(logp)) logp = -INFINITY;                         // :158-159
    if (which == 1) return logl;
    if (which == 2) return logp;
    return combine_logpost(logp, logl);
}

// stand-alone evaluation (set_state / pointwise): scalar logs ride on lanes 14, 15 when free
__device__ __forceinline__ double cp_logpost(const CPParams& P, const double* __restrict__ xs,
                                             const double* __restrict__ cy,
                                             const double* __restrict__ cyy, int lane, int k,
                                             double cx, double cv_, double sig, int which) {
    const int bu = (lane < k) ? upper_bound(xs, P.M, P.P2, cx) : P.M;
    const CPSeg sg = cp_segments(P, cy, cyy, lane, k, cx, cv_, bu);
    const double s2v = sig * sig;
    const double lg = log(sg.gap);
    const double sl = log((lane & 1) ? 1.0 / s2v : s2v);
    const double log_s2 = __shfl_sync(0xffffffffu, sl, 0, LANES);
    const double lsig = __shfl_sync(0xffffffffu, sl, 1, LANES);
    return cp_combine(P, k, sig, sg, (lane <= k) ? lg : 0.0, log_s2, lsig, which);
}

template <bool INJ, bool SMEMDATA, bool DIAG>
#ifndef RMN_CP_MINBLOCKS
#define RMN_CP_MINBLOCKS 8   /* 64 regs/thread, 32 warps/SM: measured +3.5 % over 6 */
#endif
__global__ void __launch_bounds__(128, RMN_CP_MINBLOCKS)
changepoint_kernel(const __grid_constant__ CPParams P, const double* __restrict__ gdata, CPState st,
                   int64_t K, int64_t T, int64_t step0, uint64_t seed, int64_t chain_offset,
                   const double* __restrict__ tape, rmn_trace_t tr) {
    extern __shared__ double smem[];
    const double* xs = gdata;
    if (SMEMDATA) {
        const int n = 3 * P.M + 2;
        for (int i = threadIdx.x; i < n; i += blockDim.x) smem[i] = gdata[i];
        __syncthreads();
        xs = smem;
    }
    const double* cy = xs + P.M;
    const double* cyy = cy + P.M + 1;

    const int lane = threadIdx.x & (LANES - 1);
    const int64_t c_raw = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LANES;
    const bool live = c_raw < K;
    const int64_t c = live ? c_raw : K - 1;       // dead groups shadow the last chain, never store

    int k = st.k[c];
    double cx = st.cpx[c * LANES + lane];
    double cv = st.cpv[c * LANES + lane];
    double sig = st.sig[c];
    double lp = st.lp[c];
    int bu = (lane < k) ? upper_bound(xs, P.M, P.P2, cx) : P.M;   // cached run boundaries of the state
    long long nacc = 0, novf = 0;
    double s1 = 0.0, s2 = 0.0;                    // lane i < NDIAG accumulates functional i
    const RngKey rk(seed, (uint64_t)(chain_offset + c));
    const TraceSel ts{tr.first, tr.thin > 0 ? tr.thin : 1};
    const bool tracing = tr.d_k || tr.d_cpx || tr.d_cpv || tr.d_sig || tr.d_logpost;

    for (int64_t t = 0; t < T; ++t) {
        const uint64_t step = (uint64_t)(step0 + t);
        double snew, du, uacc, xi;
        int nrand, mv;
        bool birth;
        if (INJ) {
            const double* row = tape + (t * K + c) * RMN_CP_NSLOT;
            const double u1 = row[RMN_CP_SLOT_SEL1], u2 = row[RMN_CP_SLOT_SEL2], u3 = row[RMN_CP_SLOT_SEL3];
            const double ubd = row[RMN_CP_SLOT_BD];
            snew = row[RMN_CP_SLOT_S]; du = row[RMN_CP_SLOT_DU];
            nrand = (int)row[RMN_CP_SLOT_N]; uacc = row[RMN_CP_SLOT_ACC];
            xi = row[RMN_CP_SLOT_XI + lane];
            // which block moves (fresh uniform per elif, test_changepoint.py:48-54); birth/death :59
            mv = (u1 < P.p1) ? 0 : ((u2 < P.p2) ? 1 : ((u3 < P.p3) ? 2 : 3));
            birth = (k == 0) || (ubd > 0.5);
        } else {
            // ONE Philox block per lane and step: words x,y -> this lane's normal; the spare
            // words z,w of lanes 0..3 carry the chain-level uniforms
            const uint4 r = rk.block(step, (uint32_t)lane);
            float n0, n1;
            box_muller(r.x, r.y, n0, n1);
            xi = (double)n0;
            const uint32_t z0 = __shfl_sync(0xffffffffu, r.z, 0, LANES), w0 = __shfl_sync(0xffffffffu, r.w, 0, LANES);
            const uint32_t z1 = __shfl_sync(0xffffffffu, r.z, 1, LANES), w1 = __shfl_sync(0xffffffffu, r.w, 1, LANES);
            const uint32_t z2 = __shfl_sync(0xffffffffu, r.z, 2, LANES), w2 = __shfl_sync(0xffffffffu, r.w, 2, LANES);
            const uint32_t z3 = __shfl_sync(0xffffffffu, r.z, 3, LANES), w3 = __shfl_sync(0xffffffffu, r.w, 3, LANES);
            mv = (z0 < P.t1) ? 0 : ((w0 < P.t2) ? 1 : ((z1 < P.t3) ? 2 : 3));   // same tests on raw words
            birth = (k == 0) || (w1 >= 0x80000000u);                            // u > 0.5
            snew = P.xmin + (P.xmax - P.xmin) * u01_fast(z2);
            du = -0.1 + 0.2 * u01_fast(w2);
            nrand = (int)(u01_fast(z3) * (double)k);
            uacc = u01_fast(w3);
        }
        nrand = max(0, min(nrand, k - 1));

        // ---- build the proposal by selection (the two chains of a warp never diverge on mv)
        int kk = k, nbu = bu;
        double nx = cx, nv = cv, nsig = sig, jarg = 1.0;
        bool ovf = false;
        if (mv == 0) {
            if (lane < k) nx = __dadd_rn(cx, __dmul_rn(P.sx[k], xi));   // randomwalk.py:26, scale = 1
        } else if (mv == 1) {
            if (lane <= k) nv = __dadd_rn(cv, __dmul_rn(P.sv, xi));
        }
        const double xi0 = __shfl_sync(0xffffffffu, xi, 0, LANES);
        if (mv == 2) nsig = __dadd_rn(sig, __dmul_rn(P.ss, xi0));
        if (__any_sync(0xffffffffu, mv == 3)) {                         // warp-uniform (35 % of steps)
            const double px = __shfl_up_sync(0xffffffffu, cx, 1, LANES);
            const double pv = __shfl_up_sync(0xffffffffu, cv, 1, LANES);
            const double qx = __shfl_down_sync(0xffffffffu, cx, 1, LANES);
            const double qv = __shfl_down_sync(0xffffffffu, cv, 1, LANES);
            const int nb = __popc(group_ballot(lane < k && cx < snew)); // searchsorted(cpx, s), :206
            const double hb = __shfl_sync(0xffffffffu, cv, nb, LANES);
            const double h1d = __shfl_sync(0xffffffffu, cv, nrand, LANES);
            const double h2d = __shfl_sync(0xffffffffu, cv, nrand + 1, LANES);
            const int pbu = __shfl_up_sync(0xffffffffu, bu, 1, LANES);
            const int qbu = __shfl_down_sync(0xffffffffu, bu, 1, LANES);
            if (mv == 3) {
                if (birth) {
                    const double u = 0.5 + du / P.sqrtM;                // :61
                    const double f = sqrt((1.0 - u) / u);               // changepoint.py:57
                    jarg = fabs(hb / (u * (1.0 - u)));                  // |J|, :72-74
                    if (k + 1 > LANES - 1) {
                        ovf = true;
                    } else {
                        nx = (lane < nb) ? cx : ((lane == nb) ? snew : px);
                        nv = (lane < nb) ? cv : ((lane == nb) ? hb / f : ((lane == nb + 1) ? hb * f : pv));
                        nbu = (lane <= nb) ? bu : pbu;          // lane nb is searched below
                        kk = k + 1;
                    }
                } else {
                    const double h = sqrt(h1d * h2d);                   // changepoint.py:67
                    const double u = 1.0 / (1.0 + h2d / h1d);           // :68
                    jarg = fabs(h / (u * (1.0 - u)));                   // 1/|J^-1|, :76-78
                    nx = (lane < nrand) ? cx : qx;
                    nv = (lane < nrand) ? cv : ((lane == nrand) ? h : qv);
                    nbu = (lane < nrand) ? bu : qbu;
                    kk = k - 1;
                }
            }
        }
        // lanes beyond the new extent hold zeros (canonical padding)
        if (lane >= kk) { nx = 0.0; nbu = P.M; }
        if (lane > kk) nv = 0.0;
        // run boundaries only move when a location moves: a cpx block move (every lane) or a
        // birth (the new lane).  Warp-uniform, taken on about half of the steps.
        const bool moved_x = (mv == 0) || (mv == 3 && birth && !ovf);
        if (__any_sync(0xffffffffu, moved_x)) {
            const int sb = upper_bound(xs, P.M, P.P2, nx);
            if (moved_x && lane < kk) nbu = sb;
        }

        // ---- log-posterior of the proposal; every fp64 log of the step in one call:
        //      lanes 0..kk take log(gap); lanes 12..15 take log sigma^2, log 1/sigma^2, log u, log|J|
        //      (if a chain has more than 11 changepoints those four go through a second call)
        const CPSeg sg = cp_segments(P, cy, cyy, lane, kk, nx, nv, nbu);
        const double s2n = nsig * nsig;
        const int sl = lane & 3;
        const double sarg = (sl == 0) ? s2n : ((sl == 1) ? 1.0 / s2n : ((sl == 2) ? uacc : jarg));
        const bool crowded = kk > 11;
        const double l1 = log((lane >= 12 && !crowded) ? sarg : sg.gap);
        double l2 = 0.0;
        if (__any_sync(0xffffffffu, crowded)) l2 = log(sarg);
        const double lsrc = crowded ? l2 : l1;
        const int sbase = crowded ? 0 : 12;
        const double log_s2 = __shfl_sync(0xffffffffu, lsrc, sbase + 0, LANES);
        const double lsig = __shfl_sync(0xffffffffu, lsrc, sbase + 1, LANES);
        const double logu = __shfl_sync(0xffffffffu, lsrc, sbase + 2, LANES);
        const double ljac = __shfl_sync(0xffffffffu, lsrc, sbase + 3, LANES);
        const double lpn = cp_combine(P, kk, nsig, sg, (lane <= kk) ? l1 : 0.0, log_s2, lsig, 0);
        const double lqr = (mv == 3) ? (birth ? ljac : -ljac) : 0.0;

        // sampler.py:83-84 with Python's min(0, nan) == 0
        const double delta = lpn - lp - lqr;
        const double mh = (delta < 0.0) ? delta : 0.0;
        const bool acc = !ovf && (logu < mh);
        if (acc) { k = kk; cx = nx; cv = nv; sig = nsig; lp = lpn; bu = nbu; }
        nacc += acc ? 1 : 0;
        novf += ovf ? 1 : 0;

        if (DIAG && (step % RMN_CP_DIAG_EVERY) == 0) {      // thinned accumulation (warp-uniform)
            double f = (lane == 0) ? sig : (double)k;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int cnt = __popc(group_ballot(lane < k && cx < P.xq[q]));
                const double yq = __shfl_sync(0xffffffffu, cv, cnt, LANES);
                if (lane == 2 + q) f = yq;
            }
            if (lane < RMN_CP_NDIAG) { s1 += f; s2 += f * f; }
        }

        if (live) {
            if (lane == 0) {
                if (tr.d_prop_logpost) tr.d_prop_logpost[t * K + c] = lpn;
                if (tr.d_accepted) tr.d_accepted[t * K + c] = acc ? 1 : 0;
                if (tr.d_logqratio) tr.d_logqratio[t * K + c] = lqr;
                if (tr.d_prop_k) tr.d_prop_k[t * K + c] = kk;
                if (tr.d_prop_sig) tr.d_prop_sig[t * K + c] = nsig;
            }
            if (tr.d_prop_cpx) tr.d_prop_cpx[(t * K + c) * LANES + lane] = nx;
            if (tr.d_prop_cpv) tr.d_prop_cpv[(t * K + c) * LANES + lane] = nv;
            if (tracing) {
                const long long r = ts.slot(t + 1);
                if (r >= 0) {
                    if (tr.d_cpx) tr.d_cpx[(r * K + c) * LANES + lane] = cx;
                    if (tr.d_cpv) tr.d_cpv[(r * K + c) * LANES + lane] = cv;
                    if (lane == 0) {
                        if (tr.d_k) tr.d_k[r * K + c] = k;
                        if (tr.d_sig) tr.d_sig[r * K + c] = sig;
                        if (tr.d_logpost) tr.d_logpost[r * K + c] = lp;
                    }
                }
            }
        }
    }

    if (live) {
        st.cpx[c * LANES + lane] = cx;
        st.cpv[c * LANES + lane] = cv;
        if (lane == 0) {
            st.k[c] = k;
            st.sig[c] = sig;
            st.lp[c] = lp;
            st.dacc[c] += nacc;
            st.dovf[c] += novf;
        }
        if (DIAG && lane < RMN_CP_NDIAG) {
            st.S1[(int64_t)lane * K + c] += s1;
            st.S2[(int64_t)lane * K + c] += s2;
        }
    }
}

// evaluate n states given in the canonical layout (pointwise parity entry + set_state)
__global__ void __launch_bounds__(128)
cp_eval_kernel(const __grid_constant__ CPParams P, const double* __restrict__ gdata, int which,
               int64_t n, const int32_t* __restrict__ kin, const double* __restrict__ cpx,
               const double* __restrict__ cpv, const double* __restrict__ sig,
               double* __restrict__ out) {
    const int lane = threadIdx.x & (LANES - 1);
    const int64_t c_raw = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LANES;
    const bool live = c_raw < n;
    const int64_t c = live ? c_raw : n - 1;
    const double* xs = gdata;
    const double* cy = xs + P.M;
    const double* cyy = cy + P.M + 1;
    const double v = cp_logpost(P, xs, cy, cyy, lane, kin[c], cpx[c * LANES + lane],
                                cpv[c * LANES + lane], sig[c], which);
    if (live && lane == 0) out[c] = v;
}

__global__ void cp_copy_state_kernel(int64_t K, const int32_t* k_in, const double* cpx_in,
                                     const double* cpv_in, const double* sig_in, int32_t* k_out,
                                     double* cpx_out, double* cpv_out, double* sig_out,
                                     const double* lp_in, double* lp_out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= K * LANES) return;
    const int64_t c = i / LANES;
    const int lane = (int)(i % LANES);
    const int k = k_in[c];
    if (cpx_out) cpx_out[i] = (lane < k) ? cpx_in[i] : 0.0;
    if (cpv_out) cpv_out[i] = (lane <= k) ? cpv_in[i] : 0.0;
    if (lane == 0) {
        if (k_out) k_out[c] = k;
        if (sig_out) sig_out[c] = sig_in[c];
        if (lp_out && lp_in) lp_out[c] = lp_in[c];
    }
}

static CPParams make_params(const rmn_model* m, const rmn_proposal* p) {
    CPParams P{};
    P.M = m->M;
    int p2 = 1;
    while (p2 * 2 <= m->M) p2 *= 2;
    P.P2 = (m->M >= 1) ? p2 : 0;
    P.alpha_is_one = (m->alpha == 1.0);
    P.xmin = m->xmin; P.xmax = m->xmax; P.alpha = m->alpha; P.beta = m->beta;
    P.cv = m->cv;
    P.logL = log(m->xmax - m->xmin);
    P.ycenter = m->ycenter;
    P.Mlog2pi = (double)m->M * log(2.0 * M_PI);
    P.sqrtM = sqrt((double)m->M);
    for (int ks = 0; ks <= LANES; ++ks) {
        P.tab1[ks] = (ks >= 1) ? (double)ks * log(m->lamb) - lgamma((double)ks) - m->lamb : 0.0;
        P.tab2[ks] = lgamma(2.0 * ks + 1.0);
    }
    const double hs = p ? p->hscale : 1.0;
    for (int k = 0; k <= LANES; ++k) P.sx[k] = sqrt(0.01 * (m->xmax - m->xmin) / (double)(k + 1));
    P.sv = sqrt(0.01 * (hs * hs) / (double)m->M);
    P.ss = sqrt(0.01 * hs);
    P.p1 = p ? p->p_cum[0] : 0.2; P.p2 = p ? p->p_cum[1] : 0.4; P.p3 = p ? p->p_cum[2] : 0.6;
    // (x + 0.5) 2^-32 < p  <=>  x < ceil(p 2^32 - 0.5)
    auto thr = [](double pc) {
        const double v = ceil(pc * 4294967296.0 - 0.5);
        return (uint32_t)(v <= 0.0 ? 0.0 : (v >= 4294967295.0 ? 4294967295.0 : v));
    };
    P.t1 = thr(P.p1); P.t2 = thr(P.p2); P.t3 = thr(P.p3); P.tpad = 0;
    for (int q = 0; q < NQ; ++q)
        P.xq[q] = m->xmin + (m->xmax - m->xmin) * (q + 0.5) / (double)NQ;
    return P;
}

struct ChangepointSampler : SamplerImpl {
    rmn_sampler* s;
    CPState st{};
    CPParams P;
    bool use_smem;
    size_t smem_bytes;
    explicit ChangepointSampler(rmn_sampler* s_) : s(s_) {
        P = make_params(s->model, s->prop);
        smem_bytes = (size_t)(3 * P.M + 2) * 8;
        use_smem = smem_bytes <= 96 * 1024;
    }
    size_t workspace_bytes() const override {
        const size_t K = (size_t)s->K;
        return align256(K * 4) + 2 * align256(K * LANES * 8) + 4 * align256(K * 8) +
               2 * align256(RMN_CP_NDIAG * K * 8) + 256;
    }
    int bind(void* ws) override {
        const size_t K = (size_t)s->K;
        char* p = (char*)ws;
        st.k = (int32_t*)p; p += align256(K * 4);
        st.cpx = (double*)p; p += align256(K * LANES * 8);
        st.cpv = (double*)p; p += align256(K * LANES * 8);
        st.sig = (double*)p; p += align256(K * 8);
        st.lp = (double*)p; p += align256(K * 8);
        st.dacc = (long long*)p; p += align256(K * 8);
        st.dovf = (long long*)p; p += align256(K * 8);
        st.S1 = (double*)p; p += align256(RMN_CP_NDIAG * K * 8);
        st.S2 = (double*)p; p += align256(RMN_CP_NDIAG * K * 8);
        RMN_CUDA(cudaMemset(ws, 0, workspace_bytes()));
        if (use_smem && smem_bytes > 48 * 1024) {
            RMN_CUDA(cudaFuncSetAttribute(changepoint_kernel<false, true, true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
            RMN_CUDA(cudaFuncSetAttribute(changepoint_kernel<true, true, true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        }
        return RMN_OK;
    }
    unsigned grid() const { return (unsigned)((s->K * LANES + 127) / 128); }

    int cp_set_state(const int32_t* d_k, const double* d_cpx, const double* d_cpv,
                     const double* d_sig, cudaStream_t stream) override {
        const int64_t n = s->K * LANES;
        cp_copy_state_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(
            s->K, d_k, d_cpx, d_cpv, d_sig, st.k, st.cpx, st.cpv, st.sig, nullptr, nullptr);
        RMN_KERNEL_CHECK();
        cp_eval_kernel<<<grid(), 128, 0, stream>>>(P, s->model->d_cpdata, 0, s->K, st.k, st.cpx,
                                                   st.cpv, st.sig, st.lp);
        RMN_KERNEL_CHECK();
        launches += 2;
        return RMN_OK;
    }
    int cp_get_state(int32_t* d_k, double* d_cpx, double* d_cpv, double* d_sig, double* d_lp,
                     cudaStream_t stream) override {
        const int64_t n = s->K * LANES;
        cp_copy_state_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(
            s->K, st.k, st.cpx, st.cpv, st.sig, d_k, d_cpx, d_cpv, d_sig, st.lp, d_lp);
        RMN_KERNEL_CHECK();
        launches++;
        return RMN_OK;
    }
    template <bool INJ>
    void launch(int64_t T, const double* tape, const rmn_trace_t& t0, cudaStream_t stream) {
#ifdef RMN_WITH_TPC
        {
            const char* e = getenv("RMN_CP_KERNEL");
            if (!(e && !strcmp(e, "lanes")) && cp::tpc_smem_bytes(P.M) <= 160 * 1024) {
                cp::tpc_launch(INJ, P, s->model->d_cpdata, st, s->K, T, step0, s->seed, s->chain_offset, tape, t0, stream);
                return;
            }
        }
#endif
        if (use_smem)
            changepoint_kernel<INJ, true, true><<<grid(), 128, smem_bytes, stream>>>(
                P, s->model->d_cpdata, st, s->K, T, step0, s->seed, s->chain_offset, tape, t0);
        else
            changepoint_kernel<INJ, false, true><<<grid(), 128, 0, stream>>>(
                P, s->model->d_cpdata, st, s->K, T, step0, s->seed, s->chain_offset, tape, t0);
    }
    int run(int64_t T, const rmn_inject_t* inj, const rmn_trace_t* tr, cudaStream_t stream) override {
        rmn_trace_t t0{};
        if (tr) t0 = *tr;
        if (t0.thin <= 0) t0.thin = 1;
        if (inj) {
            RMN_REQUIRE(inj->d_tape, "injected changepoint run needs d_tape");
            launch<true>(T, inj->d_tape, t0, stream);
        } else {
            launch<false>(T, nullptr, t0, stream);
        }
        RMN_KERNEL_CHECK();
        launches++;
        // samples = #{ s in [step0, step0+T) : s % EVERY == 0 }
        const int64_t E = RMN_CP_DIAG_EVERY;
        diag_samples += (step0 + T + E - 1) / E - (step0 + E - 1) / E;
        step0 += T; diag_steps += T;
        return RMN_OK;
    }
    int diag_dim() const override { return RMN_CP_NDIAG; }
    int reset_diag(cudaStream_t stream) override {
        RMN_CUDA(cudaMemsetAsync(st.S1, 0, (size_t)RMN_CP_NDIAG * s->K * 8, stream));
        RMN_CUDA(cudaMemsetAsync(st.S2, 0, (size_t)RMN_CP_NDIAG * s->K * 8, stream));
        RMN_CUDA(cudaMemsetAsync(st.dacc, 0, (size_t)s->K * 8, stream));
        RMN_CUDA(cudaMemsetAsync(st.dovf, 0, (size_t)s->K * 8, stream));
        diag_steps = 0; diag_samples = 0;
        return RMN_OK;
    }
    int reduce_diag(double* d_block, cudaStream_t stream) override {
        launches++;
        return rmn_reduce_diag_block(s->K, RMN_CP_NDIAG, diag_samples, diag_steps, st.S1, st.S2, st.dacc, st.dovf,
                                     d_block, stream);
    }
};

}  // namespace

SamplerImpl* make_changepoint_sampler(rmn_sampler* s) { return new ChangepointSampler(s); }

int cp_pointwise(rmn_model* m, int which, int64_t n, const int32_t* d_k, const double* d_cpx,
                 const double* d_cpv, const double* d_sig, double* d_out, cudaStream_t st) {
    if (n <= 0) return RMN_OK;
    const CPParams P = make_params(m, nullptr);
    cp_eval_kernel<<<(unsigned)((n * LANES + 127) / 128), 128, 0, st>>>(P, m->d_cpdata, which, n, d_k,
                                                                        d_cpx, d_cpv, d_sig, d_out);
    RMN_KERNEL_CHECK();
    return RMN_OK;
}

