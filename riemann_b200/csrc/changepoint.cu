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
#include <stdlib.h>

#ifndef RMN_CP_DEFAULT_GL
#define RMN_CP_DEFAULT_GL 4
#endif


namespace {
using namespace cp;

// ---------------------------------------------------------------------------------------
// Group geometry.  GL lanes cooperate on one chain (32/GL chains per warp); lane l holds the
// EPL = LANES/GL elements e = l + GL*j, j = 0..EPL-1 ("row" j) of cpx, cpv and of the cached
// run boundaries in registers, statically indexed.  Rows beyond the warp's largest extent hold
// the canonical padding (cpx = 0, cpv = 0, boundary = M) and are skipped by a warp-uniform
// guard, so with the typical 3..8 changepoints a GL = 4 warp spends its instructions on 8
// chains and two rows instead of on 2 chains and ten idle lanes.
// ---------------------------------------------------------------------------------------
template <int GL> struct Geo {
    static_assert(GL == 4 || GL == 8 || GL == 16, "GL must be 4, 8 or 16");
    static constexpr int EPL = LANES / GL;
    static constexpr int LOG = (GL == 4) ? 2 : ((GL == 8) ? 3 : 4);
    static constexpr int NS = (RMN_CP_NDIAG + GL - 1) / GL;     // diagnostics slots per lane
    // rows below ALWAYS are processed unconditionally (straight-line code, the rows' dependent chains
    // interleave); only the rows above carry the warp-uniform guard.  12 elements cover k <= 10.
    static constexpr int ALWAYS = (GL == 4) ? 3 : EPL;
};

template <int GL>
__device__ __forceinline__ unsigned gballot(bool pred) {
    const unsigned full = __ballot_sync(0xffffffffu, pred);
    return (full >> (threadIdx.x & 31 & ~(GL - 1))) & ((1u << GL) - 1u);
}

template <int GL>
__device__ __forceinline__ double gsum(double v) {
#pragma unroll
    for (int o = GL / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, GL);
    return v;
}

template <int GL>
__device__ __forceinline__ double gprod(double v) {
#pragma unroll
    for (int o = GL / 2; o > 0; o >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, o, GL);
    return v;
}

// row j is live for this warp-step iff it can hold an element of any chain of the warp:
// elements 0..k+1 (k+1 after a birth), kw = warp maximum of k
#define ROW_ON(j) ((j) < Geo<GL>::ALWAYS || (j) * GL <= kw + 1)

// value of element e-1 for every row (element -1 = `first`)
template <int GL, typename T>
__device__ __forceinline__ void shift_prev(const T (&a)[Geo<GL>::EPL], T (&out)[Geo<GL>::EPL], T first,
                                           int lane, int kw) {
#pragma unroll
    for (int j = 0; j < Geo<GL>::EPL; ++j) {
        out[j] = a[j];
        if (ROW_ON(j)) {
            const T up = __shfl_up_sync(0xffffffffu, a[j], 1, GL);
            T w = first;
            if (j > 0) w = __shfl_sync(0xffffffffu, a[j > 0 ? j - 1 : 0], GL - 1, GL);
            out[j] = (lane == 0) ? w : up;
        }
    }
}

// value of element e+1 for every row (element LANES = `pad`)
template <int GL, typename T>
__device__ __forceinline__ void shift_next(const T (&a)[Geo<GL>::EPL], T (&out)[Geo<GL>::EPL], T pad,
                                           int lane, int kw) {
    constexpr int EPL = Geo<GL>::EPL;
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
        out[j] = a[j];
        if (ROW_ON(j)) {
            const T dn = __shfl_down_sync(0xffffffffu, a[j], 1, GL);
            T w = pad;
            if (j + 1 < EPL) w = __shfl_sync(0xffffffffu, a[j + 1 < EPL ? j + 1 : j], 0, GL);
            out[j] = (lane == GL - 1) ? w : dn;
        }
    }
}

// value of element e + off, off in {-1, 0, +1} chosen per chain (birth: -1 above the insertion point,
// death: +1 from the removed element on, everything else 0); elements -1 and LANES read `pad`.
// Two shuffles per row: the row itself and the neighbouring row the wrap-around lane reads from.
template <int GL, typename T>
__device__ __forceinline__ void shift_by(const T (&a)[Geo<GL>::EPL], T (&out)[Geo<GL>::EPL], T pad, int lane,
                                         int dir, const int (&off)[Geo<GL>::EPL], int kw) {
    constexpr int EPL = Geo<GL>::EPL;
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
        out[j] = a[j];
        if (ROW_ON(j)) {
            const int src = (lane + off[j]) & (GL - 1);
            const T lo = (j > 0) ? a[j > 0 ? j - 1 : 0] : pad;            // row below (read by lane 0 when off = -1)
            const T hi = (j + 1 < EPL) ? a[j + 1 < EPL ? j + 1 : j] : pad; // row above (read by lane GL-1 when off = +1)
            const T m = __shfl_sync(0xffffffffu, a[j], src, GL);
            const T w = __shfl_sync(0xffffffffu, (dir < 0) ? lo : hi, src, GL);
            const bool wrap = (off[j] < 0 && lane == 0) || (off[j] > 0 && lane == GL - 1);
            out[j] = wrap ? w : m;
        }
    }
}

// a[idx] for a per-chain element index idx in [0, LANES)
template <int GL>
__device__ __forceinline__ double elem_at(const double (&a)[Geo<GL>::EPL], int idx, int kw) {
    double r = 0.0;
#pragma unroll
    for (int j = 0; j < Geo<GL>::EPL; ++j) {
        if (ROW_ON(j)) {
            const double t = __shfl_sync(0xffffffffu, a[j], idx & (GL - 1), GL);
            if ((idx >> Geo<GL>::LOG) == j) r = t;
        }
    }
    return r;
}

// #{e < k : cpx[e] < s}   (np.searchsorted(cpx, s), changepoint.py:206)
template <int GL>
__device__ __forceinline__ int count_below(const double (&cx)[Geo<GL>::EPL], int lane, int k, double s, int kw) {
    int n = 0;
#pragma unroll
    for (int j = 0; j < Geo<GL>::EPL; ++j)
        if (ROW_ON(j)) n += __popc(gballot<GL>(lane + GL * j < k && cx[j] < s));
    return n;
}

// ---------------------------------------------------------------------------------------
// Log-posterior of one state held across the GL lanes of a group, in three pieces, so that a move
// which leaves part of the state alone reuses the cached pieces of the current state:
//   cp_terms     per element: residual sum of squares of its run of data (prefix sums) -> ss,
//                height-prior term (changepoint.py:18-19,136) -> vt, gap to the previous changepoint
//                (:142-143) -> product gp.  The sum over changepoints of log(gap) is taken as ONE log of
//                the product of the gaps (a negative gap poisons the product with nan, exactly what a sum
//                containing log(negative) gives numpy; a zero gap gives -inf either way).
//   log4         the step's fp64 logarithms -- log gp, log sigma^2, log u (accept test), log|J| -- batched
//                on lanes 0..3 of the group: ONE `log` call per warp-step.
//   cp_assemble  the scalar arithmetic of changepoint.py:118-125,128-160 and model.py:50-54;
//                log(1/sigma^2) (:148) is -log(sigma^2) (+inf where 1/sigma^2 overflows, as in numpy).
// Every floating-point operation that is shared between the general path and the move-specific fast
// paths of the kernel is written with explicit round-to-nearest intrinsics, so all paths give the same
// bits (no context-dependent FMA contraction).  Differences from numpy's term-by-term sums are O(1e-15)
// relative (tests: 1e-9).
// ---------------------------------------------------------------------------------------
template <int GL>
__device__ __forceinline__ void cp_terms(const CPParams& P, const double* __restrict__ cy,
                                         const double* __restrict__ cyy, int lane, int kw, int kk,
                                         const double (&nx)[Geo<GL>::EPL], const double (&nv)[Geo<GL>::EPL],
                                         const int (&nbu)[Geo<GL>::EPL], bool wss, bool wvt, bool wgp,
                                         double& ss, double& vt, double& gp) {
    // wss / wvt / wgp are WARP-UNIFORM: which of the three sums any chain of the warp needs recomputed
    constexpr int EPL = Geo<GL>::EPL;
    if (wss) {
        int pb[EPL];
        shift_prev<GL, int>(nbu, pb, 0, lane, kw);
        double a = 0.0;
#pragma unroll
        for (int j = 0; j < EPL; ++j) {
            if (ROW_ON(j) && lane + GL * j <= kk) {
                const int bu = nbu[j], bl = pb[j];
                const double n = (double)(bu - bl);
                const double s1 = __dsub_rn(cy[bu], cy[bl]);
                const double s2 = __dsub_rn(cyy[bu], cyy[bl]);
                const double vc = __dsub_rn(nv[j], P.ycenter);
                a = __dadd_rn(a, __fma_rn(__dmul_rn(n, vc), vc, __fma_rn(-2.0 * vc, s1, s2)));   // n vc^2 - 2 vc s1 + s2
            }
        }
        ss = gsum<GL>(a);
    }
    if (wgp) {
        double px[EPL];
        shift_prev<GL, double>(nx, px, P.xmin, lane, kw);
        double g = 1.0;
#pragma unroll
        for (int j = 0; j < EPL; ++j) {
            const int e = lane + GL * j;
            if (ROW_ON(j) && e <= kk) {
                const double gap = __dsub_rn((e < kk) ? nx[j] : P.xmax, px[j]);           // changepoint.py:142-143
                g = __dmul_rn(g, (gap < 0.0) ? NAN : gap);
            }
        }
        gp = gprod<GL>(g);
    }
    if (wvt) {
        double b = 0.0;
#pragma unroll
        for (int j = 0; j < EPL; ++j) {
            if (ROW_ON(j) && lane + GL * j <= kk) {
                double t = __fma_rn(-P.beta, nv[j], P.cv);                                // changepoint.py:18-19,136
                if (!P.alpha_is_one) t = __fma_rn(P.alpha - 1.0, log(nv[j]), t);
                else if (!(nv[j] > 0.0)) t = NAN;                                         // 0 * log(v<=0) is nan in numpy
                b = __dadd_rn(b, t);
            }
        }
        vt = gsum<GL>(b);
    }
}

// log of four per-chain arguments with one call: lane (l & 3) of the group takes argument l & 3
template <int GL>
__device__ __forceinline__ void log4(int lane, double a0, double a1, double a2, double a3, double& l0, double& l1,
                                     double& l2, double& l3) {
    const int sl = lane & 3;
    const double larg = (sl == 0) ? a0 : ((sl == 1) ? a1 : ((sl == 2) ? a2 : a3));
    const double lres = log(larg);
    l0 = __shfl_sync(0xffffffffu, lres, 0, GL);
    l1 = __shfl_sync(0xffffffffu, lres, 1, GL);
    l2 = __shfl_sync(0xffffffffu, lres, 2, GL);
    l3 = __shfl_sync(0xffffffffu, lres, 3, GL);
}

// which: 0 = log posterior (model.py:43-55), 1 = log likelihood, 2 = log prior
__device__ __forceinline__ double cp_assemble(const CPParams& P, int kk, double ss, double vt, double lg, double nsig,
                                              double log_s2, int which) {
    const double s2v = __dmul_rn(nsig, nsig);
    const int ks = kk + 1;                                     // number of steps
    double logl = -0.5 * __dadd_rn(__fma_rn((double)P.M, log_s2, __ddiv_rn(ss, s2v)), P.Mlog2pi);   // :118-120
    if (isnan(logl)) logl = -INFINITY;                         // :124-125
    double lsig = (s2v < 5.562684646268003e-309) ? INFINITY : -log_s2;      // log(1/sigma^2), :148
    if (nsig < 0.0) lsig = NAN;                                // :146-147
    const double lps = __fma_rn(-(double)ks, P.logL, __dadd_rn(P.tab2[ks], lg));
    double logp = __dadd_rn(__dadd_rn(__dadd_rn(P.tab1[ks], vt), lps), lsig);              // :156
    if (isnan(logp)) logp = -INFINITY;                         // :158-159
    if (which == 1) return logl;
    if (which == 2) return logp;
    return combine_logpost(logp, logl);
}

// full evaluation (pointwise entry, state caches at kernel entry)
template <int GL>
__device__ __forceinline__ double cp_logpost_rows(const CPParams& P, const double* __restrict__ cy,
                                                  const double* __restrict__ cyy, int lane, int kw, int kk,
                                                  const double (&nx)[Geo<GL>::EPL], const double (&nv)[Geo<GL>::EPL],
                                                  const int (&nbu)[Geo<GL>::EPL], double nsig, int which,
                                                  double& ss, double& vt, double& lg, double& log_s2) {
    double gp, d2, d3;
    cp_terms<GL>(P, cy, cyy, lane, kw, kk, nx, nv, nbu, true, true, true, ss, vt, gp);
    log4<GL>(lane, gp, __dmul_rn(nsig, nsig), 1.0, 1.0, lg, log_s2, d2, d3);
    return cp_assemble(P, kk, ss, vt, lg, nsig, log_s2, which);
}

#ifndef RMN_CP_MINBLOCKS
#define RMN_CP_MINBLOCKS 4   /* 128 regs, 16 warps/SM, no spills: best of 4, 5, 6 with the move-specific guards (gpurun r2e) */
#endif
#ifndef RMN_CP_FASTPATHS
#define RMN_CP_FASTPATHS 1   /* move-specific paths for warps whose chains all drew the same move type */
#endif

// ---------------------------------------------------------------------------------------
// The T-step kernel.
//
// Move schedule (Philox mode).  Which of the four moves a chain attempts at step t is drawn
// independently of its state (three fresh uniforms, test_changepoint.py:48-54).  shared_mv = 1: the
// 32/GL chains of a warp -- global chain ids aligned to 32/GL, so the grouping does not depend on how the
// chains are sharded -- take the three selection words from the group's first chain.  Each chain still
// sees an i.i.d. sequence of move types with the reference's probabilities, independent of its state,
// so each is an exact replica of the reference sampler; given the schedule the chains of a group are
// independent, and in stationarity their ergodic averages are uncorrelated (every schedule's kernel
// preserves the target).  The warp then executes ONE move type per step: a real branch to a
// move-specific path instead of the union of all four by selection.  shared_mv = 0: per-chain move
// types (the general path every step).  Injected mode replays per-chain move types from the tape and
// takes a fast path whenever the warp happens to be uniform (always for K = 1), so the fast paths are
// covered by the reference-stream parity tests.
//
// Fast paths (cached per chain: ss, vt, lg = log gap product, ls2 = log sigma^2 of the CURRENT state):
//   sigma move  test_changepoint.py:54-56   likelihood and prior from the cached sums, no search
//   cpv move    :51-53                      residual / height-prior sums; gaps and run boundaries cached
//   cpx move    :48-50                      searches, residual sums, gap product; height-prior sum cached
//   birth/death :57-71                      the general path
// ---------------------------------------------------------------------------------------
// GUARD = false: the instantiation for per-chain move schedules in Philox mode -- the warp executes the union of the
// moves every step, so the votes and the guards that a uniform warp would use are compiled out (same arithmetic, same bits).
template <bool INJ, int DATA, int GL, bool GUARD>
__device__ __forceinline__ void cp_block(const CPParams& P, const double* __restrict__ xs, CPState st,
                                         int64_t K, int64_t t_begin, int64_t t_end, int64_t step0, uint64_t seed,
                                         int64_t chain_offset, const double* __restrict__ tape, const rmn_trace_t& tr,
                                         int shared_mv, int64_t blk) {
    // iterations [t_begin, t_end) of the launch for the 128 / GL chains of block `blk`; the chain state is read from and
    // written back to global memory (L2-coherent loads: in the time-sliced kernel another SM wrote it)
    constexpr int EPL = Geo<GL>::EPL;
    constexpr int NS = Geo<GL>::NS;
    constexpr int CPW = 32 / GL;                  // chains per warp = chains per schedule group
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int LOGP2 = (DATA == 2) ? 6 : -1;
    const double* cy = xs + P.XP;
    const double* cyy = cy + P.M + 1;

    const int lane = threadIdx.x & (GL - 1);
    // schedule groups are aligned to GLOBAL chain ids: `head` leading groups of the grid are dead
    const int head = shared_mv ? (int)(chain_offset & (CPW - 1)) : 0;
    const int64_t c_raw = (blk * (int64_t)blockDim.x + threadIdx.x) / GL - head;
    const bool live = c_raw >= 0 && c_raw < K;
    const int64_t c = c_raw < 0 ? 0 : (c_raw < K ? c_raw : K - 1);   // dead groups shadow a live chain, never store

    int k = __ldcg(st.k + c);
    double cx[EPL], cv[EPL];
    int bu[EPL];                                  // cached run boundaries of the state
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
        const int e = lane + GL * j;
        cx[j] = __ldcg(st.cpx + c * LANES + e);
        cv[j] = __ldcg(st.cpv + c * LANES + e);
        bu[j] = (e < k) ? upper_bound<LOGP2>(xs, P.P2, cx[j]) : P.M;
    }
    double sig = __ldcg(st.sig + c);
    double lp = __ldcg(st.lp + c);
    double ss_c, vt_c, lg_c, ls2_c;               // pieces of lp (see cp_terms / cp_assemble)
    cp_logpost_rows<GL>(P, cy, cyy, lane, LANES, k, cx, cv, bu, sig, 0, ss_c, vt_c, lg_c, ls2_c);
    int nacc = 0, novf = 0;
    double s1[NS], s2[NS];                        // lane l, slot s accumulates functional l + GL*s
#pragma unroll
    for (int s = 0; s < NS; ++s) { s1[s] = 0.0; s2[s] = 0.0; }
    const RngKey rk(seed, (uint64_t)(chain_offset + c_raw));
    const TraceSel ts{tr.first, tr.thin > 0 ? tr.thin : 1};
    const bool tracing = tr.d_k || tr.d_cpx || tr.d_cpv || tr.d_sig || tr.d_logpost;
    const bool tracing_prop = tr.d_prop_logpost || tr.d_accepted || tr.d_logqratio || tr.d_prop_k ||
                              tr.d_prop_sig || tr.d_prop_cpx || tr.d_prop_cpv;

    // per-step proposal record (parity harness): what Proposal.propose returned, what the sampler decided
    auto trace_prop = [&](int64_t t, double lpn, bool acc, double lqr, int kk, double nsig,
                          const double (&nx)[EPL], const double (&nv)[EPL]) {
        if (!(live && tracing_prop)) return;
        if (lane == 0) {
            if (tr.d_prop_logpost) tr.d_prop_logpost[t * K + c] = lpn;
            if (tr.d_accepted) tr.d_accepted[t * K + c] = acc ? 1 : 0;
            if (tr.d_logqratio) tr.d_logqratio[t * K + c] = lqr;
            if (tr.d_prop_k) tr.d_prop_k[t * K + c] = kk;
            if (tr.d_prop_sig) tr.d_prop_sig[t * K + c] = nsig;
        }
#pragma unroll
        for (int j = 0; j < EPL; ++j) {
            const int e = lane + GL * j;
            if (tr.d_prop_cpx) tr.d_prop_cpx[(t * K + c) * LANES + e] = nx[j];
            if (tr.d_prop_cpv) tr.d_prop_cpv[(t * K + c) * LANES + e] = nv[j];
        }
    };

    // the Philox block of step t+1 is issued during step t (it depends on nothing but the counter), so its
    // integer multiplies fill the fp64 / LDS latency of the evaluation
    uint4 rA = make_uint4(0, 0, 0, 0);
    if (!INJ) rA = rk.block((uint64_t)(step0 + t_begin), (uint32_t)lane);

    for (int64_t t = t_begin; t < t_end; ++t) {
        const uint64_t step = (uint64_t)(step0 + t);
        const int kw = __reduce_max_sync(FULL, k);
        double snew, du, uacc;
        int nrand, mv;
        bool birth;
        const double* row = nullptr;
        const uint4 r = rA;
        if (INJ) {
            row = tape + (t * K + c) * RMN_CP_NSLOT;
            const double u1 = row[RMN_CP_SLOT_SEL1], u2 = row[RMN_CP_SLOT_SEL2], u3 = row[RMN_CP_SLOT_SEL3];
            const double ubd = row[RMN_CP_SLOT_BD];
            snew = row[RMN_CP_SLOT_S]; du = row[RMN_CP_SLOT_DU];
            nrand = (int)row[RMN_CP_SLOT_N]; uacc = row[RMN_CP_SLOT_ACC];
            // which block moves (fresh uniform per elif, test_changepoint.py:48-54); birth/death :59
            mv = (u1 < P.p1) ? 0 : ((u2 < P.p2) ? 1 : ((u3 < P.p3) ? 2 : 3));
            birth = (k == 0) || (ubd > 0.5);
        } else {
            // ONE Philox block per lane and step: words x,y -> Box-Muller pair = the normals of this lane's
            // rows 0 and 1; the spare words z,w of lanes 0..3 carry the chain-level uniforms (the three
            // move-selection words come from the schedule group's first chain when shared_mv).
            const int mvw = shared_mv ? 32 : GL;
            const uint32_t z0 = __shfl_sync(FULL, r.z, 0, mvw), w0 = __shfl_sync(FULL, r.w, 0, mvw);
            const uint32_t z1 = __shfl_sync(FULL, r.z, 1, mvw), w1 = __shfl_sync(FULL, r.w, 1, GL);
            const uint32_t z2 = __shfl_sync(FULL, r.z, 2, GL), w2 = __shfl_sync(FULL, r.w, 2, GL);
            const uint32_t z3 = __shfl_sync(FULL, r.z, 3, GL), w3 = __shfl_sync(FULL, r.w, 3, GL);
            mv = (z0 < P.t1) ? 0 : ((w0 < P.t2) ? 1 : ((z1 < P.t3) ? 2 : 3));   // same tests on raw words
            birth = (k == 0) || (w1 >= 0x80000000u);                            // u > 0.5
            snew = P.xmin + (P.xmax - P.xmin) * u01_fast(z2);
            du = -0.1 + 0.2 * u01_fast(w2);
            nrand = (int)(u01_fast(z3) * (double)k);
            uacc = u01_fast(w3);
        }
        nrand = max(0, min(nrand, k - 1));
        // What the warp as a whole needs this step (warp-uniform).  With a shared move schedule all chains of the
        // warp drew the same move, so most of the step's blocks are skipped by real branches; with per-chain move types
        // the guards are almost always taken and the warp executes the union, selecting per chain.
        const bool any0 = __any_sync(FULL, mv == 0), any1 = __any_sync(FULL, mv == 1);
        const bool any3 = __any_sync(FULL, mv == 3);
        const bool skip = RMN_CP_FASTPATHS && GUARD && (INJ || shared_mv);   // per-chain Philox mode keeps ONE instruction stream
        const bool need_xi = !skip || __any_sync(FULL, mv != 3);

        // normals of the block moves: rows 0,1 from this step's block, rows 2,3 (only when some chain of the
        // warp has an element there and a location / height move is being made) from a second block
        double xi[EPL];
#pragma unroll
        for (int j = 0; j < EPL; ++j) xi[j] = 0.0;
        if (INJ) {
#pragma unroll
            for (int j = 0; j < EPL; ++j) xi[j] = row[RMN_CP_SLOT_XI + lane + GL * j];
        } else if (need_xi) {
            float n0, n1;
            box_muller(r.x, r.y, n0, n1);
            xi[0] = (double)n0;
            if (EPL > 1) xi[1 % EPL] = (double)n1;
            if (EPL > 2 && kw >= 2 * GL && (!skip || any0 || any1)) {
                const uint4 q = rk.block(step, (uint32_t)(GL + lane));
                box_muller(q.x, q.y, n0, n1);
                xi[2 % EPL] = (double)n0; xi[3 % EPL] = (double)n1;
            }
        }

        // ---- build the proposal by selection (the chains of a warp never diverge on mv)
        int kk = k, nbu[EPL];
        double nx[EPL], nv[EPL], nsig = sig, jarg = 1.0;
        bool ovf = false;
        const double sxk = P.sx[k];
#pragma unroll
        for (int j = 0; j < EPL; ++j) {
            const int e = lane + GL * j;
            nx[j] = cx[j]; nv[j] = cv[j]; nbu[j] = bu[j];
            if (ROW_ON(j)) {
                if (mv == 0 && e < k) nx[j] = __dadd_rn(cx[j], __dmul_rn(sxk, xi[j]));   // randomwalk.py:26, scale = 1
                if (mv == 1 && e <= k) nv[j] = __dadd_rn(cv[j], __dmul_rn(P.sv, xi[j]));
            }
        }
        const double xi0 = __shfl_sync(FULL, xi[0], 0, GL);
        if (mv == 2) nsig = __dadd_rn(sig, __dmul_rn(P.ss, xi0));
        int nb = 0;
        if (any3) {                                                     // warp-uniform
            nb = count_below<GL>(cx, lane, k, snew, kw);                // searchsorted(cpx, s), :206
            const double hb = elem_at<GL>(cv, nb, kw);
            const double h1d = elem_at<GL>(cv, nrand, kw);
            const double h2d = elem_at<GL>(cv, nrand + 1, kw);
            // birth (changepoint.py:57,61,72-74) and death (:67-68,76-78) share one instruction
            // stream: the operands of each division / square root are selected per chain
            const double q1 = birth ? du / P.sqrtM : h2d / h1d;
            const double ub = 0.5 + q1;                                 // birth: u, :61
            const double q2 = birth ? (1.0 - ub) / ub : 1.0 / (1.0 + q1);       // birth: f^2; death: u, :68
            const double u = birth ? ub : q2;
            const double rr = sqrt(birth ? q2 : h1d * h2d);             // birth: f, :57; death: h, :67
            jarg = fabs((birth ? hb : rr) / (u * (1.0 - u)));           // |J| resp. 1/|J^-1|
            const double hbf = hb / rr;
            ovf = (mv == 3) && birth && (k + 1 > LANES - 1);
            const bool td = (mv == 3) && !ovf;
            if (!(mv == 3)) jarg = 1.0;
            // insert / delete = every element reads its neighbour below (birth, above the insertion
            // point) or above (death, from the removed element on); other chains read themselves
            const int dir = birth ? -1 : 1;
            const int pivot = birth ? nb : nrand - 1;                   // elements <= pivot stay
            int off[EPL];
#pragma unroll
            for (int j = 0; j < EPL; ++j) off[j] = (td && lane + GL * j > pivot) ? dir : 0;
            double sx_[EPL], sv_[EPL];
            int sb_[EPL];
            shift_by<GL, double>(cx, sx_, 0.0, lane, dir, off, kw);
            shift_by<GL, double>(cv, sv_, 0.0, lane, dir, off, kw);
            shift_by<GL, int>(bu, sb_, P.M, lane, dir, off, kw);
            if (td) {
                kk = birth ? k + 1 : k - 1;
#pragma unroll
                for (int j = 0; j < EPL; ++j) {
                    const int e = lane + GL * j;
                    nx[j] = sx_[j]; nv[j] = sv_[j]; nbu[j] = sb_[j];
                    if (birth) {
                        if (e == nb) { nx[j] = snew; nv[j] = hbf; nbu[j] = bu[j]; }   // boundary searched below
                        if (e == nb + 1) nv[j] = hb * rr;
                    } else if (e == nrand) {
                        nv[j] = rr;
                    }
                }
            }
            // elements beyond the new extent hold zeros (canonical padding)
#pragma unroll
            for (int j = 0; j < EPL; ++j) {
                const int e = lane + GL * j;
                if (e >= kk) { nx[j] = 0.0; nbu[j] = P.M; }
                if (e > kk) nv[j] = 0.0;
            }
        }
        // run boundaries only move when a location moves: a cpx block move (every element) or a
        // birth (the new element).
        const bool moved_x = (mv == 0) || (mv == 3 && birth && !ovf);
        if (!skip || any0 || any3) {
            constexpr int NA = Geo<GL>::ALWAYS;
            double qa[NA];
            int sa[NA];
#pragma unroll
            for (int j = 0; j < NA; ++j) qa[j] = nx[j];
            upper_bound_rows<LOGP2, NA>(xs, P.P2, qa, sa);
#pragma unroll
            for (int j = 0; j < EPL; ++j) {
                const int e = lane + GL * j;
                int sb = 0;
                if (j < NA) sb = sa[j < NA ? j : 0];
                else if (ROW_ON(j)) sb = upper_bound<LOGP2>(xs, P.P2, nx[j]);
                if (moved_x && e < kk && (mv == 0 || e == nb)) nbu[j] = sb;
            }
        }
        if (!INJ) rA = rk.block(step + 1, (uint32_t)lane);
        // ---- log-posterior of the proposal from the pieces that changed (a sigma move: none of the three sums;
        //      a height move: ss, vt; a location move: ss, gp; birth / death: all), every fp64 log of the step in
        //      one call.  A chain whose move leaves a piece alone recomputes the cached value bit for bit.
        double nss = ss_c, nvt = vt_c, gp = 1.0, nlg, nls2, logu, ljac;
        const bool wss = !skip || any0 || any1 || any3, wvt = !skip || any1 || any3, wgp = !skip || any0 || any3;
        cp_terms<GL>(P, cy, cyy, lane, kw, kk, nx, nv, nbu, wss, wvt, wgp, nss, nvt, gp);
        log4<GL>(lane, gp, __dmul_rn(nsig, nsig), uacc, jarg, nlg, nls2, logu, ljac);
        if (!wgp) nlg = lg_c;
        const double lpn = cp_assemble(P, kk, nss, nvt, nlg, nsig, nls2, 0);
        const double lqr = (mv == 3) ? (birth ? ljac : -ljac) : 0.0;

        // sampler.py:83-84 with Python's min(0, nan) == 0
        const double delta = lpn - lp - lqr;
        const double mh = (delta < 0.0) ? delta : 0.0;
        const bool acc = !ovf && (logu < mh);
        trace_prop(t, lpn, acc, lqr, kk, nsig, nx, nv);
        if (acc) {
            k = kk; sig = nsig; lp = lpn;
            ss_c = nss; vt_c = nvt; lg_c = nlg; ls2_c = nls2;
#pragma unroll
            for (int j = 0; j < EPL; ++j) { cx[j] = nx[j]; cv[j] = nv[j]; bu[j] = nbu[j]; }
        }
        nacc += acc ? 1 : 0;
        novf += ovf ? 1 : 0;

        if ((step % RMN_CP_DIAG_EVERY) == 0) {              // thinned accumulation (warp-uniform)
            const int kd = __reduce_max_sync(FULL, k);      // (an accepted birth may have raised the extent)
            double f[NS];
#pragma unroll
            for (int s = 0; s < NS; ++s) f[s] = 0.0;
            if (lane == 0) f[0] = sig;
            if (lane == 1 % GL) f[1 / GL] = (double)k;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int cnt = count_below<GL>(cx, lane, k, P.xq[q], kd);
                const double yq = elem_at<GL>(cv, cnt, kd);
                if (lane == (2 + q) % GL) f[(2 + q) / GL] = yq;
            }
#pragma unroll
            for (int s = 0; s < NS; ++s)
                if (lane + GL * s < RMN_CP_NDIAG) { s1[s] += f[s]; s2[s] += f[s] * f[s]; }
        }

        if (live && tracing) {
            const long long rrec = ts.slot(t + 1);
            if (rrec >= 0) {
#pragma unroll
                for (int j = 0; j < EPL; ++j) {
                    const int e = lane + GL * j;
                    if (tr.d_cpx) tr.d_cpx[(rrec * K + c) * LANES + e] = cx[j];
                    if (tr.d_cpv) tr.d_cpv[(rrec * K + c) * LANES + e] = cv[j];
                }
                if (lane == 0) {
                    if (tr.d_k) tr.d_k[rrec * K + c] = k;
                    if (tr.d_sig) tr.d_sig[rrec * K + c] = sig;
                    if (tr.d_logpost) tr.d_logpost[rrec * K + c] = lp;
                }
            }
        }
    }

    if (live) {
#pragma unroll
        for (int j = 0; j < EPL; ++j) {
            const int e = lane + GL * j;
            st.cpx[c * LANES + e] = cx[j];
            st.cpv[c * LANES + e] = cv[j];
        }
        if (lane == 0) {
            st.k[c] = k;
            st.sig[c] = sig;
            st.lp[c] = lp;
            st.dacc[c] = __ldcg(st.dacc + c) + nacc;
            st.dovf[c] = __ldcg(st.dovf + c) + novf;
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int i = lane + GL * s;
            if (i < RMN_CP_NDIAG) {
                st.S1[(int64_t)i * K + c] = __ldcg(st.S1 + (int64_t)i * K + c) + s1[s];
                st.S2[(int64_t)i * K + c] = __ldcg(st.S2 + (int64_t)i * K + c) + s2[s];
            }
        }
    }
}

// DATA: 0 = data tables read from global memory (too large for shared memory), 1 = staged in shared
// memory, 2 = staged and 64 <= M < 128 (P2 = 64: fully unrolled 7-level search; the bench shape)
template <bool INJ, int DATA, int GL, bool GUARD = true>
__global__ void __launch_bounds__(128, (GL == 16) ? 8 : RMN_CP_MINBLOCKS)
changepoint_kernel(const __grid_constant__ CPParams P, const double* __restrict__ gdata, CPState st,
                   int64_t K, int64_t T, int64_t step0, uint64_t seed, int64_t chain_offset,
                   const double* __restrict__ tape, rmn_trace_t tr, int shared_mv) {
    extern __shared__ double smem[];
    const double* xs = gdata;
    if (DATA != 0) {
        const int n = P.XP + 2 * P.M + 2;
        for (int i = threadIdx.x; i < n; i += blockDim.x) smem[i] = gdata[i];
        __syncthreads();
        xs = smem;
    }
    cp_block<INJ, DATA, GL, GUARD>(P, xs, st, K, 0, T, step0, seed, chain_offset, tape, tr, shared_mv, blockIdx.x);
}

// Time-sliced form of the same run (plain Philox runs without traces).  A launch of B blocks that do not fit the SMs'
// resident-block slots S in a whole number of waves ends with a partly filled wave (2,048 blocks on 592 slots = 3.46 waves
// at the bench shape: 4.5 % of the launch).  Here S persistent blocks take the items (time slice j, block b), index
// j B + b, from an atomic counter in increasing order: every persistent block gets the same amount of work to within one
// item, the state of a slice's chains travels through global memory, and a flag per chain block orders the slices of the
// same chains (the predecessor is B >> S items earlier, so the wait is almost never taken).  Philox is keyed by
// (step, chain), so the chains are bit-identical to the one-slice launch.
template <int DATA, int GL, bool GUARD = true>
__global__ void __launch_bounds__(128, (GL == 16) ? 8 : RMN_CP_MINBLOCKS)
changepoint_sliced_kernel(const __grid_constant__ CPParams P, const double* __restrict__ gdata, CPState st,
                          int64_t K, int64_t T, int64_t step0, uint64_t seed, int64_t chain_offset, int shared_mv,
                          int nblk, int slice, int nslice, int* __restrict__ done) {
    extern __shared__ double smem[];
    const double* xs = gdata;
    if (DATA != 0) {
        const int n = P.XP + 2 * P.M + 2;
        for (int i = threadIdx.x; i < n; i += blockDim.x) smem[i] = gdata[i];
        __syncthreads();
        xs = smem;
    }
    rmn_trace_t tr{};
    __shared__ int s_item;
    const int items = nblk * nslice;
    for (;;) {
        // items are handed out in increasing order to blocks that are RUNNING: the predecessor of an item was taken
        // earlier, so its owner is resident whatever else shares the device -- the wait below cannot deadlock
        if (threadIdx.x == 0) {
            const int it = atomicAdd(done + nblk, 1);
            const int j = it / nblk, b = it % nblk;
            if (it < items && j > 0) {
                while (*reinterpret_cast<volatile int*>(done + b) < j) __nanosleep(64);
                __threadfence();
            }
            s_item = it;
        }
        __syncthreads();
        const int it = s_item;
        __syncthreads();
        if (it >= items) break;
        const int j = it / nblk, b = it % nblk;
        const int64_t t0 = (int64_t)j * slice, t1 = (t0 + slice < T) ? t0 + slice : T;
        cp_block<false, DATA, GL, GUARD>(P, xs, st, K, t0, t1, step0, seed, chain_offset, nullptr, tr, shared_mv, b);
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) *reinterpret_cast<volatile int*>(done + b) = j + 1;
    }
}

// evaluate n states given in the canonical layout (pointwise parity entry + set_state)
__global__ void __launch_bounds__(128)
cp_eval_kernel(const __grid_constant__ CPParams P, const double* __restrict__ gdata, int which,
               int64_t n, const int32_t* __restrict__ kin, const double* __restrict__ cpx,
               const double* __restrict__ cpv, const double* __restrict__ sig,
               double* __restrict__ out) {
    constexpr int GL = 16;
    const int lane = threadIdx.x & (GL - 1);
    const int64_t c_raw = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / GL;
    const bool live = c_raw < n;
    const int64_t c = live ? c_raw : n - 1;
    const double* xs = gdata;
    const double* cy = xs + P.XP;
    const double* cyy = cy + P.M + 1;
    const int k = kin[c];
    const int kw = LANES;                                        // every row on
    double x1[1] = {cpx[c * LANES + lane]}, v1[1] = {cpv[c * LANES + lane]};
    int b1[1] = {(lane < k) ? upper_bound<-1>(xs, P.P2, x1[0]) : P.M};
    if (lane >= k) x1[0] = 0.0;
    if (lane > k) v1[0] = 0.0;
    double ss, vt, lg, ls2;
    const double v = cp_logpost_rows<GL>(P, cy, cyy, lane, kw, k, x1, v1, b1, sig[c], which, ss, vt, lg, ls2);
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
    P.XP = 2 * P.P2;
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
    int gl = RMN_CP_DEFAULT_GL;     // lanes per chain (4, 8 or 16); RMN_CP_GL overrides (A/B measurements)
    int shared_mv = 1;              // Philox mode: move types shared by aligned groups of 32/gl chains (see the kernel);
                                    // rmn_sampler_set_move_schedule / RMN_CP_SCHEDULE=chain select per-chain move types
    explicit ChangepointSampler(rmn_sampler* s_) : s(s_) {
        P = make_params(s->model, s->prop);
        if (const char* e = getenv("RMN_CP_GL")) {
            const int v = atoi(e);
            if (v == 4 || v == 8 || v == 16) gl = v;
        }
        if (const char* e = getenv("RMN_CP_SCHEDULE")) shared_mv = (e[0] == 'c' || e[0] == '0') ? 0 : 1;
        if (const char* e = getenv("RMN_CP_SLICED")) sliced = (e[0] == '0') ? 0 : 1;
        if (const char* e = getenv("RMN_CP_GUARD")) force_guard = (e[0] == '1') ? 1 : 0;
        smem_bytes = (size_t)(P.XP + 2 * P.M + 2) * 8;
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
            RMN_CUDA(set_smem_attr<4>());
            RMN_CUDA(set_smem_attr<8>());
            RMN_CUDA(set_smem_attr<16>());
        }
        return RMN_OK;
    }
    template <int GL> cudaError_t set_smem_attr() {
        cudaError_t e = rmn_raise_dyn_smem((const void*)changepoint_kernel<false, 1, GL>, smem_bytes);
        if (e != cudaSuccess) return e;
        e = rmn_raise_dyn_smem((const void*)changepoint_kernel<true, 1, GL>, smem_bytes);
        if (e != cudaSuccess || GL != 4) return e;
        e = rmn_raise_dyn_smem((const void*)changepoint_kernel<false, 1, 4, false>, smem_bytes);
        if (e != cudaSuccess) return e;
        e = rmn_raise_dyn_smem((const void*)changepoint_sliced_kernel<1, 4, false>, smem_bytes);
        if (e != cudaSuccess) return e;
        return rmn_raise_dyn_smem((const void*)changepoint_sliced_kernel<1, 4>, smem_bytes);
    }
    unsigned grid(int gl = LANES) const { return (unsigned)((s->K * gl + 127) / 128); }
    // grid of the T-step kernel: schedule groups are aligned to global chain ids, so up to 32/gl - 1 leading
    // groups of the grid are dead when the shard does not start on a group boundary
    unsigned run_grid(int gl, int shared) const {
        const int64_t head = shared ? (s->chain_offset & (int64_t)(32 / gl - 1)) : 0;
        return (unsigned)(((s->K + head) * gl + 127) / 128);
    }
    int set_move_schedule(int mode) override {
        RMN_REQUIRE(mode == 0 || mode == 1, "move schedule: 0 = per chain, 1 = shared by aligned groups of chains");
        shared_mv = mode;
        return RMN_OK;
    }

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
        switch (gl) {
            case 16: launch_gl<INJ, 16>(T, tape, t0, stream); break;
            case 8: launch_gl<INJ, 8>(T, tape, t0, stream); break;
            default: launch_gl<INJ, 4>(T, tape, t0, stream); break;
        }
    }
    // time-sliced launch (changepoint_sliced_kernel): plain Philox runs whose blocks do not fill a whole number of waves
    int* d_done = nullptr;
    int done_cap = 0;
    int sliced = 1;                 // RMN_CP_SLICED=0 switches it off (A/B)
    int force_guard = 0;            // RMN_CP_GUARD=1: per-chain schedules through the guarded instantiation (A/B)
    ~ChangepointSampler() override { if (d_done) cudaFree(d_done); }
    bool last_sliced = false;       // which kernel the last plain launch used (the timer label is corrected after the launch)
    template <int DATA, int GL, bool GUARD>
    bool launch_sliced(int64_t T, unsigned nblk, int sh, size_t smem, cudaStream_t stream) {
        constexpr int SLICE_MIN = 100;          // iterations: the state round trip and the entry evaluation cost about one
        last_sliced = false;
        if (!sliced || T < 2 * SLICE_MIN) return false;
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, changepoint_sliced_kernel<DATA, GL, GUARD>, 128, smem) != cudaSuccess || per_sm < 1)
            return false;
        const unsigned slots = (unsigned)(per_sm * sms);
        if (nblk <= slots || nblk % slots == 0) return false;       // one wave, or whole waves: nothing to gain
        // slices: the count with the least modelled time = rounds of the striped walk per item, times the slice overhead
        // (state round trip + entry evaluation, about 2.5 iterations' worth)
        int nslice = 2;
        double best = 0.0;
        for (int v = 2; v <= 16 && (int64_t)v * SLICE_MIN <= T; ++v) {
            const double items = (double)nblk * v;
            const double cost = ceil(items / slots) / (items / slots) * (1.0 + 2.5 * v / (double)T);
            if (best == 0.0 || cost < best - 1e-9) { best = cost; nslice = v; }
        }
        if (best >= ceil((double)nblk / slots) / ((double)nblk / slots)) return false;     // the plain launch is as good
        const int slice = (int)((T + nslice - 1) / nslice);
        nslice = (int)((T + slice - 1) / slice);
        if ((int)nblk > done_cap) {
            if (d_done) cudaFree(d_done);
            d_done = nullptr; done_cap = 0;
            if (cudaMalloc(&d_done, ((size_t)nblk + 1) * sizeof(int)) != cudaSuccess) return false;     // flags + item counter
            done_cap = (int)nblk;
        }
        if (cudaMemsetAsync(d_done, 0, ((size_t)nblk + 1) * sizeof(int), stream) != cudaSuccess) return false;
        changepoint_sliced_kernel<DATA, GL, GUARD><<<slots, 128, smem, stream>>>(
            P, s->model->d_cpdata, st, s->K, T, step0, s->seed, s->chain_offset, sh, (int)nblk, slice, nslice, d_done);
        last_sliced = true;
        return true;
    }
    template <bool INJ, int GL>
    void launch_gl(int64_t T, const double* tape, const rmn_trace_t& t0, cudaStream_t stream) {
        if constexpr (!INJ && GL == 4) {
            // per-chain move schedules: the instantiation without the warp-uniform guards (RMN_CP_GUARD=1 keeps them: A/B
            // and the bit-identity test)
            if (!shared_mv && !force_guard) { launch_gl_g<INJ, GL, false>(T, tape, t0, stream); return; }
        }
        launch_gl_g<INJ, GL, true>(T, tape, t0, stream);
    }
    template <bool INJ, int GL, bool GUARD>
    void launch_gl_g(int64_t T, const double* tape, const rmn_trace_t& t0, cudaStream_t stream) {
        const int sh = INJ ? 0 : shared_mv;
        const unsigned g = run_grid(GL, sh);
        last_sliced = false;
        const bool traced = t0.d_k || t0.d_cpx || t0.d_cpv || t0.d_sig || t0.d_logpost || t0.d_prop_logpost || t0.d_accepted ||
                            t0.d_logqratio || t0.d_prop_k || t0.d_prop_sig || t0.d_prop_cpx || t0.d_prop_cpv;
        if constexpr (!INJ && GL == 4) {
            if (!traced && use_smem && P.P2 == 64) { if (launch_sliced<2, GL, GUARD>(T, g, sh, smem_bytes, stream)) return; }
            else if (!traced && use_smem) { if (launch_sliced<1, GL, GUARD>(T, g, sh, smem_bytes, stream)) return; }
        }
        if (use_smem && P.P2 == 64)
            changepoint_kernel<INJ, 2, GL, GUARD><<<g, 128, smem_bytes, stream>>>(
                P, s->model->d_cpdata, st, s->K, T, step0, s->seed, s->chain_offset, tape, t0, sh);
        else if (use_smem)
            changepoint_kernel<INJ, 1, GL, GUARD><<<g, 128, smem_bytes, stream>>>(
                P, s->model->d_cpdata, st, s->K, T, step0, s->seed, s->chain_offset, tape, t0, sh);
        else
            changepoint_kernel<INJ, 0, GL, GUARD><<<g, 128, 0, stream>>>(
                P, s->model->d_cpdata, st, s->K, T, step0, s->seed, s->chain_offset, tape, t0, sh);
    }
    int run(int64_t T, const rmn_inject_t* inj, const rmn_trace_t* tr, cudaStream_t stream) override {
        rmn_trace_t t0{};
        if (tr) t0 = *tr;
        if (t0.thin <= 0) t0.thin = 1;
        ktimer.begin("changepoint_kernel", stream);
        if (inj) {
            RMN_REQUIRE(inj->d_tape, "injected changepoint run needs d_tape");
            launch<true>(T, inj->d_tape, t0, stream);
        } else {
            launch<false>(T, nullptr, t0, stream);
            if (last_sliced) ktimer.name = "changepoint_sliced_kernel";
        }
        ktimer.end(stream);
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
    int chain_sums(const double** S1, const double** S2, int64_t* n) override {
        *S1 = st.S1; *S2 = st.S2; *n = diag_samples;
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

