// riemann_b200 -- thread-per-chain changepoint kernel (see changepoint.cu for the model / proposal
// semantics and the reference citations; both kernels replay identical injected streams bit-for-bit).
//
// BUILD NOTE: this translation unit is compiled with `-Xcicc -O1` (ptxas stays at -O3).  With the NVVM
// mid-level optimiser at -O2/-O3 (CUDA 12.9) the kernel hangs or raises "misaligned address" for certain
// warps whose 32 chains take different move types -- reproducible from a saved state, absent with
// -Xcicc -O1 or -G, independent of ptxas -O level, and not tied to any one loop (bisected in round 1).
#include "changepoint.cuh"

namespace cp {

// =======================================================================================
// Thread-per-chain variant.  One thread owns one chain; cpx / cpv / a backup of the moved block /
// two generations of run boundaries live in shared memory as [slot][thread] (bank-conflict free),
// so variable k needs no register indexing.  Moves are applied IN PLACE and undone on reject
// (acceptance is ~17 %).  Every Philox block, every fp64 log and every search step is executed
// once per warp for 32 chains (the 16-lane kernel above serves 2).  The sum of log-gaps becomes
// one log of their product (fp64 has the range: at most 16 factors in (0, xmax-xmin]).
// =======================================================================================
// Every loop of the step runs to a WARP-UNIFORM trip count (warp maximum of the per-thread bound)
// with a predicated body, so no loop has a divergent exit and no loop sits inside a divergent branch.
__device__ __forceinline__ int wmax(int v) { return __reduce_max_sync(0xffffffffu, v); }

template <bool INJ, int NT>
__global__ void __launch_bounds__(NT)
changepoint_tpc_kernel(const __grid_constant__ CPParams P, const double* __restrict__ gdata, CPState st,
                       int64_t K, int64_t T, int64_t step0, uint64_t seed, int64_t chain_offset,
                       const double* __restrict__ tape, rmn_trace_t tr) {
    extern __shared__ double smem[];
    const int ndata = 3 * P.M + 2;
    const int nd_al = (ndata + 1) & ~1;
    for (int i = threadIdx.x; i < ndata; i += NT) smem[i] = gdata[i];
    const double* xs = smem;
    const double* cy = xs + P.M;
    const double* cyy = cy + P.M + 1;
    double* CX = smem + nd_al + threadIdx.x;                 // element j at CX[j * NT]
    double* CV = CX + LANES * NT;                            // CV follows CX: CX[LANES*NT + j*NT] == CV[j*NT]
    double* BK = CV + LANES * NT;
    int* BU = reinterpret_cast<int*>(smem + nd_al + 3 * LANES * NT) + threadIdx.x;   // two generations
    constexpr int GEN = LANES * NT;
    double* XI = reinterpret_cast<double*>(BU + 2 * GEN - threadIdx.x) + threadIdx.x;   // [LANES][NT] noise of the step

    const int64_t c_raw = blockIdx.x * (int64_t)NT + threadIdx.x;
    const bool live = c_raw < K;
    const int64_t c = live ? c_raw : K - 1;

    int k = st.k[c];
    double sig = st.sig[c], lp = st.lp[c];
    __syncthreads();
    for (int j = 0; j < LANES; ++j) {
        const double x = st.cpx[c * LANES + j];
        CX[j * NT] = x;
        CV[j * NT] = st.cpv[c * LANES + j];
        BU[j * NT] = (j < k) ? upper_bound(xs, P.M, P.P2, x) : P.M;
    }
    int bsel = 0;
    long long nacc = 0, novf = 0;
    double s1[RMN_CP_NDIAG], s2[RMN_CP_NDIAG];
#pragma unroll
    for (int i = 0; i < RMN_CP_NDIAG; ++i) { s1[i] = 0.0; s2[i] = 0.0; }
    const RngKey rk(seed, (uint64_t)(chain_offset + c));
    const TraceSel ts{tr.first, tr.thin > 0 ? tr.thin : 1};
    const bool tracing = tr.d_k || tr.d_cpx || tr.d_cpv || tr.d_sig || tr.d_logpost;

    for (int64_t t = 0; t < T; ++t) {
        const uint64_t step = (uint64_t)(step0 + t);
        const double* row = INJ ? tape + (t * K + c) * RMN_CP_NSLOT : nullptr;
        double snew, du, uacc;
        int nrand, mv;
        bool birth;
        if (INJ) {
            const double u1 = row[RMN_CP_SLOT_SEL1], u2 = row[RMN_CP_SLOT_SEL2], u3 = row[RMN_CP_SLOT_SEL3];
            mv = (u1 < P.p1) ? 0 : ((u2 < P.p2) ? 1 : ((u3 < P.p3) ? 2 : 3));        // test_changepoint.py:48-54
            birth = (k == 0) || (row[RMN_CP_SLOT_BD] > 0.5);                          // :59
            snew = row[RMN_CP_SLOT_S]; du = row[RMN_CP_SLOT_DU];
            nrand = (int)row[RMN_CP_SLOT_N]; uacc = row[RMN_CP_SLOT_ACC];
        } else {
            const uint4 a = rk.block(step, RMN_BLOCK_AUX);
            const uint4 b = rk.block(step, RMN_BLOCK_AUX2);
            mv = (a.x < P.t1) ? 0 : ((a.y < P.t2) ? 1 : ((a.z < P.t3) ? 2 : 3));
            birth = (k == 0) || (a.w >= 0x80000000u);
            snew = P.xmin + (P.xmax - P.xmin) * u01_fast(b.x);
            du = -0.1 + 0.2 * u01_fast(b.y);
            nrand = (int)(u01_fast(b.z) * (double)k);
            uacc = u01_fast(b.w);
        }
        nrand = max(0, min(nrand, k - 1));

        const int ocur = bsel * GEN, onew = (bsel ^ 1) * GEN;     // offsets of the two boundary generations
        int kk = k;
        double nsig = sig, jarg = 1.0;
        bool ovf = false;
        const int nnorm = (mv == 0) ? k : ((mv == 1) ? k + 1 : ((mv == 2) ? 1 : 0));
        const int aoff = (mv == 0) ? 0 : GEN;                   // moved block: CX or CV
        const double mscale = (mv == 0) ? P.sx[k] : P.sv;
        const int kw = wmax(k);                                 // warp-uniform loop bounds
        const int maxn = wmax(nnorm);

        // ---- noise of the step (uniform loop)
        for (int b4 = 0; b4 < maxn; b4 += 4) {
            double xi[4];
            if (INJ) {
#pragma unroll
                for (int q = 0; q < 4; ++q) xi[q] = (b4 + q < LANES) ? row[RMN_CP_SLOT_XI + b4 + q] : 0.0;
            } else {
                normal4(rk.block(step, (uint32_t)(b4 >> 2)), xi);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) XI[(b4 + q) * NT] = xi[q];
        }
        // ---- block moves: theta + scale * L xi with diagonal L (randomwalk.py:26), in place
        for (int j = 0; j < maxn; ++j) {
            if (mv < 2 && j < nnorm) {
                const double old = CX[aoff + j * NT];
                BK[j * NT] = old;
                CX[aoff + j * NT] = __dadd_rn(old, __dmul_rn(mscale, XI[j * NT]));
            }
        }
        if (mv == 2) nsig = __dadd_rn(sig, __dmul_rn(P.ss, XI[0]));
        // run boundaries of moved locations (cpx block move): fixed-trip search, predicated store
        for (int j = 0; j < kw; ++j) {
            const int b = upper_bound(xs, P.M, P.P2, CX[j * NT]);
            if (mv == 0 && j < k) BU[onew + j * NT] = b;
        }
        bool newb = (mv == 0);

        // ---- trans-dimensional moves (changepoint.py:193-240), in place with an undo record
        const bool dimv = (mv == 3);
        const bool want_b = dimv && birth;
        const bool is_d = dimv && !birth;
        int nb = 0;
        for (int j = 0; j < kw; ++j) nb += (want_b && j < k && CX[j * NT] < snew) ? 1 : 0;   // searchsorted(cpx, s)
        const int dn = is_d ? nrand : 0;
        const double hb = CV[nb * NT];
        const double h1 = CV[dn * NT], h2 = CV[min(dn + 1, LANES - 1) * NT];
        const double ub = 0.5 + du / P.sqrtM;                                          // test_changepoint.py:61
        const double fb = sqrt((1.0 - ub) / ub);                                       // changepoint.py:57
        const double hd = sqrt(h1 * h2);                                               // :67
        const double ud = 1.0 / (1.0 + h2 / h1);                                       // :68
        if (want_b) jarg = fabs(hb / (ub * (1.0 - ub)));                               // |J|
        if (is_d) jarg = fabs(hd / (ud * (1.0 - ud)));                                 // 1/|J^-1|
        ovf = want_b && (k + 1 > LANES - 1);
        const bool is_b = want_b && !ovf;
        const int ubs = upper_bound(xs, P.M, P.P2, snew);
        const double sv0 = is_b ? hb : CX[dn * NT];                                    // undo record
        const int dpos = is_b ? nb : dn;
        // birth: shift up (descending, predicated)
        for (int j = kw; j >= 1; --j) {
            if (is_b && j > nb && j <= k) {
                CX[j * NT] = CX[(j - 1) * NT];
                CV[(j + 1) * NT] = CV[j * NT];
            }
        }
        if (is_b) { CX[nb * NT] = snew; CV[nb * NT] = hb / fb; CV[(nb + 1) * NT] = hb * fb; }
        // death: shift down (ascending, predicated)
        for (int j = 0; j < kw; ++j) {
            if (is_d && j >= dn && j < k - 1) CX[j * NT] = CX[(j + 1) * NT];
            if (is_d && j >= dn + 1 && j < k) CV[j * NT] = CV[(j + 1) * NT];
        }
        if (is_d) CV[dn * NT] = hd;
        // boundaries of the new generation
        for (int j = 0; j <= kw; ++j) {
            if (is_b && j <= k) {
                const int src = (j < nb) ? BU[ocur + j * NT] : ((j == nb) ? ubs : BU[ocur + (j - 1) * NT]);
                BU[onew + j * NT] = src;
            }
            if (is_d && j < k - 1) BU[onew + j * NT] = (j < dn) ? BU[ocur + j * NT] : BU[ocur + (j + 1) * NT];
        }
        if (is_b) kk = k + 1;
        if (is_d) kk = k - 1;
        newb = newb || is_b || is_d;

        // ---- log-posterior of the (in-place) proposal (uniform loop, predicated body)
        const int eoff = newb ? onew : ocur;
        const int kkw = wmax(kk);
        double SS = 0.0, prod = 1.0, vsum = 0.0, vprod = 1.0, prevx = P.xmin;
        int bl = 0;
        bool bad = false;
        for (int j = 0; j <= kkw; ++j) {
            if (j <= kk) {
                const double hi = (j < kk) ? CX[j * NT] : P.xmax;
                const int bj = (j < kk) ? BU[eoff + j * NT] : P.M;
                const double v = CV[j * NT];
                const double gap = hi - prevx;
                bad |= !(gap > 0.0) || !(v > 0.0);
                prod *= gap;
                const double n = (double)(bj - bl);
                const double a1 = cy[bj] - cy[bl], a2 = cyy[bj] - cyy[bl];
                const double vc = v - P.ycenter;
                SS += n * vc * vc - 2.0 * vc * a1 + a2;
                vsum += v;
                if (!P.alpha_is_one) vprod *= v;
                prevx = hi; bl = bj;
            }
        }
        const int ks = kk + 1;
        const double s2n = nsig * nsig;
        const double log_s2 = log(s2n);
        const double lg = log(prod);
        const double logu = log(uacc);
        const double ljac = log(jarg);
        double vt = -P.beta * vsum + (double)ks * P.cv;
        if (!P.alpha_is_one) vt += (P.alpha - 1.0) * log(vprod);
        double logl = -0.5 * ((SS / s2n + (double)P.M * log_s2) + P.Mlog2pi);
        if (isnan(logl)) logl = -INFINITY;
        double lsig = -log_s2;                                   // log(1/sigma^2), within 1 ulp
        if (nsig < 0.0) lsig = NAN;
        const double lps = (P.tab2[ks] + lg) - (double)ks * P.logL;
        double logp = ((P.tab1[ks] + vt) + lps) + lsig;
        if (isnan(logp) || bad) logp = -INFINITY;
        const double lpn = combine_logpost(logp, logl);
        const double lqr = dimv ? (birth ? ljac : -ljac) : 0.0;

        const double delta = lpn - lp - lqr;
        const double mh = (delta < 0.0) ? delta : 0.0;               // Python min(0, nan) == 0
        const bool acc = !ovf && (logu < mh);

        if (live) {
            if (tr.d_prop_logpost) tr.d_prop_logpost[t * K + c] = lpn;
            if (tr.d_accepted) tr.d_accepted[t * K + c] = acc ? 1 : 0;
            if (tr.d_logqratio) tr.d_logqratio[t * K + c] = lqr;
            if (tr.d_prop_k) tr.d_prop_k[t * K + c] = kk;
            if (tr.d_prop_sig) tr.d_prop_sig[t * K + c] = nsig;
            if (tr.d_prop_cpx)
                for (int j = 0; j < LANES; ++j) tr.d_prop_cpx[(t * K + c) * LANES + j] = (j < kk) ? CX[j * NT] : 0.0;
            if (tr.d_prop_cpv)
                for (int j = 0; j < LANES; ++j) tr.d_prop_cpv[(t * K + c) * LANES + j] = (j <= kk) ? CV[j * NT] : 0.0;
        }

        // ---- accept, or undo the in-place move (uniform loops, predicated bodies)
        const bool undo_blk = !acc && mv < 2;
        const bool undo_b = !acc && is_b, undo_d = !acc && is_d;
        for (int j = 0; j < maxn; ++j)
            if (undo_blk && j < nnorm) CX[aoff + j * NT] = BK[j * NT];
        for (int j = 0; j < kw; ++j) {                       // undo insert at dpos (old k)
            if (undo_b && j >= dpos && j < k) {
                CX[j * NT] = CX[(j + 1) * NT];
                CV[(j + 1) * NT] = CV[(j + 2) * NT];
            }
        }
        if (undo_b) CV[dpos * NT] = sv0;
        for (int j = kw; j >= 1; --j) {                      // undo delete at dpos (old k)
            if (undo_d && j > dpos && j <= k - 1) CX[j * NT] = CX[(j - 1) * NT];
            if (undo_d && j > dpos + 1 && j <= k) CV[j * NT] = CV[(j - 1) * NT];
        }
        if (undo_d) { CX[dpos * NT] = sv0; CV[dpos * NT] = h1; CV[(dpos + 1) * NT] = h2; }
        novf += ovf ? 1 : 0;
        if (acc) {
            k = kk; sig = nsig; lp = lpn;
            if (newb) bsel ^= 1;
            nacc += 1;
        }

        if ((step % RMN_CP_DIAG_EVERY) == 0) {
            int cnt[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) cnt[q] = 0;
            const int kw2 = wmax(k);
            for (int j = 0; j < kw2; ++j) {
                const double cc = CX[j * NT];
#pragma unroll
                for (int q = 0; q < NQ; ++q) cnt[q] += (j < k && cc < P.xq[q]) ? 1 : 0;
            }
            s1[0] += sig; s2[0] += sig * sig;
            s1[1] += (double)k; s2[1] += (double)k * (double)k;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const double yq = CV[cnt[q] * NT];
                s1[2 + q] += yq; s2[2 + q] += yq * yq;
            }
        }

        if (live && tracing) {
            const long long r = ts.slot(t + 1);
            if (r >= 0) {
                if (tr.d_cpx)
                    for (int j = 0; j < LANES; ++j) tr.d_cpx[(r * K + c) * LANES + j] = (j < k) ? CX[j * NT] : 0.0;
                if (tr.d_cpv)
                    for (int j = 0; j < LANES; ++j) tr.d_cpv[(r * K + c) * LANES + j] = (j <= k) ? CV[j * NT] : 0.0;
                if (tr.d_k) tr.d_k[r * K + c] = k;
                if (tr.d_sig) tr.d_sig[r * K + c] = sig;
                if (tr.d_logpost) tr.d_logpost[r * K + c] = lp;
            }
        }
    }

    if (live) {
        for (int j = 0; j < LANES; ++j) {
            st.cpx[c * LANES + j] = (j < k) ? CX[j * NT] : 0.0;
            st.cpv[c * LANES + j] = (j <= k) ? CV[j * NT] : 0.0;
        }
        st.k[c] = k; st.sig[c] = sig; st.lp[c] = lp;
        st.dacc[c] += nacc; st.dovf[c] += novf;
#pragma unroll
        for (int i = 0; i < RMN_CP_NDIAG; ++i) {
            st.S1[(int64_t)i * K + c] += s1[i];
            st.S2[(int64_t)i * K + c] += s2[i];
        }
    }
}


static constexpr int TPC_NT = 64;

size_t tpc_smem_bytes(int M) {
    const size_t nd_al = ((size_t)(3 * M + 2) + 1) & ~size_t(1);
    return (nd_al + 4 * LANES * TPC_NT) * 8 + (size_t)2 * LANES * TPC_NT * 4;
}

void tpc_launch(bool inj, const CPParams& P, const double* gdata, const CPState& st, int64_t K, int64_t T,
                int64_t step0, uint64_t seed, int64_t chain_offset, const double* tape, const rmn_trace_t& tr,
                cudaStream_t stream) {
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(changepoint_tpc_kernel<false, TPC_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(changepoint_tpc_kernel<true, TPC_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr = true;
    }
    const size_t sm = tpc_smem_bytes(P.M);
    const unsigned grid = (unsigned)((K + TPC_NT - 1) / TPC_NT);
    if (inj) changepoint_tpc_kernel<true, TPC_NT><<<grid, TPC_NT, sm, stream>>>(P, gdata, st, K, T, step0, seed, chain_offset, tape, tr);
    else changepoint_tpc_kernel<false, TPC_NT><<<grid, TPC_NT, sm, stream>>>(P, gdata, st, K, T, step0, seed, chain_offset, tape, tr);
}

}  // namespace cp
