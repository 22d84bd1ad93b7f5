// fnn_modes.cuh — the Relaxed and Random selection strategies on the device.
// Included by fnn_order.cu inside its anonymous namespace (needs DevState).
//
//   Random  (NeighborNetRandom.java:31-48, :130-178): fully device-resident.  One lane replays the
//           reference's java.util.Random walk (the draw sequence is inherently serial: each bound
//           depends on the node reached by the previous draw), the block evaluates the sampled Q
//           values in parallel and reduces on (Q, draw index) = "first strict minimum in draw order".
//   Relaxed (NeighborNetLocal.java:88-126): k_rowmin scans one node's row over ALL active nodes in
//           position order and returns the minimum and every exact tie; the sampling / mutual-
//           nearest logic of :170-264 is control flow and runs on the host (fnn_order.cu).
#pragma once

namespace modes {

constexpr int THREADS = 1024;
constexpr int MAX_TIES = 4096;

// ---- java.util.Random on the device -------------------------------------------------------
__device__ __forceinline__ int jr_next(unsigned long long& s, int bits) {
    s = (s * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
    return (int)((long long)s >> (48 - bits));
}
__device__ inline int jr_next_int(unsigned long long& s, int bound) {
    int r = jr_next(s, 31);
    const int m = bound - 1;
    if ((bound & m) == 0) return (int)(((long long)bound * (long long)r) >> 31);
    for (int u = r;; u = jr_next(s, 31)) {
        r = u % bound;
        if ((int)((unsigned)u - (unsigned)r + (unsigned)m) >= 0) break;   // Java int wrap-around
    }
    return r;
}

// cluster distance with the reference's role order: p first, q second (e.g. NeighborNetRandom.java:161-168)
__device__ __forceinline__ double dpq_roles(const double* __restrict__ D, int64_t ld, int sp, int sq, int P2) {
    const int spn = sp < P2 ? (sp ^ 1) : -1, sqn = sq < P2 ? (sq ^ 1) : -1;
    const double* rp = D + (int64_t)sp * ld;
    if (spn < 0 && sqn < 0) return rp[sq];
    if (spn >= 0 && sqn < 0) return (rp[sq] + D[(int64_t)spn * ld + sq]) * 0.5;
    if (spn < 0) return (rp[sq] + rp[sqn]) * 0.5;
    const double* rpn = D + (int64_t)spn * ld;
    return (((rp[sq] + rp[sqn]) + rpn[sq]) + rpn[sqn]) * 0.25;
}

// ceil(log10(total)) as Java computes it (Math.log10 is exact on powers of ten): smallest k with 10^k >= total
__device__ __forceinline__ int ceil_log10(int total) {
    int k = 0;
    long long p = 1;
    while (p < total) { p *= 10; ++k; }
    return k;
}

// ---------------------------------------------------------------- Random findNodes
__global__ void __launch_bounds__(THREADS, 1)
k_random_select(const double* __restrict__ D, int64_t ld, const double* __restrict__ Sx, const int* __restrict__ pos,
                const int* __restrict__ p2s, DevState* st) {
    if (st->done || st->mode < 2 || st->m <= st->fallback) return;
    __shared__ int si[THREADS], sj[THREADS];
    __shared__ int scount;
    __shared__ double wq[THREADS / 32];
    __shared__ unsigned long long wk[THREADS / 32];
    __shared__ int wi[THREADS / 32], wj[THREADS / 32];
    const int m = st->m, P2 = st->P2, tid = threadIdx.x;
    const double cm2 = (double)st->c - 2.0;
    long long amount;   // findSearchAmount (:31-48)
    if (st->mode == 4) amount = ceil_log10(m);
    else if (st->mode == 2) amount = m;
    else amount = (long long)ceil_log10(m) * m;
    const long long searchAmount = (long long)st->mult * amount;

    unsigned long long rng = st->rng;
    int cur = 0;
    if (tid == 0) cur = jr_next_int(rng, m);
    double bq = 0.0;
    unsigned long long bk = ~0ull;
    int bi = -1, bj = -1;
    for (long long base = 0; base < searchAmount; base += THREADS) {
        if (tid == 0) {
            const int cnt = (int)min((long long)THREADS, searchAmount - base);
            for (int k = 0; k < cnt; ++k) {   // the walk of :140-158
                const int i = cur;
                const int s = p2s[i];
                int j;
                if (s < P2) {
                    const int iNbr = pos[s ^ 1];
                    j = jr_next_int(rng, m - 2);
                    if (i == j && m - 1 == iNbr) j = m - 2;
                    else if (i == j && m - 1 != iNbr) j = m - 1;
                    else if (iNbr == j && m - 2 == i) j = m - 1;
                    else if (iNbr == j && m - 2 != i) j = m - 2;
                } else {
                    j = jr_next_int(rng, m - 1);
                    if (i == j) j = m - 1;
                }
                si[k] = i; sj[k] = j;
                cur = j;
            }
            scount = cnt;
        }
        __syncthreads();
        if (tid < scount) {
            const int i = si[tid], j = sj[tid];
            const int sp = p2s[i], sq = p2s[j];
            const double q = (cm2 * dpq_roles(D, ld, sp, sq, P2) - Sx[sp]) - Sx[sq];
            const unsigned long long key = (unsigned long long)(base + tid);
            if (bi < 0 || q < bq || (q == bq && key < bk)) { bq = q; bk = key; bi = i; bj = j; }
        }
        __syncthreads();
    }
    // block reduce on (Q, draw index); invalid lanes carry bi < 0
    for (int off = 16; off > 0; off >>= 1) {
        const double oq = __shfl_down_sync(0xffffffffu, bq, off);
        const unsigned long long ok = __shfl_down_sync(0xffffffffu, bk, off);
        const int oi = __shfl_down_sync(0xffffffffu, bi, off), oj = __shfl_down_sync(0xffffffffu, bj, off);
        if (oi >= 0 && (bi < 0 || oq < bq || (oq == bq && ok < bk))) { bq = oq; bk = ok; bi = oi; bj = oj; }
    }
    if ((tid & 31) == 0) { wq[tid >> 5] = bq; wk[tid >> 5] = bk; wi[tid >> 5] = bi; wj[tid >> 5] = bj; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < THREADS / 32; ++w)
            if (wi[w] >= 0 && (bi < 0 || wq[w] < bq || (wq[w] == bq && wk[w] < bk))) { bq = wq[w]; bk = wk[w]; bi = wi[w]; bj = wj[w]; }
        st->cx_pos = bi;   // Cx = p, Cy = q (:173-176)
        st->cy_pos = bj;
        st->rng = rng;
    }
}

// ---------------------------------------------------------------- Relaxed findRowMin
struct RowMinOut { double value; int count; int overflow; int pos[MAX_TIES]; };

__global__ void __launch_bounds__(THREADS, 1)
k_rowmin(const double* __restrict__ D, int64_t ld, const double* __restrict__ Sx, const int* __restrict__ p2s,
         const DevState* st, int ip, RowMinOut* out) {
    __shared__ double wmin[THREADS / 32];
    __shared__ double gmin;
    __shared__ int cnt;
    const int m = st->m, P2 = st->P2, tid = threadIdx.x;
    const double cm2 = (double)st->c - 2.0;
    const int sp = p2s[ip];
    const int spn = sp < P2 ? (sp ^ 1) : -1;
    const double Sp = Sx[sp];
    double mn = INFINITY;
    for (int iq = tid; iq < m; iq += THREADS) {   // every active q except p and p.nbr (:100-105)
        const int sq = p2s[iq];
        if (sq == sp || sq == spn) continue;
        const double q = (cm2 * dpq_roles(D, ld, sp, sq, P2) - Sp) - Sx[sq];
        mn = fmin(mn, q);
    }
    for (int off = 16; off > 0; off >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, off));
    if ((tid & 31) == 0) wmin[tid >> 5] = mn;
    if (tid == 0) cnt = 0;
    __syncthreads();
    if (tid == 0) {
        double g = wmin[0];
        for (int w = 1; w < THREADS / 32; ++w) g = fmin(g, wmin[w]);
        gmin = g;
    }
    __syncthreads();
    const double g = gmin;
    for (int iq = tid; iq < m; iq += THREADS) {   // all exact ties (:118-124), gathered then ordered by position
        const int sq = p2s[iq];
        if (sq == sp || sq == spn) continue;
        const double q = (cm2 * dpq_roles(D, ld, sp, sq, P2) - Sp) - Sx[sq];
        if (q == g) {
            const int k = atomicAdd(&cnt, 1);
            if (k < MAX_TIES) out->pos[k] = iq;
        }
    }
    __syncthreads();
    if (tid == 0) {
        const int n = min(cnt, MAX_TIES);
        for (int a = 1; a < n; ++a) {   // insertion sort: the list has one or two entries in practice
            const int v = out->pos[a];
            int b = a - 1;
            while (b >= 0 && out->pos[b] > v) { out->pos[b + 1] = out->pos[b]; --b; }
            out->pos[b + 1] = v;
        }
        out->value = g;
        out->count = n;
        out->overflow = cnt > MAX_TIES;
    }
}


// ---------------------------------------------------------------- Relaxed -additive look-ahead
// findAgglomeratedQ (NeighborNetLocal.java:280-386) with agg3wayLocal (:388-414) / agg4wayLocal (:416-466):
// simulate joining (Cx, Cy) and return the Q of the merged cluster against testNode, plus the Q before
// the join (:241-245).  Cx, Cy are taken as selected (no id swap, as in the reference call at :245).
struct LookOut { double origQ, newQ; };

__global__ void __launch_bounds__(THREADS, 1)
k_lookahead(const double* __restrict__ D, int64_t ld, const double* __restrict__ Sx, const int* __restrict__ p2s,
            const DevState* st, int cx_pos, int cy_pos, int test_pos, LookOut* out) {
    extern __shared__ unsigned char smem_raw[];
    xsum::Smem* xs = reinterpret_cast<xsum::Smem*>(smem_raw);
    __shared__ double rx[4], crs[1];
    __shared__ int sh[2];
    const int m = st->m, c = st->c, P2 = st->P2, tid = threadIdx.x;
    const int Cx = p2s[cx_pos], Cy = p2s[cy_pos], T = p2s[test_pos];
    const int Cxn = Cx < P2 ? (Cx ^ 1) : -1, Cyn = Cy < P2 ? (Cy ^ 1) : -1;
    const int L = (m + xsum::THREADS - 1) / xsum::THREADS;
    auto d = [&](int a, int b) -> double { return D[(int64_t)a * ld + b]; };
    const int zs[4] = {Cx, Cxn, Cy, Cyn};
    if (Cxn >= 0 || Cyn >= 0) {
        auto load = [&](int r, int t, int k) -> double {
            const int s = p2s[t * L + k];
            const double v = d(zs[r], s);
            const bool full = (s >= P2) || s == Cx || s == Cxn || s == Cy || s == Cyn;
            return full ? v : v * 0.5;
        };
        xsum::block_exact_seq_sum<4>(xs, m, load, [&](int r) { return zs[r] >= 0; }, rx);
    } else {
        if (tid < 4) rx[tid] = 0.0;
        __syncthreads();
    }
    if (tid == 0) {   // the 4-candidate pick (:311-335)
        int x = Cx, y = Cy;
        const double f = (double)(c + (Cxn >= 0) + (Cyn >= 0)) - 2.0;
        double best = (f * d(Cx, Cy) - rx[0]) - rx[2];
        if (Cxn >= 0) { const double q = (f * d(Cxn, Cy) - rx[1]) - rx[2]; if (q < best) { x = Cxn; y = Cy; best = q; } }
        if (Cyn >= 0) { const double q = (f * d(Cx, Cyn) - rx[0]) - rx[3]; if (q < best) { x = Cx; y = Cyn; best = q; } }
        if (Cxn >= 0 && Cyn >= 0) { const double q = (f * d(Cxn, Cyn) - rx[1]) - rx[3]; if (q < best) { x = Cxn; y = Cyn; best = q; } }
        sh[0] = x; sh[1] = y;
    }
    __syncthreads();
    const int x = sh[0], y = sh[1];
    const int xn = x < P2 ? (x ^ 1) : -1, yn = y < P2 ? (y ^ 1) : -1;
    int kind, X = -1, Y = -1, Z = -1, W = -1;
    if (xn < 0 && yn < 0) kind = 2;
    else if (xn < 0) { kind = 3; X = x; Y = y; Z = yn; }
    else if (yn < 0 || m == 4) { kind = 3; X = y; Y = x; Z = xn; }
    else { kind = 4; X = xn; Y = x; Z = y; W = yn; }   // (x2, x, y, y2)
    const double TT = 2.0 / 3.0;
    auto term = [&](int i) -> double {   // the per-position contribution to clusterRowSum
        const int p = p2s[i];
        const bool pair = p < P2;
        if (pair && (p & 1)) return 0.0;                       // not a representative
        const int pn = pair ? p + 1 : -1;
        if (kind == 2) {
            if (p == x || p == y) return 0.0;
            if (!pair) return (d(p, x) + d(p, y)) * 0.5;
            return (((d(p, x) + d(p, y)) + d(pn, x)) + d(pn, y)) * 0.25;
        } else if (kind == 3) {
            if (p == X || p == Y || p == Z || pn == Y || pn == Z) return 0.0;
            const double Dup = TT * d(X, p) + d(Y, p) / 3.0, Dvp = TT * d(Z, p) + d(Y, p) / 3.0;
            if (!pair) return (Dup + Dvp) * 0.5;
            const double Duq = TT * d(X, pn) + d(Y, pn) / 3.0, Dvq = TT * d(Z, pn) + d(Y, pn) / 3.0;
            return (((Dup + Dvp) + Duq) + Dvq) * 0.25;
        } else {
            if (p == X || p == Y || p == Z || p == W || pn == X || pn == Y || pn == Z || pn == W) return 0.0;
            const double Dup = TT * d(X, p) + d(Y, p) / 3.0, Dvp = TT * d(Z, p) + d(Y, p) / 3.0;
            const double Dup2 = TT * Dup + Dvp / 3.0, Dvp2 = TT * d(W, p) + Dvp / 3.0;
            if (!pair) return (Dup2 + Dvp2) * 0.5;
            const double Duq = TT * d(X, pn) + d(Y, pn) / 3.0, Dvq = TT * d(Z, pn) + d(Y, pn) / 3.0;
            const double Duq2 = TT * Duq + Dvq / 3.0, Dvq2 = TT * d(W, pn) + Dvq / 3.0;
            return (((Dup2 + Dvp2) + Duq2) + Dvq2) * 0.25;
        }
    };
    auto load1 = [&](int, int t, int k) -> double { return term(t * L + k); };
    xsum::block_exact_seq_sum<1>(xs, m, load1, [](int) { return true; }, crs);
    if (tid == 0) {
        const double dCxT = dpq_roles(D, ld, Cx, T, P2), dCyT = dpq_roles(D, ld, Cy, T, P2);
        const double subtracted = (Sx[T] - dCxT) - dCyT;                                  // :337
        double cd;
        if (kind == 2) {
            const int Tn = T < P2 ? (T ^ 1) : -1;
            // calculateClusterDistLocal(x, testNode) (:349, :468-476); a singleton testNode makes the JAR throw there -
            // the oracle's documented substitute is the 2-term form against (x, y)
            cd = Tn >= 0 ? (d(x, T) + d(x, Tn)) * 0.5 : (d(T, x) + d(T, y)) * 0.5;
        } else cd = term(test_pos);
        const double A = ((double)c - 1.0) - 2.0;
        out->newQ = (A * cd - crs[0]) - (subtracted + cd);                                // :361 / :412 / :464
        out->origQ = (((double)c - 2.0) * dCxT - Sx[Cx]) - Sx[T];                         // :241-243
    }
}

}  // namespace modes
