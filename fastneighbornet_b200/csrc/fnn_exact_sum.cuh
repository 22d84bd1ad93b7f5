// fnn_exact_sum.cuh — bit-exact LEFT-TO-RIGHT fp64 summation in parallel.
//
// The reference accumulates ComputeRx (NetMakerOriginal.java:549-561) and u.Sx (:530-535) with
// `sum += x` in position order; those bits feed later argmins, so the rounding sequence must be
// reproduced, and a dependent DADD chain costs ~5 ns per element on one lane.
//
// Observation: while the running sum s stays inside one binade [2^e, 2^(e+1)) and the addends
// are non-negative, every step is s <- s + RN_ulp(a) with ulp = 2^(e-52), i.e. INTEGER addition
// of R(a) = round-to-nearest(a / ulp) to the 53-bit significand (a tie a/ulp = k + 1/2 is the only
// case where the increment depends on s, via round-to-even).  Integer addition is associative, so
// a run of such steps collapses to one add of the pre-summed increments.
//
// Algorithm (one block of 1024 threads, one pass over the chain):
//   A. thread t owns the contiguous segment [t*L, (t+1)*L); an approximate prefix sum (any order)
//      places the segment in a binade e; if its whole uncertainty interval lies in that binade and
//      its elements are non-negative, finite and not ties, the segment gets the integer increment
//      C = sum rint(a * 2^(52-e)) (exact in fp64 while < 2^53).  32 segments combine to a warp
//      summary when they agree on e.
//   B. one warp per chain walks: 32 warp summaries -> 32 segment summaries of a failing warp ->
//      the elements of a failing segment (plain sequential adds).  At the two upper levels the
//      warp applies the longest applicable PREFIX of the 32 summaries in one cooperative step.
//      Every application is VERIFIED exactly: exponent(s) == e before, significand + C < 2^53
//      after.  Because increments are non-negative, that proves no step of the run left the
//      binade, hence the collapsed result equals the sequential one bit for bit.  Whatever is
//      not verifiable (binade crossings, ties, negative / non-finite addends, s == 0) is summed
//      sequentially.  The approximation only steers efficiency, never the result.
#pragma once

namespace xsum {

constexpr int THREADS = 1024;
constexpr int LEAF_MAX = 128;        // max segment length (chains up to 131072 elements)
constexpr int PROLOGUE = 512;        // the first ~512 elements hold most binade crossings: summed sequentially up front
constexpr int LEAF_BUF = PROLOGUE + LEAF_MAX;
constexpr double TWO53 = 9007199254740992.0;
constexpr unsigned long long M52 = (1ull << 52) - 1;
constexpr int E_INVALID = 0, E_IDENT = -1;   // summary exponent codes (else: biased exponent 1..2046)

template <int NR>
struct SmemN {
    double wtot[NR][32];
    double wC[NR][32];
    int wE[NR][32];
    double cC[NR][THREADS];
    int cE[NR][THREADS];
    double leaf[NR][LEAF_BUF];
    double result[NR];
};
using Smem = SmemN<4>;

// longest applicable prefix of the 32 lane summaries (lanes < start are already consumed).
// s is uniform across the warp.  Returns the index of the first summary that was NOT applied.
__device__ __forceinline__ int coop_apply(double& s, int e_lane, double C_lane, int start, int lane) {
    double P = (lane >= start && e_lane > 0) ? C_lane : 0.0;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, P, off);
        if (lane >= off) P += t;
    }
    const unsigned long long b = (unsigned long long)__double_as_longlong(s);
    const int es = (int)(b >> 52);                         // sign bit set -> never equals a summary exponent
    const double msd = (double)((b & M52) | (1ull << 52));
    const bool structural = (lane < start) || e_lane == E_IDENT || (e_lane > 0 && e_lane == es);
    const unsigned bad = ~__ballot_sync(0xffffffffu, structural);
    const int f = bad ? (__ffs(bad) - 1) : 32;
    const bool ok = lane >= start && lane < f && (msd + P < TWO53);
    const int p = __popc(__ballot_sync(0xffffffffu, ok));   // ok lanes form a contiguous run from `start` (P is monotone)
    if (p > 0) {
        const double Pp = __shfl_sync(0xffffffffu, P, start + p - 1);
        if (Pp > 0.0) {
            const unsigned long long tot = (unsigned long long)(msd + Pp);
            s = __longlong_as_double((long long)(((unsigned long long)es << 52) | (tot & M52)));
        }
    }
    return start + p;
}

// buf unused (kept for the serial variant's signature symmetry).  All THREADS threads call.
// load(r, t, k): element k of segment t (= element t*L + k) of chain r, with L = ceil(len/1024) - callers may store
// chains segment-transposed so that the 32 lanes of a warp read consecutive addresses; present(r): whether chain r
// exists; out[r] = sequential sum.
template <int NR, typename SmemT, typename Loader, typename Present>
__device__ void block_exact_seq_sum(SmemT* sm, int len, Loader load, Present present, double* out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = (len + THREADS - 1) / THREADS;            // <= LEAF_MAX
    const int i0 = min(tid * L, len), i1 = min(i0 + L, len);
    // ---- A1: local sums -> approximate exclusive prefix
#ifdef FNN_XSUM_TIMING
    long long t0_ = clock64();
#endif
    double ls[NR], incl[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) ls[r] = 0.0;
    for (int i = i0; i < i1; i += 4) {   // 4 x NR independent loads in flight
        double a[NR][4];
#pragma unroll
        for (int r = 0; r < NR; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) a[r][q] = (present(r) && i + q < i1) ? load(r, tid, i + q - i0) : 0.0;
#pragma unroll
        for (int r = 0; r < NR; ++r) ls[r] += (a[r][0] + a[r][1]) + (a[r][2] + a[r][3]);
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        double v = ls[r];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, v, off);
            if (lane >= off) v += t;
        }
        incl[r] = v;
        if (lane == 31) sm->wtot[r][warp] = v;
    }
    __syncthreads();
    if (warp < NR) {
        double v = sm->wtot[warp][lane];
        const double own = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, v, off);
            if (lane >= off) v += t;
        }
        sm->wtot[warp][lane] = v - own;
    }
    __syncthreads();
#ifdef FNN_XSUM_TIMING
    long long t1_ = clock64();
#endif
    // ---- A2: segment summaries, warp summaries
    {
        int eb[NR];
        double scale[NR], C[NR];
        bool bad[NR], anynz[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const double pstart = sm->wtot[r][warp] + (incl[r] - ls[r]);
            const double pend = pstart + ls[r];
            const double eps = 4.0e-11;   // >> error of the approximate prefix (<= ~2^17 terms, tree order)
            const unsigned long long lo = (unsigned long long)__double_as_longlong(pstart * (1.0 - eps));
            const unsigned long long hi = (unsigned long long)__double_as_longlong(pend * (1.0 + eps));
            eb[r] = (int)(lo >> 52);
            bad[r] = !(pstart > 0.0) || eb[r] != (int)(hi >> 52) || eb[r] < 123 || eb[r] > 1923;
            scale[r] = __longlong_as_double((long long)(2098 - min(max(eb[r], 123), 1923)) << 52);   // 2^(52 - e)
            C[r] = 0.0;
            anynz[r] = false;
        }
        for (int i = i0; i < i1; i += 4) {
            double a[NR][4];
#pragma unroll
            for (int r = 0; r < NR; ++r)
#pragma unroll
                for (int q = 0; q < 4; ++q) a[r][q] = (present(r) && i + q < i1) ? load(r, tid, i + q - i0) : 0.0;
#pragma unroll
            for (int r = 0; r < NR; ++r)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double x = a[r][q] * scale[r];
                    // round-to-nearest-even to an integer without FRND.F64: exact for 0 <= x < 2^52 under the default
                    // rounding mode (no FMA contraction, no fast-math), and x >= 2^52 is already an integer; x < 0 marks
                    // the segment invalid below, whatever R is
                    const double R = (x < 4503599627370496.0) ? (x + 4503599627370496.0) - 4503599627370496.0 : x;
                    bad[r] |= (a[r][q] < 0.0) | (fabs(x - R) == 0.5);
                    anynz[r] |= (a[r][q] != 0.0);
                    C[r] += R;
                }
        }
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            int e;
            double Cr = C[r];
            if (!present(r) || i1 <= i0 || !anynz[r]) { e = E_IDENT; Cr = 0.0; }
            else if (bad[r] || !(Cr < TWO53)) { e = E_INVALID; Cr = 0.0; }
            else e = eb[r];
            sm->cE[r][tid] = e;
            sm->cC[r][tid] = Cr;
            // warp summary: all segments applicable and agreeing on the binade (identities are neutral)
            const unsigned real = __ballot_sync(0xffffffffu, e > 0);
            const int e0 = real ? __shfl_sync(0xffffffffu, e, __ffs(real) - 1) : E_IDENT;
            const bool agree = __all_sync(0xffffffffu, e == E_IDENT || e == e0);
            double c = Cr;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
            if (lane == 0) {
                const bool okw = agree && (c < TWO53);
                sm->wE[r][warp] = okw ? e0 : E_INVALID;
                sm->wC[r][warp] = okw ? c : 0.0;
            }
        }
    }
    __syncthreads();
#ifdef FNN_XSUM_TIMING
    long long t2_ = clock64(), tp_ = 0;
    int ncoop_ = 0, nleaf_ = 0;
#endif
    // ---- B: warp r walks chain r
    if (warp < NR) {
        const int r = warp;
        double s = 0.0;
        if (present(r)) {
            const int we = sm->wE[r][lane];
            const double wc = sm->wC[r][lane];
            // sequential adds of elements [j0, j1) staged through shared memory (s stays uniform across the warp)
            auto leaf_sum = [&](int j0, int j1) {
#ifdef FNN_XSUM_TIMING
                nleaf_ += j1 - j0;
#endif
                for (int k = lane; k < j1 - j0; k += 32) sm->leaf[r][k] = load(r, (j0 + k) / L, (j0 + k) % L);
                __syncwarp();
                const double* e = sm->leaf[r];
                int k = 0;
                for (; k + 4 <= j1 - j0; k += 4) {
                    const double v0 = e[k], v1 = e[k + 1], v2 = e[k + 2], v3 = e[k + 3];
                    s += v0; s += v1; s += v2; s += v3;
                }
                for (; k < j1 - j0; ++k) s += e[k];
                __syncwarp();
            };
            // open warp summary w from its segment s1: cooperative prefixes, sequential leaves
            auto open_warp = [&](int w, int s1) {
                const int te = sm->cE[r][w * 32 + lane];
                const double tc = sm->cC[r][w * 32 + lane];
                while (s1 < 32) {
#ifdef FNN_XSUM_TIMING
                    ++ncoop_;
#endif
                    s1 = coop_apply(s, te, tc, s1, lane);
                    if (s1 >= 32) break;
                    const int t = w * 32 + s1;
                    leaf_sum(min(t * L, len), min(t * L + L, len));
                    s1 += 1;
                }
            };
            // prologue: the running sum doubles every few elements at the start (binade crossings at ~2^j / mean),
            // so the first segments are never collapsible - add them sequentially in one go
            const int S0 = min(32, (PROLOGUE + L - 1) / L);
            leaf_sum(0, min(S0 * L, len));
#ifdef FNN_XSUM_TIMING
            tp_ = clock64();
#endif
            open_warp(0, S0);
            int s2 = 1;
            while (s2 < 32) {
#ifdef FNN_XSUM_TIMING
                ++ncoop_;
#endif
                s2 = coop_apply(s, we, wc, s2, lane);
                if (s2 >= 32) break;
                open_warp(s2, 0);
                s2 += 1;
            }
        }
        if (lane == 0) sm->result[r] = s;
#ifdef FNN_XSUM_TIMING
        if (lane == 0) printf("chain %d: A1=%lld A2=%lld prologue=%lld walk(after prologue)=%lld cycles, coop steps=%d leaf elements=%d L=%d\n", r, t1_ - t0_, t2_ - t1_, tp_ - t2_, clock64() - tp_, ncoop_, nleaf_, L);
#endif
    }
    __syncthreads();
    if (tid < NR) out[tid] = sm->result[tid];
    __syncthreads();
}

}  // namespace xsum
