// fnn_exact_sum.cuh — bit-exact LEFT-TO-RIGHT fp64 summation in parallel.
//
// The reference accumulates ComputeRx (NetMakerOriginal.java:549-561) and u.Sx (:530-535) with
// `sum += x` in position order; those bits feed later argmins, so the rounding sequence must be
// reproduced, and a dependent DADD chain costs ~4.8 ns per element on one lane.
//
// Observation: while the running sum s stays inside one binade [2^e, 2^(e+1)) and the addends
// are non-negative, every step is s <- s + RN_ulp(a) with ulp = 2^(e-52), i.e. INTEGER addition
// of R(a) = round-to-nearest(a / ulp) to the 53-bit significand (a tie a/ulp = k + 1/2 is the only
// case where the increment depends on s, via round-to-even).  Integer addition is associative, so
// a run of such steps collapses to one add of the pre-summed increments.
//
// Algorithm per 4096-element tile (1024 threads x 4 consecutive elements):
//   1. approximate prefix sums (any order) place each 4-element chunk in a binade e;
//   2. each chunk whose whole uncertainty interval lies in one binade and whose elements are
//      non-negative, normal (or zero), not ties, gets the integer increment C = sum R(a);
//      chunk -> warp -> tile summaries (all members same e) are combined by shuffles;
//   3. one walker lane per chain applies summaries top-down.  Every application is VERIFIED
//      exactly: exponent(s) == e before, significand + C < 2^53 after.  Because increments are
//      non-negative, that proves no step of the run left the binade, hence the collapsed result
//      equals the sequential one bit for bit.  Anything unverifiable (binade crossings, ties,
//      negative / subnormal / non-finite addends, s == 0) falls back to plain sequential adds
//      of that chunk.  The approximation only steers efficiency, never the result.
#pragma once

namespace xsum {

constexpr int THREADS = 1024;
constexpr int L = 4;                 // elements per thread
constexpr int TILE = THREADS * L;    // 4096
constexpr unsigned long long M52 = (1ull << 52) - 1;
constexpr unsigned long long B53 = 1ull << 53;

// summary word: 0 = not collapsible; else (biased exponent << 53) | increment (< 2^53)
__device__ __forceinline__ unsigned long long chunk_summary(const double* a, double pstart, double pend) {
    const double eps = 3.7e-12;   // >= 2^-38: bounds the error of the approximate in-tile prefix
    if (!(pstart > 0.0)) return 0;
    const unsigned long long lo = (unsigned long long)__double_as_longlong(pstart * (1.0 - eps));
    const unsigned long long hi = (unsigned long long)__double_as_longlong(pend * (1.0 + eps));
    const unsigned long long eb = (lo >> 52) & 0x7ff;
    if (eb == 0 || eb == 0x7ff || ((hi >> 52) & 0x7ff) != eb || (lo >> 63)) return 0;
    unsigned long long C = 0;
#pragma unroll
    for (int j = 0; j < L; ++j) {
        const unsigned long long b = (unsigned long long)__double_as_longlong(a[j]);
        if ((b << 1) == 0) continue;                       // +-0
        const unsigned long long ea = (b >> 52) & 0x7ff;
        if ((b >> 63) || ea == 0 || ea == 0x7ff || ea > eb) return 0;   // negative, subnormal, inf/nan, too large
        const unsigned long long ma = (b & M52) | (1ull << 52);
        const unsigned sh = (unsigned)(eb - ea);
        if (sh == 0) C += ma;
        else if (sh <= 53) {
            const unsigned long long half = 1ull << (sh - 1);
            const unsigned long long rem = ma & ((half << 1) - 1);
            if (rem == half) return 0;                     // tie: increment depends on the parity of s
            C += (ma >> sh) + (rem > half ? 1ull : 0ull);
        }                                                  // sh >= 54: a < ulp/2, increment 0
    }
    if (C >= B53) return 0;
    return (eb << 53) | C;
}

// combine 32 lane summaries of one warp: collapsible iff all are and all share the exponent
__device__ __forceinline__ unsigned long long warp_combine(unsigned long long w) {
    const unsigned long long e0 = __shfl_sync(0xffffffffu, w, 0) >> 53;
    const bool ok = __all_sync(0xffffffffu, w != 0 && (w >> 53) == e0);
    unsigned long long c = w & (B53 - 1);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    return (ok && c < B53) ? ((e0 << 53) | c) : 0ull;
}

__device__ __forceinline__ bool try_apply(double& s, unsigned long long w) {
    if (w == 0) return false;
    const unsigned long long b = (unsigned long long)__double_as_longlong(s);
    if ((b >> 52) != (w >> 53)) return false;              // sign bit set or different binade
    const unsigned long long tot = ((b & M52) | (1ull << 52)) + (w & (B53 - 1));
    if (tot >= B53) return false;                          // would leave the binade
    s = __longlong_as_double((long long)(((w >> 53) << 52) | (tot & M52)));
    return true;
}

struct Smem {
    double carry[4];
    double wtot[4][32];
    unsigned long long wsum[4][32];
    unsigned long long csum[4][THREADS];
};

// buf: [NR][TILE] doubles of dynamic shared memory; sm: scratch above.  All THREADS threads call.
// load(r, i) returns element i of chain r (only called for i < len).  out[r] = sequential sum.
template <int NR, typename Loader>
__device__ void block_exact_seq_sum(double (*buf)[TILE], Smem* sm, int len, Loader load, double* out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < NR) sm->carry[tid] = 0.0;
    __syncthreads();
    for (int base = 0; base < len; base += TILE) {
        long long t0_ = clock64();
        double a[NR][L], ls[NR], incl[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) {
#pragma unroll
            for (int j = 0; j < L; ++j) {
                const int i = base + tid * L + j;
                a[r][j] = (i < len) ? load(r, i) : 0.0;
                buf[r][tid * L + j] = a[r][j];
            }
            ls[r] = ((a[r][0] + a[r][1]) + a[r][2]) + a[r][3];
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
        if (warp < NR) {   // warp r turns the 32 warp totals of chain r into exclusive offsets
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
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const double pstart = sm->carry[r] + (sm->wtot[r][warp] + (incl[r] - ls[r]));
            const unsigned long long cs = chunk_summary(a[r], pstart, pstart + ls[r]);
            sm->csum[r][tid] = cs;
            const unsigned long long ws = warp_combine(cs);
            if (lane == 0) sm->wsum[r][warp] = ws;
        }
        __syncthreads();
        long long t1_ = clock64();
        if (warp < NR) {   // warp r walks chain r through this tile
            const int r = warp;
            const unsigned long long tsum = warp_combine(sm->wsum[r][lane]);
            if (lane == 0) {
                double s = sm->carry[r];
                if (!try_apply(s, tsum)) {
                    for (int w = 0; w < 32; ++w) {
                        if (try_apply(s, sm->wsum[r][w])) continue;
                        for (int t = w * 32; t < w * 32 + 32; ++t) {
                            if (try_apply(s, sm->csum[r][t])) continue;
                            const double* e = &buf[r][t * L];
                            s += e[0]; s += e[1]; s += e[2]; s += e[3];
                        }
                    }
                }
                sm->carry[r] = s;
            }
        }
        __syncthreads();
#ifdef FNN_XSUM_TIMING
        if (tid == 0) printf("tile base=%d prep=%lld walk=%lld cycles\n", base, t1_ - t0_, clock64() - t1_);
#endif
    }
    if (tid < NR) out[tid] = sm->carry[tid];
    __syncthreads();
}

}  // namespace xsum
