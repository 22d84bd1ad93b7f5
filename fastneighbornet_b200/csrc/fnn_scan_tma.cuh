// fnn_scan_tma.cuh — K2, the selection scan, as a TMA-fed shared-memory pipeline (sm_100a).
//
// One persistent CTA per SM: a producer warp issues 2-D tiled TMA loads (cp.async.bulk.tensor,
// SASS UTMALDG) of 8-row x 256-column boxes of the distance matrix into a 6-stage ring
// (6 x 32 KB), completion signalled through mbarriers; eight consumer warps evaluate
// Q = (c-2)*Dpq - Sx[p] - Sx[q] on the staged rows and keep a per-thread (Q, i, j) min-loc.
// Memory-level parallelism comes from the ring (up to 192 KB in flight per SM), not from
// registers or occupancy.  Included by fnn_order.cu (needs DevState, Partial, better()).
#pragma once
#include <cuda.h>

namespace tma {

constexpr int CONSUMERS = 256;              // 8 consumer warps
constexpr int THREADS = CONSUMERS + 32;     // + 1 producer warp
constexpr int BOX_W = 256;                  // TMA box: 256 columns (2 KB) ...
constexpr int BOX_R = 8;                    // ... x 8 rows
constexpr int TILE_COLS = 2 * BOX_W;        // 512 columns = 2 boxes per stage
constexpr int TILE_ROWS = 32;               // rows per tile = 4 chunks
constexpr int STAGES = 6;
constexpr int STAGE_BYTES = 2 * BOX_R * BOX_W * 8;   // 32 KB
constexpr int KPB = TILE_COLS / TILE_ROWS;  // row tiles per 512-row band
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}

struct TileIter {   // identical tile sequence for producer and consumers
    long long t, total;
    int stride;
    __device__ TileIter(int m, int first, int stride_) : t(first), stride(stride_) {
        const int nRowTiles = (m + TILE_ROWS - 1) / TILE_ROWS;
        const int gFull = nRowTiles / KPB, rRem = nRowTiles % KPB;
        total = (long long)KPB * gFull * (gFull + 1) / 2 + (long long)rRem * (gFull + 1);
    }
    __device__ bool valid() const { return t < total; }
    __device__ void next() { t += stride; }
    __device__ void decode(int& r0, int& cb0) const {
        long long g = (long long)((sqrt(8.0 * (double)t / KPB + 1.0) - 1.0) * 0.5);
        while ((long long)KPB * g * (g + 1) / 2 > t) --g;
        while ((long long)KPB * (g + 1) * (g + 2) / 2 <= t) ++g;
        const long long rem = t - (long long)KPB * g * (g + 1) / 2;
        r0 = ((int)g * KPB + (int)(rem / (g + 1))) * TILE_ROWS;
        cb0 = (int)(rem % (g + 1)) * TILE_COLS;
    }
};

struct __align__(16) RowData { double S; long long pos; };   // one LDS.128 per row

#define FNN_CONSIDER(QV, RP, CP, ROWP)                                                                   \
    if ((QV) <= bq) {                                                                                    \
        const unsigned long long key_ = (ROWP) ? (((unsigned long long)(RP) << 32) | (unsigned)(CP))     \
                                               : (((unsigned long long)(CP) << 32) | (unsigned)(RP));    \
        if (better((QV), key_, bq, bk)) { bq = (QV); bk = key_; }                                        \
    }

__global__ void __launch_bounds__(THREADS, 1)
k_scan_tma(const __grid_constant__ CUtensorMap tmap, const double* __restrict__ Sx, const int* __restrict__ pos,
           DevState* st, Partial* partials) {
    if (st->done) return;
    extern __shared__ __align__(1024) unsigned char smem[];
    double* ring = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES];
    __shared__ RowData rowdata[2][TILE_ROWS];
    __shared__ Partial wbest[THREADS / 32];
    __shared__ bool amLast;

    const int m = st->m, P2 = st->P2;
    const double cm2 = (double)st->c - 2.0;
    const int tid = threadIdx.x;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    double bq = INFINITY;
    unsigned long long bk = ~0ull;

    if (tid >= CONSUMERS) {
        // ===================== producer warp: one lane issues the TMA loads =====================
        if (tid == CONSUMERS) {
            int stage = 0;
            uint32_t phase = 0;
            for (TileIter it(m, blockIdx.x, gridDim.x); it.valid(); it.next()) {
                int r0, cb0;
                it.decode(r0, cb0);
                const int rEnd = min(r0 + TILE_ROWS, m);
                for (int rc = r0; rc < rEnd; rc += BOX_R) {
                    const int rcEnd = min(rc + BOX_R, rEnd);
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    const bool second = (cb0 + BOX_W < rcEnd - 1) && (cb0 + BOX_W < m);
                    mbar_expect_tx(&full_bar[stage], second ? STAGE_BYTES : STAGE_BYTES / 2);
                    double* dst = ring + (size_t)stage * (STAGE_BYTES / 8);
                    tma_load_2d(dst, &tmap, cb0, rc, &full_bar[stage]);
                    if (second) tma_load_2d(dst + BOX_R * BOX_W, &tmap, cb0 + BOX_W, rc, &full_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===================== consumer warps =====================
        int stage = 0;
        uint32_t phase = 0;
        const int box = tid >> 7, lc = (tid & 127) * 2;   // which box of the stage, local column
        int tileParity = 0;
        for (TileIter it(m, blockIdx.x, gridDim.x); it.valid(); it.next(), tileParity ^= 1) {
            int r0, cb0;
            it.decode(r0, cb0);
            const int rEnd = min(r0 + TILE_ROWS, m);
            if (tid < TILE_ROWS && r0 + tid < m) {
                rowdata[tileParity][tid].S = Sx[r0 + tid];
                rowdata[tileParity][tid].pos = pos[r0 + tid];
            }
            const int c0 = cb0 + box * BOX_W + lc;
            const bool cvalid = (c0 < m) && (c0 < rEnd - 1);   // some row of the tile lies strictly below column c0
            const bool colPair = c0 < P2;
            double cS0 = 0.0, cS1 = 0.0;
            int cP0 = 0, cP1 = 0;
            if (cvalid) {
                cS0 = Sx[c0]; cP0 = pos[c0];
                if (c0 + 1 < m) { cS1 = Sx[c0 + 1]; cP1 = pos[c0 + 1]; }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory");
            const RowData* rd = rowdata[tileParity];

            for (int rc = r0; rc < rEnd; rc += BOX_R) {
                const int rcEnd = min(rc + BOX_R, rEnd);
                mbar_wait(&full_bar[stage], phase);
                const double* sm = ring + (size_t)stage * (STAGE_BYTES / 8) + box * (BOX_R * BOX_W) + lc;
                if (cvalid && c0 < rcEnd - 1) {
                    const bool full_rows = (rcEnd - rc == BOX_R);
                    if (colPair) {
                        if (full_rows && rcEnd <= P2 && c0 < rc) {
                            // ---- pair rows x pair column, interior
#pragma unroll
                            for (int u = 0; u < BOX_R; u += 2) {
                                const double2 e0 = *reinterpret_cast<const double2*>(sm + u * BOX_W);
                                const double2 e1 = *reinterpret_cast<const double2*>(sm + (u + 1) * BOX_W);
                                const RowData r = rd[rc - r0 + u];
                                const int rp = (int)r.pos;
                                const bool rowP = rp > cP0;
                                const double t1 = rowP ? e0.y : e1.x;
                                const double t2 = rowP ? e1.x : e0.y;
                                const double dpq = (((e0.x + t1) + t2) + e1.y) * 0.25;
                                const double s1 = rowP ? r.S : cS0;
                                const double s2 = rowP ? cS0 : r.S;
                                const double q = (cm2 * dpq - s1) - s2;
                                FNN_CONSIDER(q, rp, cP0, rowP)
                            }
                        } else if (full_rows && rc >= P2) {
                            // ---- single rows x pair column (always below the diagonal)
#pragma unroll
                            for (int u = 0; u < BOX_R; ++u) {
                                const double2 e = *reinterpret_cast<const double2*>(sm + u * BOX_W);
                                const RowData r = rd[rc - r0 + u];
                                const int rp = (int)r.pos;
                                const bool rowP = rp > cP0;
                                const double dpq = (e.x + e.y) * 0.5;
                                const double s1 = rowP ? r.S : cS0;
                                const double s2 = rowP ? cS0 : r.S;
                                const double q = (cm2 * dpq - s1) - s2;
                                FNN_CONSIDER(q, rp, cP0, rowP)
                            }
                        } else {
                            // ---- generic pair column (diagonal / region boundary / ragged end)
                            for (int rr = rc; rr < rcEnd;) {
                                const RowData r = rd[rr - r0];
                                const int rp = (int)r.pos;
                                const bool rowP = rp > cP0;
                                const double s1 = rowP ? r.S : cS0;
                                const double s2 = rowP ? cS0 : r.S;
                                if (rr < P2) {
                                    if (c0 < rr) {
                                        const double2 e0 = *reinterpret_cast<const double2*>(sm + (rr - rc) * BOX_W);
                                        const double2 e1 = *reinterpret_cast<const double2*>(sm + (rr - rc + 1) * BOX_W);
                                        const double t1 = rowP ? e0.y : e1.x;
                                        const double t2 = rowP ? e1.x : e0.y;
                                        const double dpq = (((e0.x + t1) + t2) + e1.y) * 0.25;
                                        const double q = (cm2 * dpq - s1) - s2;
                                        FNN_CONSIDER(q, rp, cP0, rowP)
                                    }
                                    rr += 2;
                                } else {
                                    const double2 e = *reinterpret_cast<const double2*>(sm + (rr - rc) * BOX_W);
                                    const double dpq = (e.x + e.y) * 0.5;
                                    const double q = (cm2 * dpq - s1) - s2;
                                    FNN_CONSIDER(q, rp, cP0, rowP)
                                    rr += 1;
                                }
                            }
                        }
                    } else {
                        if (full_rows && rc >= P2 && c0 + 1 < rc && c0 + 1 < m) {
                            // ---- single rows x two single columns, interior
#pragma unroll
                            for (int u = 0; u < BOX_R; ++u) {
                                const double2 e = *reinterpret_cast<const double2*>(sm + u * BOX_W);
                                const RowData r = rd[rc - r0 + u];
                                const int rp = (int)r.pos;
                                {
                                    const bool rowP = rp > cP0;
                                    const double s1 = rowP ? r.S : cS0;
                                    const double s2 = rowP ? cS0 : r.S;
                                    const double q = (cm2 * e.x - s1) - s2;
                                    FNN_CONSIDER(q, rp, cP0, rowP)
                                }
                                {
                                    const bool rowP = rp > cP1;
                                    const double s1 = rowP ? r.S : cS1;
                                    const double s2 = rowP ? cS1 : r.S;
                                    const double q = (cm2 * e.y - s1) - s2;
                                    FNN_CONSIDER(q, rp, cP1, rowP)
                                }
                            }
                        } else {
                            // ---- generic single columns
                            for (int rr = max(max(rc, P2), c0 + 1); rr < rcEnd; ++rr) {
                                const double2 e = *reinterpret_cast<const double2*>(sm + (rr - rc) * BOX_W);
                                const RowData r = rd[rr - r0];
                                const int rp = (int)r.pos;
                                {
                                    const bool rowP = rp > cP0;
                                    const double s1 = rowP ? r.S : cS0;
                                    const double s2 = rowP ? cS0 : r.S;
                                    const double q = (cm2 * e.x - s1) - s2;
                                    FNN_CONSIDER(q, rp, cP0, rowP)
                                }
                                if (c0 + 1 < rr) {
                                    const bool rowP = rp > cP1;
                                    const double s1 = rowP ? r.S : cS1;
                                    const double s2 = rowP ? cS1 : r.S;
                                    const double q = (cm2 * e.y - s1) - s2;
                                    FNN_CONSIDER(q, rp, cP1, rowP)
                                }
                            }
                        }
                    }
                }
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(&empty_bar[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    }

    // ---- block min-loc, then the last block to finish reduces the per-block partials
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double oq = __shfl_down_sync(0xffffffffu, bq, off);
        unsigned long long ok = __shfl_down_sync(0xffffffffu, bk, off);
        if (better(oq, ok, bq, bk)) { bq = oq; bk = ok; }
    }
    if ((tid & 31) == 0) wbest[tid >> 5] = Partial{bq, bk};
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < THREADS / 32; ++w)
            if (better(wbest[w].q, wbest[w].key, bq, bk)) { bq = wbest[w].q; bk = wbest[w].key; }
        partials[blockIdx.x] = Partial{bq, bk};
        __threadfence();
        unsigned int tk = atomicAdd(&st->ticket, 1u);
        amLast = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (amLast) {
        __threadfence();
        bq = INFINITY; bk = ~0ull;
        for (int b = tid; b < (int)gridDim.x; b += THREADS) {
            const double pq = __ldcg(&partials[b].q);
            const unsigned long long pk = __ldcg(&partials[b].key);
            if (better(pq, pk, bq, bk)) { bq = pq; bk = pk; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            double oq = __shfl_down_sync(0xffffffffu, bq, off);
            unsigned long long ok = __shfl_down_sync(0xffffffffu, bk, off);
            if (better(oq, ok, bq, bk)) { bq = oq; bk = ok; }
        }
        __syncthreads();
        if ((tid & 31) == 0) wbest[tid >> 5] = Partial{bq, bk};
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < THREADS / 32; ++w)
                if (better(wbest[w].q, wbest[w].key, bq, bk)) { bq = wbest[w].q; bk = wbest[w].key; }
            st->selQ = bq;
            st->sel_i = (int)(bk >> 32);
            st->sel_j = (int)(bk & 0xffffffffu);
            st->ticket = 0;
        }
    }
}

}  // namespace tma
