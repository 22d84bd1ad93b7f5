// fnn_tile_iter.h - the tile sequence of the selection scan (csrc/fnn_scan_tma.cuh) and its multi-GPU partition.
//
// The scan walks the lower triangle in tiles of TILE_ROWS x TILE_COLS; tiles are numbered band by band (band g = 512 rows,
// KPB row tiles per band, g+1 column tiles each).  CTA b of rank r in a world of w starts at tile r + w*b and strides by
// w * gridDim: rank r owns tiles t = r (mod w).  Plain C++ with host/device qualifiers: the kernels use it on the device,
// and the CPU (gloo) tests of the N>1 path drive THIS code through the oracle library (oracle_tile_*), not a mirror.
#pragma once
#include <math.h>
#if defined(__CUDACC__)
#define FNN_TI_HD __host__ __device__
#else
#define FNN_TI_HD
#endif

namespace tma {

constexpr int TILE_COLS = 512;              // 2 TMA boxes per stage
constexpr int TILE_ROWS = 32;               // rows per tile = 4 chunks of 8
constexpr int KPB = TILE_COLS / TILE_ROWS;  // row tiles per 512-row band

struct TileIter {   // identical tile sequence for producer and consumers (and for every rank)
    long long t, total;
    int stride;
    FNN_TI_HD TileIter(int m, int first, int stride_) : t(first), stride(stride_) {
        const int nRowTiles = (m + TILE_ROWS - 1) / TILE_ROWS;
        const int gFull = nRowTiles / KPB, rRem = nRowTiles % KPB;
        total = (long long)KPB * gFull * (gFull + 1) / 2 + (long long)rRem * (gFull + 1);
    }
    FNN_TI_HD bool valid() const { return t < total; }
    FNN_TI_HD void next() { t += stride; }
    FNN_TI_HD void decode(int& r0, int& cb0) const {
        long long g = (long long)((sqrt(8.0 * (double)t / KPB + 1.0) - 1.0) * 0.5);
        while ((long long)KPB * g * (g + 1) / 2 > t) --g;
        while ((long long)KPB * (g + 1) * (g + 2) / 2 <= t) ++g;
        const long long rem = t - (long long)KPB * g * (g + 1) / 2;
        r0 = ((int)g * KPB + (int)(rem / (g + 1))) * TILE_ROWS;
        cb0 = (int)(rem % (g + 1)) * TILE_COLS;
    }
};


}  // namespace tma
