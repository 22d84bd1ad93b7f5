// fnn_host.cu — host-side pieces of libfastnn.so: error state, the one-shot B1 seam (its Phylip input goes through
// csrc/fnn_phylip.cpp), and the
// device generator for the synthetic additive-tree metrics.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <mutex>
#include "fastnn.h"
#include "fnn_common.h"

namespace fnn {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }
}  // namespace fnn

// The one-shot seam keeps its last context (device matrix, node tables, tensor map, instantiated CUDA graph) so that a
// host which orders several matrices of one size - the reference CLI in a loop, a bootstrap - pays cudaMalloc, the
// tensor-map encode and the graph instantiation once.  fnn_release_cache() gives the memory back.
namespace {
std::mutex g_cache_mu;
fnn_ctx* g_cache = nullptr;
fnn_opts g_cache_opts;
int64_t g_cache_n = 0;
}  // namespace

extern "C" void fnn_release_cache(void) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    if (g_cache) { fnn_ctx_destroy(g_cache); g_cache = nullptr; }
}

extern "C" int fnn_order(const fnn_opts* o, const double* Dh, const char* phylip_path, int64_t n, int32_t* ordering) {
    if (!ordering || n < 1 || ((Dh == nullptr) == (phylip_path == nullptr))) {
        fnn::set_error("fnn_order: need n>=1, ordering_out, and exactly one of D_rowmajor / phylip_path");
        return FNN_E_ARG;
    }
    if (n <= 3) {  // NetMakerOriginal.java:133-140
        for (int64_t i = 0; i <= n; ++i) ordering[i] = (int32_t)i;
        return FNN_OK;
    }
    std::vector<double> file_D;
    if (phylip_path) {   // native loader, csrc/fnn_phylip.cpp
        file_D.resize((size_t)n * n);
        int rc = fnn_read_phylip(phylip_path, n, file_D.data(), nullptr, 0, 0);
        if (rc) return rc;
        Dh = file_D.data();
    }
    fnn_opts key;
    if (o) key = *o; else fnn_default_opts(&key);
    std::lock_guard<std::mutex> lk(g_cache_mu);
    if (g_cache && (g_cache_n != n || memcmp(&g_cache_opts, &key, sizeof(key)) != 0)) { fnn_ctx_destroy(g_cache); g_cache = nullptr; }
    if (!g_cache) {
        int rc = fnn_ctx_create(&key, n, &g_cache);
        if (rc) { g_cache = nullptr; return rc; }
        g_cache_opts = key;
        g_cache_n = n;
    }
    int rc = fnn_ctx_load_host(g_cache, Dh);
    if (!rc) rc = fnn_ctx_order(g_cache, ordering);
    if (rc) { fnn_ctx_destroy(g_cache); g_cache = nullptr; }   // do not reuse a context whose run failed
    return rc;
}

// ---- device generator ---------------------------------------------------------------------
namespace {
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    unsigned long long z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// sparse table: tab[l*n + i] = max(h[i .. i+2^l-1])
__global__ void k_synth(double* D, int64_t ld, int n, const double* __restrict__ tab, const double* __restrict__ a,
                        const int* __restrict__ slot, unsigned long long noise_base, double eps) {
    const int64_t total = (int64_t)n * n;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int t1 = (int)(e / n), t2 = (int)(e % n);
        double v = 0.0;
        if (t1 != t2) {
            const int i = slot[t1], j = slot[t2];
            const int lo = min(i, j), hi = max(i, j);
            const int len = hi - lo;                // h indices lo .. hi-1
            const int l = 31 - __clz(len);
            const double mx = fmax(tab[(int64_t)l * n + lo], tab[(int64_t)l * n + hi - (1 << l)]);
            v = (2.0 * mx + a[lo]) + a[hi];
            if (eps != 0.0) {
                const unsigned long long tl = (unsigned long long)min(t1, t2), th = (unsigned long long)max(t1, t2);
                const unsigned long long key = noise_base + tl * (unsigned long long)n + th;
                const double u = (double)(splitmix64(key) >> 11) * (1.0 / 9007199254740992.0);
                v = v * (1.0 + eps * (2.0 * u - 1.0));
            }
        }
        D[(int64_t)t1 * ld + t2] = v;
    }
}
}  // namespace

extern "C" int fnn_ctx_synth(fnn_ctx* c, const double* h, const double* a, const int64_t* slot_of_taxon,
                             uint64_t noise_base, double eps) {
    if (!c || !h || !a || !slot_of_taxon) { fnn::set_error("fnn_ctx_synth: null argument"); return FNN_E_ARG; }
    FNN_CUDA(cudaSetDevice(fnn_ctx_device_(c)));   // a process may hold contexts on several devices
    cudaStream_t stream = fnn_ctx_stream_(c);
    double* dD; int64_t ld;
    fnn_ctx_matrix_ptr(c, &dD, &ld);
    const int64_t n = fnn_ctx_n_(c);
    int levels = 1;
    while ((1ll << levels) <= n) ++levels;
    std::vector<double> tab((size_t)levels * n, 0.0);
    for (int64_t i = 0; i + 1 < n; ++i) tab[i] = h[i];
    for (int l = 1; l < levels; ++l)
        for (int64_t i = 0; i + (1ll << l) <= n - 1; ++i)
            tab[(size_t)l * n + i] = std::max(tab[(size_t)(l - 1) * n + i], tab[(size_t)(l - 1) * n + i + (1ll << (l - 1))]);
    std::vector<int> slot(n);
    for (int64_t i = 0; i < n; ++i) slot[i] = (int)slot_of_taxon[i];
    double *d_tab = nullptr, *d_a = nullptr;
    int* d_slot = nullptr;
    DevScratch scratch;   // freed on every exit path
    FNN_CUDA(scratch.alloc((void**)&d_tab, tab.size() * sizeof(double)));
    FNN_CUDA(scratch.alloc((void**)&d_a, n * sizeof(double)));
    FNN_CUDA(scratch.alloc((void**)&d_slot, n * sizeof(int)));
    FNN_CUDA(cudaMemcpyAsync(d_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
    FNN_CUDA(cudaMemcpyAsync(d_a, a, n * sizeof(double), cudaMemcpyHostToDevice, stream));
    FNN_CUDA(cudaMemcpyAsync(d_slot, slot.data(), n * sizeof(int), cudaMemcpyHostToDevice, stream));
    k_synth<<<fnn_ctx_sms_(c) * 8, 256, 0, stream>>>(dD, ld, (int)n, d_tab, d_a, d_slot, noise_base, eps);
    FNN_CUDA(cudaGetLastError());
    FNN_CUDA(cudaStreamSynchronize(stream));
    fnn_ctx_mark_loaded_(c);
    return FNN_OK;
}
