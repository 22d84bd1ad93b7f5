// fnn_host.cu — host-side pieces of libfastnn.so: error state, the one-shot B1 seam, the
// native Phylip reader (conventions of DistancesAndNames.java:43-132), and the
// device generator for the synthetic additive-tree metrics.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
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

// ---- Phylip: line 1 = n (all whitespace stripped, FastNN.java:272-274); then `name v v v ...`
// split on single spaces, then tabs; only columns < row are consumed (DistancesAndNames.java:65-87),
// so square and lower-triangular files both load.
static int read_phylip(const char* path, int64_t n, std::vector<double>& D) {
    FILE* f = fopen(path, "rb");
    if (!f) { fnn::set_error("cannot open %s", path); return FNN_E_IO; }
    std::string line;
    auto getline_ = [&](std::string& out) -> bool {
        out.clear();
        int ch;
        bool any = false;
        while ((ch = fgetc(f)) != EOF) {
            any = true;
            if (ch == '\n') break;
            out.push_back((char)ch);
        }
        if (!out.empty() && out.back() == '\r') out.pop_back();
        return any;
    };
    if (!getline_(line)) { fclose(f); fnn::set_error("%s: empty file", path); return FNN_E_IO; }
    std::string digits;
    for (char ch : line) if (!isspace((unsigned char)ch)) digits.push_back(ch);
    const long long n_file = atoll(digits.c_str());
    if (n_file != n) { fclose(f); fnn::set_error("%s: header says %lld taxa, caller says %lld", path, n_file, (long long)n); return FNN_E_ARG; }
    D.assign((size_t)n * n, 0.0);
    int64_t row = 0;
    std::vector<const char*> toks;
    while (row < n && getline_(line)) {
        // tokens: split on ' ' then '\t'; first token is the name
        size_t p = 0;
        bool first = true;
        int64_t col = 0;
        bool empty_name = false;
        while (p <= line.size()) {
            size_t e = line.find(' ', p);
            if (e == std::string::npos) e = line.size();
            if (first) {
                first = false;
                if (e == p && line.empty()) empty_name = true;
            } else if (e > p) {
                size_t q = p;
                while (q < e) {
                    size_t te = line.find('\t', q);
                    if (te == std::string::npos || te > e) te = e;
                    if (te > q && col < row) {
                        const double v = strtod(line.substr(q, te - q).c_str(), nullptr);
                        D[(size_t)row * n + col] = v;
                        D[(size_t)col * n + row] = v;
                        ++col;
                    }
                    q = te + 1;
                }
            }
            p = e + 1;
        }
        if (empty_name) break;
        if (col < row) { fclose(f); fnn::set_error("%s: row %lld has %lld of %lld lower-triangle values", path, (long long)row, (long long)col, (long long)row); return FNN_E_IO; }
        ++row;
    }
    fclose(f);
    if (row < n) { fnn::set_error("%s: %lld rows, expected %lld", path, (long long)row, (long long)n); return FNN_E_IO; }
    return FNN_OK;
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
    if (phylip_path) {
        int rc = read_phylip(phylip_path, n, file_D);
        if (rc) return rc;
        Dh = file_D.data();
    }
    fnn_ctx* c = nullptr;
    int rc = fnn_ctx_create(o, n, &c);
    if (rc) return rc;
    rc = fnn_ctx_load_host(c, Dh);
    if (!rc) rc = fnn_ctx_order(c, ordering);
    fnn_ctx_destroy(c);
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
    FNN_CUDA(cudaMemcpy(d_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    FNN_CUDA(cudaMemcpy(d_a, a, n * sizeof(double), cudaMemcpyHostToDevice));
    FNN_CUDA(cudaMemcpy(d_slot, slot.data(), n * sizeof(int), cudaMemcpyHostToDevice));
    k_synth<<<148 * 8, 256>>>(dD, ld, (int)n, d_tab, d_a, d_slot, noise_base, eps);
    FNN_CUDA(cudaGetLastError());
    FNN_CUDA(cudaDeviceSynchronize());
    fnn_ctx_mark_loaded_(c);
    return FNN_OK;
}
