// sort_trace.cu -- per-tile phase timestamps (globaltimer) of ONE onesweep pass, to find out what
// the decoupled look-back actually waits for.  Build with -DSORT_TRACE.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../../cudafluidsimulator_b200/csrc/sph_sort.cu"

int main() {
    const int n = 16000000;
    std::vector<uint32_t> h(n);
    srand(1);
    for (int i = 0; i < n; ++i) { int d = rand() % 600 - 300; long v = (long)((double)i / n * (1 << 24)) + d; h[i] = (uint32_t)std::min<long>(std::max<long>(v, 0), (1 << 24) - 1); }
    uint32_t *keys, *scratch; uint64_t *p0, *p1; unsigned long long *trace;
    const int tiles = sph::sort_tiles(n);
    cudaMalloc(&keys, n * 4); cudaMalloc(&p0, (size_t)n * 8); cudaMalloc(&p1, (size_t)n * 8);
    cudaMalloc(&scratch, sph::sort_scratch_words(n) * 4);
    cudaMalloc(&trace, (size_t)tiles * 8 * 8);
    cudaMemcpy(keys, h.data(), n * 4, cudaMemcpyHostToDevice);
    unsigned long long *null = nullptr;
    cudaMemcpyToSymbol(g_sort_trace, &null, sizeof(null));
    for (int it = 0; it < 3; ++it) sph::sort_pairs_async(keys, p0, p1, n, 3, scratch, 148, 0, nullptr);
    // trace a single-pass sort (pass 0 only: FIRST variant)
    cudaMemset(trace, 0, (size_t)tiles * 64);
    cudaMemcpyToSymbol(g_sort_trace, &trace, sizeof(trace));
    sph::sort_pairs_async(keys, p0, p1, n, 1, scratch, 148, 0, nullptr);
    cudaDeviceSynchronize();
    std::vector<unsigned long long> t((size_t)tiles * 8);
    cudaMemcpy(t.data(), trace, t.size() * 8, cudaMemcpyDeviceToHost);
    unsigned long long t0 = t[0];
    for (int i = 0; i < tiles; ++i) t0 = std::min(t0, t[(size_t)i * 8]);
    double sum[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < tiles; ++i) for (int k = 0; k < 5; ++k) sum[k] += (double)(t[(size_t)i * 8 + k + 1] - t[(size_t)i * 8 + k]);
    printf("tiles %d; mean ns: load %.0f rank %.0f publish %.0f lookback(+2 scans) %.0f scatter %.0f\n", tiles,
           sum[0] / tiles, sum[1] / tiles, sum[2] / tiles, sum[3] / tiles, sum[4] / tiles);
    // how late is the predecessor's publish relative to my own publish?
    double late = 0; int nlate = 0; double maxlate = 0;
    for (int i = 1; i < tiles; ++i) {
        double d = (double)t[(size_t)(i - 1) * 8 + 3] - (double)t[(size_t)i * 8 + 3];
        if (d > 0) { late += d; ++nlate; maxlate = std::max(maxlate, d); }
    }
    printf("predecessor published AFTER me in %d of %d tiles, mean lateness %.0f ns, max %.0f ns\n", nlate, tiles - 1, nlate ? late / nlate : 0.0, maxlate);
    printf("first 40 tiles: start publish looked end (us rel.) sm\n");
    for (int i = 0; i < 40; ++i)
        printf("  %4d %8.2f %8.2f %8.2f %8.2f  sm %llu\n", i, (t[(size_t)i*8] - t0) / 1e3, (t[(size_t)i*8+3] - t0) / 1e3, (t[(size_t)i*8+4] - t0) / 1e3, (t[(size_t)i*8+5] - t0) / 1e3, t[(size_t)i*8+6]);
    for (int i = 1000; i < 1020; ++i)
        printf("  %4d %8.2f %8.2f %8.2f %8.2f  sm %llu\n", i, (t[(size_t)i*8] - t0) / 1e3, (t[(size_t)i*8+3] - t0) / 1e3, (t[(size_t)i*8+4] - t0) / 1e3, (t[(size_t)i*8+5] - t0) / 1e3, t[(size_t)i*8+6]);
    return 0;
}
