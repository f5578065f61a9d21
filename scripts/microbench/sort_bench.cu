// sort_bench.cu -- phase-elimination timing of the onesweep pass (build with -DSORT_EXP=mask:
// 1 no look-back, 2 no ranking, 4 linear output).  Results with SORT_EXP != 0 are wrong by
// construction; only the time matters.  Keys mimic a settled SPH state: nearly sorted
// 24-bit flat cell keys.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../../cudafluidsimulator_b200/csrc/sph_sort.cu"

int main(int argc, char **argv) {
    const int n = 16000000;
    const bool shuffled = argc > 1;
    std::vector<uint32_t> h(n);
    srand(1);
    for (int i = 0; i < n; ++i) h[i] = (uint32_t)((double)i / n * (1 << 24));
    if (shuffled) for (int i = 0; i < n; ++i) h[i] = (uint32_t)(((uint64_t)rand() * 65536 + rand()) & 0xffffff);
    else for (int i = 0; i < n; ++i) { int d = rand() % 600 - 300; long v = (long)h[i] + d; h[i] = (uint32_t)std::min<long>(std::max<long>(v, 0), (1 << 24) - 1); }
    uint32_t *keys, *scratch; uint64_t *p0, *p1;
    cudaMalloc(&keys, n * 4); cudaMalloc(&p0, (size_t)n * 8); cudaMalloc(&p1, (size_t)n * 8);
    cudaMalloc(&scratch, sph::sort_scratch_words(n) * 4);
    cudaMemcpy(keys, h.data(), n * 4, cudaMemcpyHostToDevice);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int it = 0; it < 3; ++it) sph::sort_pairs_async(keys, p0, p1, n, 3, scratch, 148, 0, nullptr);
    cudaEventRecord(a);
    const int reps = 20;
    int out = 0;
    for (int it = 0; it < reps; ++it) out = sph::sort_pairs_async(keys, p0, p1, n, 3, scratch, 148, 0, nullptr);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    std::vector<uint64_t> r(n);
    cudaMemcpy(r.data(), out ? p1 : p0, (size_t)n * 8, cudaMemcpyDeviceToHost);
    bool ok = true;
    for (int i = 1; i < n && ok; ++i) ok = r[i - 1] <= r[i];   // (key, index) pairs strictly ordered when stable
    printf("SORT_EXP=%d %s: %.3f ms per sort (hist + 3 passes), sorted+stable=%d, err=%s\n",
#ifdef SORT_EXP
           SORT_EXP,
#else
           0,
#endif
           shuffled ? "random" : "nearly-sorted", ms / reps, (int)ok, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
