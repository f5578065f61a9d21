// sph_sort.cuh -- interface of the hand-written onesweep radix sort (sph_sort.cu).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace sph {

constexpr int kRadix = 256;                             // 8-bit digits
constexpr int kSortThreads = 256;                       // == kRadix: thread d owns digit d
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 16;                          // pairs per thread
constexpr int kSortTile = kSortThreads * kSortItems;    // 4096 pairs per CTA
constexpr int kMaxPasses = 4;                           // 32-bit keys

enum { kSortStageHistogram = 0, kSortStagePass = 1 };

// Optional per-kernel instrumentation (CUDA events) supplied by the step driver.
struct SortHooks {
    void *ctx;
    void (*before)(void *ctx, int sort_stage);
    void (*after)(void *ctx, int sort_stage);
};

int sort_tiles(int n);
size_t sort_scratch_words(int capacity);  // uint32 words of scratch for up to `capacity` pairs
int sort_passes_for(uint32_t table_size); // 8-bit passes needed for keys < table_size

// Sorts (keys[i], i) for i in [0, n) by key, stably.  Pair = key << 32 | index.
// Enqueues a scratch clear, the histogram kernel and `passes` onesweep kernels on
// `stream`; returns which of pairs0 / pairs1 (0 / 1) receives the sorted pairs.
// n_dev != nullptr: the count is read from device memory by the kernels and `n` is only its
// upper bound (grid and scratch sizes) -- the slab cluster step never brings counts to the host.
int sort_pairs_async(const uint32_t *keys, uint64_t *pairs0, uint64_t *pairs1, int n, int passes,
                     uint32_t *scratch, int sm_count, cudaStream_t stream, SortHooks *hooks,
                     const int *n_dev = nullptr);

// Counting sort by cell for keys < table_entries - 1 (single-GPU step; see sph_sort.cu).  Enqueues the
// count, scan and scatter kernels: cell_start[k] (k < table_entries) = number of keys below k, and
// pairs_sorted = the (key, index) pairs grouped by key, members of one cell in arbitrary order --
// the reorder kernel ranks them by index.  `count` (table_entries words) must be zero on entry and
// is zero again on exit.  counted: count[] and pairs_tmp (key << 32 | provisional rank) were already
// produced (by the force kernel of the previous step), skip the count kernel.  n_dev: as for
// sort_pairs_async.  base: added to every cell_start entry (slot of the first sorted particle).  `scratch`: cell_sort_scratch_words(table_entries) words.
size_t cell_sort_scratch_words(uint32_t table_entries);
void cell_sort_async(const uint32_t *keys, uint64_t *pairs_sorted, uint64_t *pairs_tmp, int n,
                     uint32_t table_entries, uint32_t *count, uint32_t *cell_start, uint32_t *scratch,
                     cudaStream_t stream, SortHooks *hooks, bool counted, const int *n_dev = nullptr,
                     uint32_t base = 0);

}  // namespace sph
