// sph_sort.cu -- the step's sort of particles by cell key, hand-written for sm_100a.  Two algorithms
// with the same result, the order (key, position in the array that was sorted):
//   * counting sort by cell (default; second half of this file): per-cell counts -- taken by the
//     force kernel of the previous step --, one scan of the cell table, a scatter; the members of a
//     cell are ranked by index in the reorder kernel.  Every array is read and written once.
//   * stable LSD radix sort of (cell key, slot index) pairs: one histogram sweep + one "onesweep"
//     pass per 8-bit digit (chained scan with decoupled look-back, so every pass reads and writes
//     each pair exactly once).  SPH_SORT=radix / SphOptions.sort_algo; the A/B arm.
//
// Replaces the neighbour-search data structures of the three reference variants
// (ref: src/simulator.cu:44-55 insertList / 133-147 kernelBuildGrid for the
// lock-free lists; README.md:5 for index_sort and z_index_sort whose source is
// not mounted).  Stable => the order is deterministic.
//
// Roofline: HBM.  Algorithmic bytes per pair: radix: histogram 4 (keys read once) and
// per pass 8 read + 8 written (first pass reads only the 4-byte key, the index
// is implicit); counting sort: 32 per particle + 16 per cell (see below).
#include "sph_sort.cuh"
#include "sph_common.cuh"

// Tunables of the onesweep pass (see profiles/r01_sort_phase_elimination.txt): resident CTAs
// per SM and status words fetched per look-back step.  The look-back walk covers every
// running predecessor tile; it stays short only while
//   (resident tiles / tile time) x (latency of one step) / window  <<  1.
#ifndef SORT_CTAS_PER_SM
#define SORT_CTAS_PER_SM 4
#endif
#ifndef SORT_LOOKBACK_WINDOW
#define SORT_LOOKBACK_WINDOW 4
#endif
#ifndef SORT_USE_MATCH_INSTRUCTION
#define SORT_USE_MATCH_INSTRUCTION 0
#endif
#ifndef SORT_SPIN_SLEEP_NS
#define SORT_SPIN_SLEEP_NS 0
#endif

#ifdef SORT_TRACE
__device__ unsigned long long *g_sort_trace;   // [tile][8]: t_start, t_loaded, t_ranked, t_published, t_looked, t_end, smid, walk
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define TRACE(slot) do { if (threadIdx.x == 0 && g_sort_trace) g_sort_trace[(size_t)tile * 8 + (slot)] = gtime(); } while (0)
#else
#define TRACE(slot) do { } while (0)
#endif

namespace sph {

namespace {

constexpr uint32_t kFlagAggregate = 1u << 30;  // tile count published
constexpr uint32_t kFlagInclusive = 1u << 31;  // inclusive prefix published
constexpr uint32_t kValueMask = (1u << 30) - 1;

__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(uint32_t *p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Lanes of the warp holding the same 8-bit digit: eight ballots intersected.  (The
// MATCH.ANY instruction computes the same thing but at a small fraction of the ballot
// rate -- measured: the pass ran at IPC 0.9 with its warps parked on the match results.)
// One bit of the digit: peers &= lanes whose bit equals mine.  Test, vote, select and one 3-input
// logic op (4 SASS instructions); written in PTX because the compiler turns the C form into
// shift + and + compare + select + vote + logic (6).
template <int BIT>
__device__ __forceinline__ void ballot_step(uint32_t &peers, uint32_t d) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 t, b;\n\t"
        "and.b32 t, %1, %2;\n\t"
        "setp.ne.u32 p, t, 0;\n\t"
        "vote.sync.ballot.b32 b, p, 0xffffffff;\n\t"
        "selp.b32 t, 0, 0xffffffff, p;\n\t"   // a lane whose bit is clear matches the complement
        "xor.b32 b, b, t;\n\t"
        "and.b32 %0, %0, b;\n\t"
        "}"
        : "+r"(peers)
        : "r"(d), "n"(1u << BIT));
}

__device__ __forceinline__ uint32_t match_digit(uint32_t d) {
#if SORT_USE_MATCH_INSTRUCTION
    return __match_any_sync(0xffffffffu, d);
#endif
    uint32_t peers = 0xffffffffu;
    ballot_step<0>(peers, d); ballot_step<1>(peers, d); ballot_step<2>(peers, d); ballot_step<3>(peers, d);
    ballot_step<4>(peers, d); ballot_step<5>(peers, d); ballot_step<6>(peers, d); ballot_step<7>(peers, d);
    return peers;
}

// Exclusive scan over the 256 threads of a block; `total` = sum of all inputs.
__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t *s_warp /*8*/,
                                                             uint32_t &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t w = s_warp[lane & 7];
    uint32_t wincl = w;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, wincl, o);
        if ((lane & 7) >= o) wincl += t;
    }
    total = __shfl_sync(0xffffffffu, wincl, 7);
    const uint32_t warp_excl = __shfl_sync(0xffffffffu, wincl - w, warp);
    __syncthreads();  // s_warp may be reused by the caller
    return incl - v + warp_excl;
}

// ---- histogram of every digit in one sweep over the keys ----------------------
// Blocked arrangement (16 consecutive keys per thread) with per-thread run
// aggregation: the array being sorted is last step's sorted order, so the high
// digits of consecutive keys are almost always equal and would otherwise
// serialise on one shared-memory counter.
__global__ void __launch_bounds__(kSortThreads)
    k_histogram(const uint32_t *__restrict__ keys, int n, const int *__restrict__ n_dev, int passes,
                uint32_t *__restrict__ ghist) {
    __shared__ uint32_t s_hist[kMaxPasses * kRadix];
    if (n_dev) n = *n_dev;   // count known only on the device (slab cluster): the grid covers the capacity
    for (int i = threadIdx.x; i < kMaxPasses * kRadix; i += kSortThreads) s_hist[i] = 0;
    __syncthreads();

    const int stride = gridDim.x * kSortThreads * 16;
    for (int base = (blockIdx.x * kSortThreads + threadIdx.x) * 16; base < n; base += stride) {
        uint32_t k[16];
        if (base + 16 <= n) {
            const uint4 *src = reinterpret_cast<const uint4 *>(keys + base);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                uint4 q = __ldg(src + v);
                k[4 * v] = q.x; k[4 * v + 1] = q.y; k[4 * v + 2] = q.z; k[4 * v + 3] = q.w;
            }
        } else {
#pragma unroll
            for (int v = 0; v < 16; ++v) k[v] = (base + v < n) ? __ldg(keys + base + v) : 0xffffffffu;
        }
        const int valid = min(16, n - base);
        for (int p = 0; p < passes; ++p) {
            const int shift = 8 * p;
            uint32_t run_digit = (k[0] >> shift) & 255u;
            uint32_t run_len = 1;
#pragma unroll
            for (int v = 1; v < 16; ++v) {
                if (v < valid) {
                    const uint32_t d = (k[v] >> shift) & 255u;
                    if (d == run_digit) {
                        ++run_len;
                    } else {
                        atomicAdd(&s_hist[p * kRadix + run_digit], run_len);
                        run_digit = d;
                        run_len = 1;
                    }
                }
            }
            atomicAdd(&s_hist[p * kRadix + run_digit], run_len);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * kRadix; i += kSortThreads) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(&ghist[i], c);
    }
}

// ---- one onesweep pass ----------------------------------------------------------
// FIRST: input is the bare key array, the payload (slot index) is implicit.
// Tile = 4096 consecutive pairs, 8 warps x 512, warp-striped so that every load
// is a fully coalesced 256-byte (128-byte for FIRST) request and the order of
// (item, lane) inside a warp is the array order -- needed for stability.
template <bool FIRST>
__global__ void __launch_bounds__(kSortThreads, SORT_CTAS_PER_SM)
    k_onesweep(const uint32_t *__restrict__ keys_in, const uint64_t *__restrict__ pairs_in,
               uint64_t *__restrict__ pairs_out, int n, const int *__restrict__ n_dev, int shift,
               const uint32_t *__restrict__ ghist,  // 256 counts of this digit
               uint32_t *__restrict__ status,       // tiles x 256, zeroed
               uint32_t *__restrict__ ticket) {     // zeroed
    __shared__ uint64_t s_pairs[kSortTile];                 // 32 KB
    __shared__ uint32_t s_whist[kSortWarps][kRadix];        // 8 KB
    __shared__ uint32_t s_count[kRadix];
    __shared__ uint32_t s_dstart[kRadix];
    __shared__ uint32_t s_goff[kRadix];
    __shared__ uint32_t s_scan[8];
    __shared__ uint32_t s_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (n_dev) n = *n_dev;   // (see k_histogram) CTAs beyond the last tile leave right after their ticket
    // Tiles are handed out in launch order: a tile only ever waits on tiles with
    // a smaller ticket, which are already resident or finished.
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) s_whist[w][tid] = 0;
    s_count[tid] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    TRACE(0);
    const int base = (int)tile * kSortTile;
    if (base >= n) return;   // CTA-uniform; nobody looks back at a tile that does not exist
    const int tile_valid = min(kSortTile, n - base);

    // -- load -------------------------------------------------------------------
    uint64_t item[kSortItems];
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const int local = warp * (kSortItems * 32) + k * 32 + lane;
        const int idx = base + local;
        uint64_t v = ~0ull;  // padding sorts behind everything and is never written
        if (local < tile_valid) {
            if (FIRST) v = ((uint64_t)__ldg(keys_in + idx) << 32) | (uint32_t)idx;
            else v = __ldg(pairs_in + idx);
        }
        item[k] = v;
    }

    TRACE(1);
    // -- early counts: the tile's digit histogram, published BEFORE the (much slower) ranking.
    // Every later tile needs this aggregate for its look-back; tiles sharing an SM finish their
    // ranking at very different times (3.8 / 6.2 / 9.4 us for the 1st / 2nd / 3rd resident CTA,
    // profiles/r01_sort_trace.txt), and with the aggregate published only after ranking every
    // tile waited for the slowest running predecessor.
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const uint32_t d = (uint32_t)(item[k] >> (32 + shift)) & 255u;
        const uint32_t d0 = __shfl_sync(0xffffffffu, d, 0);
        if (__all_sync(0xffffffffu, d == d0)) {          // warp-uniform digit (nearly sorted input)
            if (lane == 0) atomicAdd(&s_count[d0], 32u);
        } else {
            atomicAdd(&s_count[d], 1u);
        }
    }
    __syncthreads();
    const int d = tid;  // kSortThreads == kRadix: thread d owns digit d
    const uint32_t count = s_count[d];
    uint32_t *my_status = status + (size_t)tile * kRadix + d;
    st_relaxed(my_status, (tile == 0 ? kFlagInclusive : kFlagAggregate) | count);
    TRACE(2);

    // -- stable rank of every item among equal digits of its warp ---------------
    // A batch of independent peer masks first, then the (serial, shared-memory) counter
    // updates of that batch.
    uint32_t rank[kSortItems];
#if defined(SORT_EXP) && (SORT_EXP & 2)   // microbenchmark only: no ranking (wrong result)
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) rank[k] = 0;
#else
    constexpr int kBatch = 8;
#pragma unroll
    for (int k0 = 0; k0 < kSortItems; k0 += kBatch) {
        uint32_t peers[kBatch];
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
            const uint32_t dd = (uint32_t)(item[k0 + j] >> (32 + shift)) & 255u;
            const uint32_t d0 = __shfl_sync(0xffffffffu, dd, 0);
            // warp-uniform digit: everyone is everyone's peer, no need for eight ballots
            peers[j] = __all_sync(0xffffffffu, dd == d0) ? 0xffffffffu : match_digit(dd);
        }
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
            const uint32_t dd = (uint32_t)(item[k0 + j] >> (32 + shift)) & 255u;
            const int leader = __ffs(peers[j]) - 1;
            const uint32_t below = __popc(peers[j] & ((1u << lane) - 1u));
            uint32_t old = 0;
            if (lane == leader) {
                old = s_whist[warp][dd];
                s_whist[warp][dd] = old + __popc(peers[j]);
            }
            old = __shfl_sync(0xffffffffu, old, leader);
            rank[k0 + j] = old + below;
            __syncwarp();
        }
    }
#endif
    __syncthreads();

    // -- per-digit: exclusive scan of the warp counters -----------------------------
    {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const uint32_t c = s_whist[w][d];
            s_whist[w][d] = run;  // exclusive over warps
            run += c;
        }
    }
    TRACE(3);
    uint32_t total;
    const uint32_t dstart = block_exclusive_scan_256(count, s_scan, total);
    const uint32_t gbase = block_exclusive_scan_256(__ldg(ghist + d), s_scan, total);

    // Decoupled look-back, a window of predecessors per step: the status words of
    // kWindow earlier tiles are fetched together (independent loads), then consumed in
    // order until an inclusive prefix closes the sum.  A tile whose word is not published
    // yet is polled again.  (All tiles of a wave start together, so a serial walk would pay
    // one L2 round trip per running predecessor -- hundreds -- before the first prefix.)
    uint32_t excl = 0;
#if defined(SORT_EXP) && (SORT_EXP & 1)   // microbenchmark only: no look-back (wrong result)
    if (false) {
#else
    if (tile > 0) {
#endif
        constexpr int kWindow = SORT_LOOKBACK_WINDOW;
        int prev = (int)tile - 1;
        bool closed = false;
        while (!closed) {
            uint32_t v[kWindow];
#pragma unroll
            for (int j = 0; j < kWindow; ++j)
                v[j] = (prev - j >= 0) ? ld_relaxed(status + (size_t)(prev - j) * kRadix + d) : kFlagInclusive;
            int used = 0;
#pragma unroll
            for (int j = 0; j < kWindow; ++j) {
                if (!closed && used == j) {
                    if (v[j] & (kFlagAggregate | kFlagInclusive)) {
                        excl += v[j] & kValueMask;
                        closed = (v[j] & kFlagInclusive) != 0;
                        ++used;
                    }
                }
            }
            prev -= used;
#if SORT_SPIN_SLEEP_NS > 0
            // Blocked on a tile that has not published yet: back off instead of hammering the
            // issue slots the publishing CTAs on this SM need.
            if (!closed && used == 0) __nanosleep(SORT_SPIN_SLEEP_NS);
#endif
        }
        st_relaxed(my_status, kFlagInclusive | (excl + count));
    }
    s_dstart[d] = dstart;
    s_goff[d] = gbase + excl - dstart;  // global slot = s_goff[digit] + slot in sorted tile
    __syncthreads();
    TRACE(4);
#ifdef SORT_TRACE
    if (tid == 0 && g_sort_trace) { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); g_sort_trace[(size_t)tile * 8 + 6] = sm; }
#endif

    // -- scatter into tile-sorted order in shared memory -------------------------
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const uint32_t dd = (uint32_t)(item[k] >> (32 + shift)) & 255u;
#if defined(SORT_EXP) && (SORT_EXP & 2)
        s_pairs[tid * kSortItems + k] = item[k];
#else
        s_pairs[s_dstart[dd] + s_whist[warp][dd] + rank[k]] = item[k];
#endif
    }
    __syncthreads();

    // -- coalesced runs out to global --------------------------------------------
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const int s = k * kSortThreads + tid;
        if (s < tile_valid) {
            const uint64_t v = s_pairs[s];
            const uint32_t dd = (uint32_t)(v >> (32 + shift)) & 255u;
#if defined(SORT_EXP) && (SORT_EXP & 4)   // microbenchmark only: linear output (wrong result)
            pairs_out[base + s] = v + dd + s_goff[dd];
#else
            pairs_out[s_goff[dd] + (uint32_t)s] = v;
#endif
        }
    }
    TRACE(5);
}


// ===== counting sort by cell (single-GPU step) ===========================================
// The keys are cell numbers below table_size, and the step needs the exclusive prefix of the
// per-cell counts anyway (cell_start).  So instead of three digit passes over the pairs:
//   k_cell_count    count[key]++ with the old value kept as the particle's provisional rank
//                   in its cell (warp-aggregated over runs of equal keys: the input is last
//                   step's sorted order, neighbours in the array mostly share a cell)
//   k_cell_scan_*   cell_start = exclusive prefix of count (tile and group sums, then one sweep
//                   that also zeroes count for the next step)
//   k_cell_scatter  pair (key, index) -> slot cell_start[key] + provisional rank
// The provisional ranks are in atomic (= arbitrary) order; the reorder kernel that follows
// replaces them by the rank of the index among the cell's members (k_reorder<COUNTED>), so
// the final order is (key, index) exactly as the stable radix sort produces it -- the
// summation order downstream, and with it every result, stays deterministic.
// Algorithmic bytes per particle: 4 + 8 (count) + 8 + 4 + 8 (scatter) = 32, plus 12 per table
// entry (scan), against 4 + 16 P - 4 for P radix passes.
constexpr int kScanThreads = 256;
constexpr int kScanVecs = 4;                                   // uint4 loads per thread
constexpr int kScanTile = kScanThreads * kScanVecs * 4;        // 4096 table entries per CTA

__global__ void __launch_bounds__(256)
    k_cell_count(const uint32_t *__restrict__ keys, int n, const int *__restrict__ n_dev,
                 uint32_t *__restrict__ count, uint64_t *__restrict__ tagged) {
    if (n_dev) n = *n_dev;   // slab cluster: the count lives on the device, the grid covers the capacity
    if ((int)(blockIdx.x * 256) >= n) return;
    const int i = blockIdx.x * 256 + threadIdx.x;
    const uint32_t key = i < n ? __ldg(keys + i) : 0u;
    const uint32_t r = cell_rank(count, key, i < n);
    if (i < n) tagged[i] = ((uint64_t)key << 32) | r;
}

// Scan, part 1: sum of every tile of 4096 counts, and of every group of `group` tiles.
__global__ void __launch_bounds__(kScanThreads)
    k_cell_scan_sums(const uint32_t *__restrict__ count, uint32_t entries, uint32_t *__restrict__ tile_sum,
                     uint32_t *__restrict__ group_sum, uint32_t group_shift) {
    __shared__ uint32_t s_warp[kScanThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = blockIdx.x;
    const uint32_t tile_base = tile * (uint32_t)kScanTile;
    uint32_t sum = 0;
#pragma unroll
    for (int v = 0; v < kScanVecs; ++v) {
        const uint32_t e = tile_base + (uint32_t)((v * kScanThreads + tid) * 4);
        if (e + 3 < entries) {
            const uint4 q = __ldg(reinterpret_cast<const uint4 *>(count + e));
            sum += q.x + q.y + q.z + q.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (e + j < entries) sum += __ldg(count + e + j);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) s_warp[warp] = sum;
    __syncthreads();
    if (tid == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < kScanThreads / 32; ++w) t += s_warp[w];
        tile_sum[tile] = t;
        if (t) atomicAdd(group_sum + (tile >> group_shift), t);
    }
}

// Scan, part 2: cell_start = exclusive prefix of count; count is cleared for the next step.  The
// tile's offset = the groups before its group + the tiles before it inside the group (at most
// 2 * sqrt(tiles) words, read by every CTA: no chain between CTAs -- a decoupled look-back was
// measured first and spent its time walking the ~1000 tiles in flight).
__global__ void __launch_bounds__(kScanThreads)
    k_cell_scan_apply(uint32_t *__restrict__ count, uint32_t *__restrict__ cell_start, uint32_t entries,
                      const uint32_t *__restrict__ tile_sum, const uint32_t *__restrict__ group_sum,
                      uint32_t group_shift, uint32_t base) {
    __shared__ uint32_t s_warp[kScanThreads / 32];
    __shared__ uint32_t s_off[kScanThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = blockIdx.x;
    const uint32_t tile_base = tile * (uint32_t)kScanTile;

    // offset of the tile (+ base: the slot of the first sorted particle, slot0 of a slab)
    uint32_t off = tid == 0 ? base : 0u;
    {
        const uint32_t g = tile >> group_shift;
        for (uint32_t k = tid; k < g; k += kScanThreads) off += __ldg(group_sum + k);
        for (uint32_t k = (g << group_shift) + tid; k < tile; k += kScanThreads) off += __ldg(tile_sum + k);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) off += __shfl_xor_sync(0xffffffffu, off, o);
        if (lane == 0) s_off[warp] = off;
    }

    // warp-striped uint4 loads: vector v of warp w covers entries tile_base + ((w*4+v)*32 + lane)*4 ..+3
    uint32_t c[kScanVecs][4];
    uint32_t vsum[kScanVecs];
#pragma unroll
    for (int v = 0; v < kScanVecs; ++v) {
        const uint32_t e = tile_base + (uint32_t)(((warp * kScanVecs + v) * 32 + lane) * 4);
        if (e + 3 < entries) {
            const uint4 q = *reinterpret_cast<const uint4 *>(count + e);
            c[v][0] = q.x; c[v][1] = q.y; c[v][2] = q.z; c[v][3] = q.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                c[v][j] = 0;
                if (e + j < entries) c[v][j] = count[e + j];
            }
        }
        vsum[v] = c[v][0] + c[v][1] + c[v][2] + c[v][3];
    }
    // exclusive prefix of every vector inside its warp
    uint32_t vexcl[kScanVecs];
    uint32_t run = 0;
#pragma unroll
    for (int v = 0; v < kScanVecs; ++v) {
        uint32_t incl = vsum[v];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        vexcl[v] = run + incl - vsum[v];
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) s_warp[warp] = run;
    __syncthreads();
    uint32_t prefix = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) {
        prefix += s_off[w];
        if (w < warp) prefix += s_warp[w];
    }
#pragma unroll
    for (int v = 0; v < kScanVecs; ++v) {
        const uint32_t e = tile_base + (uint32_t)(((warp * kScanVecs + v) * 32 + lane) * 4);
        uint4 o;
        o.x = prefix + vexcl[v];
        o.y = o.x + c[v][0];
        o.z = o.y + c[v][1];
        o.w = o.z + c[v][2];
        if (e + 3 < entries) {
            *reinterpret_cast<uint4 *>(cell_start + e) = o;
        } else {
            if (e < entries) cell_start[e] = o.x;
            if (e + 1 < entries) cell_start[e + 1] = o.y;
            if (e + 2 < entries) cell_start[e + 2] = o.z;
        }
    }
    // count is cleared only now: a store to count between the loads above made every later load
    // wait for it (the compiler cannot tell the addresses apart) -- 156 us instead of 19 us at 16 M cells
#pragma unroll
    for (int v = 0; v < kScanVecs; ++v) {
        const uint32_t e = tile_base + (uint32_t)(((warp * kScanVecs + v) * 32 + lane) * 4);
        if (e + 3 < entries) {
            *reinterpret_cast<uint4 *>(count + e) = make_uint4(0u, 0u, 0u, 0u);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (e + j < entries) count[e + j] = 0;
        }
    }
}

// Four particles per thread, all loads of a stage issued together: the kernel is a chain of two
// dependent loads and a store per particle, and one chain per thread was latency-bound (75-98 us at
// 16 M against ~45 us of traffic).
constexpr int kScatterItems = 4;
__global__ void __launch_bounds__(256)
    k_cell_scatter(const uint64_t *__restrict__ tagged, int n, const int *__restrict__ n_dev,
                   const uint32_t *__restrict__ cell_start, uint32_t base, uint64_t *__restrict__ pairs) {
    if (n_dev) n = *n_dev;
    if ((int)(blockIdx.x * (256 * kScatterItems)) >= n) return;
    const int i0 = blockIdx.x * (256 * kScatterItems) + threadIdx.x;
    uint64_t t[kScatterItems];
    uint32_t c0[kScatterItems];
#pragma unroll
    for (int k = 0; k < kScatterItems; ++k) {
        const int i = i0 + k * 256;
        t[k] = i < n ? __ldg(tagged + i) : 0ull;
    }
#pragma unroll
    for (int k = 0; k < kScatterItems; ++k) c0[k] = __ldg(cell_start + (uint32_t)(t[k] >> 32));
#pragma unroll
    for (int k = 0; k < kScatterItems; ++k) {
        const int i = i0 + k * 256;
        if (i < n) pairs[c0[k] - base + (uint32_t)t[k]] = (t[k] & 0xffffffff00000000ull) | (uint32_t)i;
    }
}

}  // namespace

int sort_tiles(int n) { return (n + kSortTile - 1) / kSortTile; }

size_t sort_scratch_words(int capacity) {
    // [ghist: kMaxPasses*256][tickets: kMaxPasses (padded to 64)][status: passes*tiles*256]
    return (size_t)kMaxPasses * kRadix + 64 + (size_t)kMaxPasses * sort_tiles(capacity) * kRadix;
}

int sort_passes_for(uint32_t table_size) {
    int bits = 1;
    while (bits < 32 && (1ull << bits) < (unsigned long long)table_size) ++bits;
    return (bits + 7) / 8;
}

// Enqueues: clear scratch, histogram, `passes` onesweep passes.  Returns the
// buffer (0 or 1) holding the sorted pairs.  `launches` is incremented per kernel.
int sort_pairs_async(const uint32_t *keys, uint64_t *pairs0, uint64_t *pairs1, int n, int passes,
                     uint32_t *scratch, int sm_count, cudaStream_t stream, SortHooks *hooks,
                     const int *n_dev) {
    const int tiles = sort_tiles(n);
    uint32_t *ghist = scratch;
    uint32_t *tickets = scratch + kMaxPasses * kRadix;
    uint32_t *status = tickets + 64;
    const size_t used_words = (size_t)kMaxPasses * kRadix + 64 + (size_t)passes * tiles * kRadix;
    cudaMemsetAsync(scratch, 0, used_words * sizeof(uint32_t), stream);

    if (hooks) hooks->before(hooks->ctx, kSortStageHistogram);
    int hist_blocks = (n + kSortThreads * 16 - 1) / (kSortThreads * 16);
    hist_blocks = max(1, min(hist_blocks, sm_count * 8));
    k_histogram<<<hist_blocks, kSortThreads, 0, stream>>>(keys, n, n_dev, passes, ghist);
    if (hooks) hooks->after(hooks->ctx, kSortStageHistogram);

    uint64_t *bufs[2] = {pairs0, pairs1};
    int out = 0;
    for (int p = 0; p < passes; ++p) {
        if (hooks) hooks->before(hooks->ctx, kSortStagePass);
        uint32_t *st = status + (size_t)p * tiles * kRadix;
        if (p == 0) {
            k_onesweep<true><<<tiles, kSortThreads, 0, stream>>>(keys, nullptr, bufs[0], n, n_dev, 0, ghist,
                                                               st, tickets + p);
            out = 0;
        } else {
            k_onesweep<false><<<tiles, kSortThreads, 0, stream>>>(
                nullptr, bufs[out], bufs[out ^ 1], n, n_dev, 8 * p, ghist + p * kRadix, st, tickets + p);
            out ^= 1;
        }
        if (hooks) hooks->after(hooks->ctx, kSortStagePass);
    }
    return out;
}


namespace {
uint32_t scan_tiles(uint32_t table_entries) { return (table_entries + kScanTile - 1) / kScanTile; }
uint32_t scan_group_shift(uint32_t tiles) {   // groups of 2^shift tiles, about sqrt(tiles) of them
    uint32_t shift = 0;
    while ((1ull << (2 * shift)) < tiles) ++shift;
    return shift;
}
}  // namespace

size_t cell_sort_scratch_words(uint32_t table_entries) {
    const uint32_t tiles = scan_tiles(table_entries);
    return (size_t)tiles + (tiles >> scan_group_shift(tiles)) + 1;   // [group sums][tile sums]
}

void cell_sort_async(const uint32_t *keys, uint64_t *pairs_sorted, uint64_t *pairs_tmp, int n,
                     uint32_t table_entries, uint32_t *count, uint32_t *cell_start, uint32_t *scratch,
                     cudaStream_t stream, SortHooks *hooks, bool counted, const int *n_dev, uint32_t base) {
    const uint32_t tiles = scan_tiles(table_entries);
    const uint32_t shift = scan_group_shift(tiles);
    const uint32_t groups = (tiles >> shift) + 1;
    uint32_t *group_sum = scratch, *tile_sum = scratch + groups;
    cudaMemsetAsync(group_sum, 0, groups * sizeof(uint32_t), stream);
    const int blocks = (n + 255) / 256;
    if (!counted) {   // (otherwise the force kernel of the previous step counted: CellCount)
        if (hooks) hooks->before(hooks->ctx, kSortStageHistogram);
        if (n > 0) k_cell_count<<<blocks, 256, 0, stream>>>(keys, n, n_dev, count, pairs_tmp);
        if (hooks) hooks->after(hooks->ctx, kSortStageHistogram);
    }
    if (hooks) hooks->before(hooks->ctx, kSortStagePass);
    k_cell_scan_sums<<<tiles, kScanThreads, 0, stream>>>(count, table_entries, tile_sum, group_sum, shift);
    if (hooks) hooks->after(hooks->ctx, kSortStagePass);
    if (hooks) hooks->before(hooks->ctx, kSortStagePass);
    k_cell_scan_apply<<<tiles, kScanThreads, 0, stream>>>(count, cell_start, table_entries, tile_sum,
                                                          group_sum, shift, base);
    if (hooks) hooks->after(hooks->ctx, kSortStagePass);
    if (hooks) hooks->before(hooks->ctx, kSortStagePass);
    if (n > 0)
        k_cell_scatter<<<(n + 256 * kScatterItems - 1) / (256 * kScatterItems), 256, 0, stream>>>(
            pairs_tmp, n, n_dev, cell_start, base, pairs_sorted);
    if (hooks) hooks->after(hooks->ctx, kSortStagePass);
}

}  // namespace sph
