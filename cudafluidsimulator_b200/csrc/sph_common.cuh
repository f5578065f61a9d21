// sph_common.cuh -- shared device-side definitions of the B200 SPH step.
//
// Physics constants and rounding rules follow the reference
// (ref: src/simulator.h:6-12, src/simulator.cu:12-14).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace sph {

constexpr float kMass = 0.02f;          // ref: simulator.h:7  MASS
constexpr float kRestDensity = 1000.f;  // ref: simulator.h:9  REST_DENSITY
constexpr float kGravity = -9.8f;       // ref: simulator.h:11 GRAVITY
constexpr float kEps = 1e-4f;           // ref: simulator.cu:14 EPS_F
constexpr float kPushStrength = 5.f;    // ref: simulator.cu:13 PUSH_STRENGTH
// GAS_CONSTANT == VISCOSITY == 1 and ELASTICITY == 0.5 are folded in place.

enum KeyMode : int { kKeyFlat = 0, kKeyMorton = 1 };

// Slab cluster (multi-GPU, csrc/sph_cluster.cu): the particle counts of a slab live in device
// memory and never cross the host inside a step; kernels are launched over the slab's capacity and
// read what they need from here.
struct SlabDyn {
    int n_total;        // entries of the cur arrays in use (live + emigrated-dead + appended immigrants)
    int n_live;         // live particles = sorted slots [slot0, slot0 + n_live) of this step
    int n_dead;         // entries among n_total that emigrated in the previous step (sort behind the live ones)
    int g_lo, g_hi;     // ghost particles installed below / above the owned slots this step
    int steps;          // steps taken (statistics)
    unsigned overflow;  // SPH_OVF_* bits: a capacity was exceeded, particles were lost
    int n_prev;         // live particles of the step that just finished (its rho / pa slots)
    int lo_count, hi_count;        // particles in the lowest / highest owned layer (this step's sort)
    unsigned long long migrated;   // particles received from neighbours so far (statistics)
    unsigned long long ghosts;     // ghost particles installed so far (statistics)
};
enum : unsigned { SPH_OVF_CAPACITY = 1u, SPH_OVF_GHOSTS = 2u, SPH_OVF_EMIGRANTS = 4u };

// Kernel parameters, passed by value (__grid_constant__) -- no __constant__
// symbol, so several simulators (e.g. one per slab) can coexist in a process.
struct Params {
    int n;          // particles in the arrays (owned + ghosts)
    int n_owned;    // particles this simulator integrates (== n without slabs)
    int nc;         // cells per dimension (int of the reference's float)
    float h;        // smoothing length == cell edge
    float h2;       // h*h, rounded once (ref: simulator.cu:89)
    float vk;       // v_kernel_coeff = 45/(pi h^6)
    float dk;       // d_kernel_coeff = 315/(64 pi h^9)
    float box;      // boxDim
    float hi;       // boxDim - h, rounded once (ref: simulator.cu:283)
    float dt;       // timestep
    int key_mode;   // KeyMode
    uint32_t table_size;  // number of distinct keys (nc^2 * ncz flat, 8^bits Morton)
    // -- slab decomposition along z (multi-GPU); single GPU: slot0 = 0, zoff = 0, ncz = nc,
    //    [zlo, zhi) = [0, nc), nz = nc, hi_z = hi
    int slab;       // 1 in slab mode
    int slot0;      // first target slot of the sorted arrays (ghost capacity in slab mode)
    int zoff;       // global z cell of local layer 0 (slab: zlo - 1, the low ghost layer)
    int ncz;        // local layers along z (slab: owned + 2 ghost layers)
    int zlo, zhi;   // owned global z-cell layers [zlo, zhi)
    int nz;         // global cells along z
    float hi_z;     // global box length along z minus h (z wall)
    uint32_t dead_key;  // slab: key given to emigrated particles; sorts behind every live key
    // -- launch over a subset of the particle CTAs (slab cluster: the interior CTAs run before the halo
    //    is waited for, the boundary CTAs after it arrived).  part 0 = every CTA; 1 = interior only
    //    (a CTA that holds a particle of the lowest / highest owned layer skips); 2 = those boundary CTAs
    //    only, launched as 2 * part_ctas blocks (low boundary, then high boundary).  The boundary
    //    ranges come from SlabDyn (n_live, lo_count, hi_count), i.e. they are decided on the device.
    int part, part_ctas;
    int cta_count;        // grid size to launch, 0 = one CTA per kBlock particles of n
    const SlabDyn *dyn;   // slab cluster: counts in device memory (n, n_owned are then upper bounds)
    // -- self-checking build (-DSPH_BOUNDS_CHECK; compute-sanitizer is closed on the GPU pool)
    int slot_begin, slot_end;   // sorted slots that hold particles this step (ghosts included)
    uint32_t *dbg;              // violation bits are OR-ed in here (see SPH_DBG_* below)
};

// Violation bits of the self-checking build, read back with sph_debug_flags().
enum : uint32_t {
    SPH_DBG_RUN_BOUNDS = 1u,    // a stencil run [s, e) is reversed or leaves the valid slot range
    SPH_DBG_TABLE_INDEX = 2u,   // a cell_start index beyond the table
    SPH_DBG_GATHER_INDEX = 4u,  // a sorted pair points outside the particle arrays
    SPH_DBG_MASK_WORDS = 8u,    // more mask words than kMaskWords were about to be written / read
    SPH_DBG_EMIGRANT = 16u,     // emigrant slot beyond the buffer (counted as overflow, not written)
};

#ifdef SPH_BOUNDS_CHECK
#define SPH_CHECK(p, cond, bit) do { if (!(cond) && (p).dbg) atomicOr((p).dbg, (bit)); } while (0)
#else
#define SPH_CHECK(p, cond, bit) do { } while (0)
#endif

// Live particles of this step: a host-known launch parameter, or the slab's device-side count.
__device__ __forceinline__ int live_count(const Params &p) { return p.dyn ? p.dyn->n_live : p.n; }

// Particle CTA (kBlockParticles consecutive sorted particles) this thread block works on, or -1 if
// the block has nothing to do in this launch (see Params::part).
constexpr int kBlockParticles = 128;
__device__ __forceinline__ int particle_cta(const Params &p) {
    const int c = (int)blockIdx.x;
    if (p.part == 0) return c;
    const int n = p.dyn->n_live;
    const int lo_end = (p.dyn->lo_count + kBlockParticles - 1) / kBlockParticles;          // CTAs [0, lo_end) touch the lowest layer
    const int hi_first = max((n - p.dyn->hi_count) / kBlockParticles, lo_end);             // CTAs [hi_first, ...) the highest
    if (p.part == 1) return (c < lo_end || c >= hi_first) ? -1 : c;
    if (c < p.part_ctas) return c < lo_end ? c : -1;
    const int h = hi_first + (c - p.part_ctas);
    return h * kBlockParticles < n ? h : -1;
}

// ---- cell coordinates and keys ------------------------------------------------
// ref: simulator.cu:57-76 getGridCell: IEEE divide by h, truncate toward zero.
// Coordinates are clamped into the table so that no state can index out of
// bounds (the reference only printf()s).
__device__ __forceinline__ int cell_coord(float x, const Params &p) {
    int c = __float2int_rz(__fdiv_rn(x, p.h));
    return min(max(c, 0), p.nc - 1);
}
// z: the GLOBAL cell (same IEEE rule), and the layer index local to this slab
__device__ __forceinline__ int cell_coord_zglobal(float z, const Params &p) {
    int c = __float2int_rz(__fdiv_rn(z, p.h));
    return min(max(c, 0), p.nz - 1);
}
__device__ __forceinline__ int cell_coord_z(float z, const Params &p) {
    return min(max(cell_coord_zglobal(z, p) - p.zoff, 0), p.ncz - 1);
}

__device__ __forceinline__ uint32_t spread3(uint32_t v) {
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

// ref: simulator.cu:78-82 flattenGridCoord, evaluated in integers (equal to the
// reference's float evaluation while nc^3 <= 2^24).
__device__ __forceinline__ uint32_t key_flat(int cx, int cy, int cz, int nc) {
    return (uint32_t)cx + (uint32_t)nc * ((uint32_t)cy + (uint32_t)nc * (uint32_t)cz);
}

__device__ __forceinline__ uint32_t key_morton(int cx, int cy, int cz) {
    return spread3((uint32_t)cx) | (spread3((uint32_t)cy) << 1) | (spread3((uint32_t)cz) << 2);
}

template <int MODE>
__device__ __forceinline__ uint32_t cell_key(int cx, int cy, int cz, int nc) {
    return MODE == kKeyFlat ? key_flat(cx, cy, cz, nc) : key_morton(cx, cy, cz);
}

// ref: simulator.cu:85-88 as nvcc contracts it for sm_100a:
// dy*dy is a rounded product, the dx and dz terms are fused.  Neighbour counts
// are bit-exact only with exactly this contraction.
__device__ __forceinline__ float dist2(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// ---- 27-cell stencil as index runs over the sorted arrays ----------------------
// cell_start[k] = first sorted slot whose key is >= k (table_size + 1 entries),
// so the particles of cell k are [cell_start[k], cell_start[k+1]).
// Flat keys: the three x-neighbours of a row are adjacent keys, so the stencil is
// 9 contiguous runs.  Morton keys: 27 single-cell runs.
// Visiting order is the reference's dz, dy, dx loop nest (ref: simulator.cu:163-176).
template <int MODE, typename F>
__device__ __forceinline__ void for_each_run(const Params &p, int cx, int cy, int cz,
                                             const uint32_t *__restrict__ cell_start, F &&f) {
#pragma unroll 1
    for (int dz = -1; dz <= 1; ++dz) {
        const int zz = cz + dz;
        if (zz < 0 || zz >= p.ncz) continue;
#pragma unroll 1
        for (int dy = -1; dy <= 1; ++dy) {
            const int yy = cy + dy;
            if (yy < 0 || yy >= p.nc) continue;
            if (MODE == kKeyFlat) {
                const int x0 = max(cx - 1, 0), x1 = min(cx + 1, p.nc - 1);
                const uint32_t row = (uint32_t)p.nc * ((uint32_t)yy + (uint32_t)p.nc * (uint32_t)zz);
                f(__ldg(cell_start + row + x0), __ldg(cell_start + row + x1 + 1));
            } else {
#pragma unroll 1
                for (int dx = -1; dx <= 1; ++dx) {
                    const int xx = cx + dx;
                    if (xx < 0 || xx >= p.nc) continue;
                    const uint32_t m = key_morton(xx, yy, zz);
                    f(__ldg(cell_start + m), __ldg(cell_start + m + 1));
                }
            }
        }
    }
}


// Counting sort by cell (sph_sort.cu): registers the calling lane's key in count[] and returns the
// lane's provisional rank inside its cell.  Lanes that are neighbours in the warp and share a key
// join in one atomic (the array is last step's sorted order: neighbours mostly share a cell).
// All 32 lanes call it together; `active` = the lane has a particle.
__device__ __forceinline__ uint32_t cell_rank(uint32_t *__restrict__ count, uint32_t key, bool active) {
    const int lane = threadIdx.x & 31;
    const uint32_t k = active ? key : 0xffffffffu;   // never equal to a cell number
    const uint32_t prev = __shfl_up_sync(0xffffffffu, k, 1);
    const bool head = lane == 0 || k != prev;
    const uint32_t heads = __ballot_sync(0xffffffffu, head);
    const uint32_t upto = (2u << lane) - 1u;         // lanes <= mine (lane 31: all)
    const int start = 31 - __clz(heads & upto);
    const uint32_t above = heads & ~upto;
    const int end = above ? __ffs(above) - 1 : 32;
    uint32_t base = 0;
    if (head && active) base = atomicAdd(count + k, (uint32_t)(end - start));
    base = __shfl_sync(0xffffffffu, base, start);
    return base + (uint32_t)(lane - start);
}
// Output of the count fused into the force kernel: next step's counts and tagged pairs.
struct CellCount {
    uint32_t *count;    // nullptr: not fused
    uint64_t *tagged;   // (key << 32 | provisional rank) per particle of the new state
};

}  // namespace sph
