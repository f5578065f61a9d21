// sph_kernels.cuh -- launch interface of the SPH step kernels (sph_kernels.cu).
#pragma once

#include "sph_common.cuh"

namespace sph {

// Device arrays of one simulator.  "cur" = the state, stored in the order of the
// last sort; "srt" = the same particles gathered into this step's sorted order.
// pos.w carries the ORIGINAL particle id (bit pattern), vel.w is unused.
struct DeviceState {
    float4 *cur_pos, *cur_vel;   // state (input of the step, overwritten with its output)
    float4 *srt_pos, *srt_vel;   // sorted copy of the pre-step state
    float4 *pair_xy;             // sorted positions again, two slots per record: {x0,x1,y0,y1}
    float2 *pair_z;              //   and {z0,z1}: operands of the packed f32x2 neighbour loop
    uint32_t *key;               // cell key of cur_pos[i]
    uint64_t *pairs[2];          // (key << 32 | slot) ping-pong buffers of the sort
    uint32_t *cell_start;        // table_size + 1 entries
    float2 *pa;                  // per sorted slot: {pressure, -MASS/(2*density)}
    float *rho;                  // per sorted slot: density
    float4 *force;               // optional: force of the last step per sorted slot
    float *out_pos;              // xyz packed, ORIGINAL particle order (host readback source)
    uint32_t *sort_scratch;
    uint32_t *cell_count;   // counting sort by cell: table_size + 1 words, zero between steps
    double *stats;               // 2 doubles
    int32_t *counts;             // optional 2*n scratch for K and C
    const uint64_t *sorted_pairs;  // pairs[sorted buffer] of this step: key of sorted slot i in the high word
    int stage_tiles;             // 1: dense CTAs run k_density_tile (TMA-staged neighbour tiles)
    int density_exact;           // 1: density summed term by term in the reference's order (bit-identical
                                 // to the CPU restatement); 0: factored sum (DensityAcc in sph_kernels.cu)
    // In-range bit masks density hands to force.  A pool of rows of 32 words; every WARP of a
    // particle CTA takes as many rows as its widest particle needs (word w of lane t = row base + w,
    // column t) from a per-step bump allocator -- 2-3 rows in the undisturbed fluid, tens in the
    // floor pile-up -- instead of a fixed kMaskWords rows per particle (4 GB at 16 M particles).
    // The pool is split into kMaskPools segments with their own cursors (warp w uses segment
    // w % kMaskPools) so that the allocating atomics do not serialise on one address.
    struct MaskPool {
        uint32_t *words;     // rows * 32 words
        uint32_t *base;      // per warp of the particle CTAs: first row, or kNoMaskRows (segment
                             // exhausted: the force kernel repeats the distance tests for that warp)
        uint32_t *cursor;    // kMaskPools cursors, 128 bytes apart: rows handed out this step
        uint32_t rows;       // capacity of ONE segment
    } masks;
    // slab mode: particles that left the owned z-layers during integration, per side
    float4 *emig_pos[2], *emig_vel[2];   // [0] towards lower z, [1] towards higher z
    uint32_t *emig_count[2];             // one counter per side (may exceed emig_capacity: overflow)
    int emig_capacity;
};

// Header of a message between neighbouring slabs (cluster mode); the payload follows it.
// count : entries in the payload (the migration message's count is the force kernel's atomic cursor)
// seq   : written by the SENDER after the payload of exchange round `seq` is complete
// ack   : written by the RECEIVER (into the sender's memory) once it has consumed round `ack`
// seq / ack are only used between slabs of different processes, where the receiver reads the
// sender's buffer directly over NVLink (peer memory); they sit in separate 64-byte blocks.
struct MsgHeader {
    uint32_t count;
    uint32_t pad0[3];
    uint32_t seq;
    uint32_t pad1[11];
    uint32_t ack;
    uint32_t pad2[15];
};
static_assert(sizeof(MsgHeader) == 128, "message payloads start 128-byte aligned");

constexpr int kBlock = 128;      // particles per CTA of the neighbour kernels (ref: simulator.cu:12)
static_assert(kBlock == kBlockParticles, "particle_cta() assumes the neighbour kernels' CTA size");
constexpr int kMaskWords = 64;   // mask capacity per particle: 64 words = up to 2048 candidates
constexpr uint32_t kNoMaskRows = 0xffffffffu;
constexpr int kMaskRowsPerWarp = 12;  // pool size: average rows per warp (48 B per particle)
#ifndef SPH_MASK_POOLS
#define SPH_MASK_POOLS 64
#endif
constexpr int kMaskPools = SPH_MASK_POOLS;

// Predicate thresholds on r^2 that are exactly equivalent to the reference's
// predicates on r = sqrt_rn(r^2) (ref: simulator.cu:110 `dist < EPS_F`,
// simulator.cu:125 `dist > h`); computed on the host by bisection over floats.
struct Thresholds {
    float r2_eps;  // smallest float t with sqrtf(t) >= EPS_F  => (dist < EPS_F) == (r2 < r2_eps)
    float r2_h;    // largest  float t with sqrtf(t) <= h      => (dist > h)     == (r2 > r2_h)
};

void launch_hash(const Params &p, const DeviceState &d, cudaStream_t s);
void launch_reorder(const Params &p, const DeviceState &d, int sorted_buf, int n_sorted, int sm_count,
                    cudaStream_t s);
// Single-GPU step after cell_sort_async(): pairs grouped by cell, cell_start already built.
void launch_reorder_counted(const Params &p, const DeviceState &d, int sorted_buf, cudaStream_t s);
void launch_density(const Params &p, const Thresholds &t, const DeviceState &d, bool counts,
                    cudaStream_t s);
// count_cells (single GPU, flat keys): also produce next step's per-cell counts (d.cell_count) and
// tagged pairs (d.pairs[1]) for the counting sort by cell.
void launch_force_integrate(const Params &p, const Thresholds &t, const DeviceState &d, cudaStream_t s,
                            bool count_cells = false);
// positions of the state (cur_pos, id in .w) -> xyz packed in ORIGINAL particle order
void launch_unpermute(const Params &p, const DeviceState &d, float *out_pos, cudaStream_t s);
void launch_push(const Params &p, const DeviceState &d, int click_x, int click_y, cudaStream_t s);
void launch_stats(const Params &p, const DeviceState &d, cudaStream_t s);
// slab mode: hash of freshly appended particles [first, first+count) of the cur arrays
void launch_hash_range(const Params &p, const DeviceState &d, int first, int count, cudaStream_t s);
// slab mode: cell ranges + pair-interleaved copy of the ghost slots [first, first+count) whose
// keys lie in [key_lo, key_hi); cell_start[k] for k in [key_lo, key_hi] is written
void launch_ghost_prepare(const Params &p, const DeviceState &d, int first, int count,
                          uint32_t key_lo, uint32_t key_hi, cudaStream_t s);

// -- slab cluster (device-resident counts; csrc/sph_cluster.cu) ---------------------------------
void launch_pack_layer(const Params &p, const DeviceState &d, bool pa, MsgHeader *out_lo, MsgHeader *out_hi,
                       int cap, SlabDyn *dyn, cudaStream_t s);
void launch_ghost_install(const Params &p, const DeviceState &d, const MsgHeader *msg, int cap, int side,
                          SlabDyn *dyn, cudaStream_t s);
void launch_ghost_pa(const Params &p, const DeviceState &d, const MsgHeader *msg_lo, const MsgHeader *msg_hi,
                     int cap, cudaStream_t s);
void launch_append_immigrants(const Params &p, const DeviceState &d, const MsgHeader *from_lo,
                              const MsgHeader *from_hi, int cap_m, const MsgHeader *sent_lo,
                              const MsgHeader *sent_hi, int capacity, SlabDyn *dyn, bool rebalance,
                              cudaStream_t s, bool count_cells = false);
void launch_rekey_emigrate(const Params &p, const DeviceState &d, cudaStream_t s);
// peer-memory hand-shake on up to two message headers (either may be null), one tiny kernel each:
// wait until seq / ack reaches `round`, or set it (after a system-wide fence)
enum MsgFlagOp : int { kWaitSeq = 0, kWaitAck = 1, kSetSeq = 2, kSetAck = 3 };
void launch_msg_flags(int op, MsgHeader *h0, MsgHeader *h1, uint32_t round, cudaStream_t s);

}  // namespace sph
