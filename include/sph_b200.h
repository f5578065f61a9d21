/*
 * sph_b200.h -- C ABI of the B200-native SPH timestep (libsph_b200.so).
 *
 * This is the drop-in boundary for the hot path of andrew-sha/CUDAFluidSimulator:
 * everything the reference's `class Simulator` (ref: src/simulator.h:53-74,
 * implemented in src/simulator.cu:370-546) does for its two callers
 * (src/main.cpp:65-76 time mode, src/display.cpp:35-64 free mode) is reachable
 * through these entry points with plain pointers and sizes.  include/simulator.h
 * is the source-compatible C++ class that forwards to them.
 *
 * Conventions
 *   - every call returns 0 on success, a cudaError_t (>0) for CUDA failures or a
 *     negative SPH_E_* code; sph_last_error() gives a thread-local message.
 *     (The reference checks nothing -- ref: SURVEY 5.3 -- so there is no error
 *     behaviour to mirror; callers that ignore the int get reference behaviour.)
 *   - particle arrays crossing the boundary are xyz-interleaved float32,
 *     3*numParticles values, indexed by ORIGINAL particle id exactly like the
 *     reference's `position[i]` (ref: simulator.cu:317, 407-409), no matter how
 *     the library orders particles internally.
 *   - there is no CPU fallback: without a CUDA device every call that needs one
 *     fails with the CUDA error.
 */
#ifndef SPH_B200_H
#define SPH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPH_B200_ABI_VERSION 1

/* Layout-identical to the reference's `struct Settings` (ref: simulator.h:19-31,
 * sizeof == 32: bool at 0, int at 4, six floats from 8), so a `Settings *` may be
 * passed directly. */
typedef struct SphSettings {
    uint8_t randomInit; /* ref: bool randomInit */
    uint8_t _pad[3];
    int32_t numParticles;
    float h;
    float v_kernel_coeff;
    float d_kernel_coeff;
    float boxDim;
    float numCellsPerDim; /* a float in the reference */
    float timestep;
} SphSettings;

/* Layout-identical to the reference's `struct Times` (ref: times.h:5-10). */
typedef struct SphTimes {
    double buildGrid; /* hash + radix sort + cell ranges + reorder, seconds */
    double sphUpdate; /* density/pressure + force/integrate, seconds        */
    double memcpy;    /* device -> host position copy, seconds              */
    int32_t iters;
} SphTimes;

typedef struct sph_sim sph_sim; /* opaque simulator handle */

enum { /* cell-key form used for the sort (north_star item 1) */
    SPH_KEY_FLAT = 0,  /* x + n*(y + n*z)          (ref: simulator.cu:78-82) */
    SPH_KEY_MORTON = 1 /* 3-D bit interleave, x in bit 0 (README.md:5)       */
};

enum {
    SPH_E_INVALID = -1,   /* bad argument / bad settings                       */
    SPH_E_STATE = -2,     /* call order (e.g. step before setup)               */
    SPH_E_NOMEM = -3,     /* host allocation failed                            */
    SPH_E_OUT_OF_BOX = -4 /* a particle lies outside [0, boxDim)^3 (the
                             reference printf()s and indexes out of bounds,
                             ref: simulator.cu:60-73)                           */
};

#define SPH_SORT_COUNT 1
#define SPH_SORT_RADIX 2

/* Extra, additive knobs; zero-initialise for reference behaviour. */
typedef struct SphOptions {
    int32_t device;       /* CUDA device ordinal (default 0)                     */
    int32_t key_mode;     /* SPH_KEY_FLAT (default) or SPH_KEY_MORTON            */
    int32_t record_force; /* keep per-particle force of the last step for
                             sph_get_density_pressure_force()                    */
    int32_t use_graph;    /* 1 (default when 0 is passed via sph_create): replay
                             the step as a CUDA graph; 2 = plain launches        */
    int32_t capacity;     /* >= numParticles; room for ghost/migrated particles
                             in slab mode (0 = numParticles)                     */
    /* slab decomposition (multi-GPU); z_cell_lo == z_cell_hi == 0 => whole box  */
    int32_t z_cell_lo;    /* first owned cell layer along z                      */
    int32_t z_cell_hi;    /* one past the last owned cell layer                  */
    int32_t no_mask_handoff; /* 1: force kernel repeats every distance test instead of
                             reading density's in-range bit masks (A/B measurements)   */
    int32_t nz_cells;     /* slab mode: global cell layers along z (0 = numCellsPerDim); the
                             global box is boxDim x boxDim x nz_cells*h                  */
    int32_t ghost_capacity; /* slab mode: max ghost particles per side (0 = capacity/4)     */
    int32_t emig_capacity;  /* unused (the cluster's migration messages hold the emigrants) */
    int32_t pipeline_readback; /* 1: sph_step() overlaps the device->host copy of step k with the
                             computation of step k+1 (started speculatively before sph_step
                             returns).  Positions handed out are exactly those of the blocking
                             mode; the simulator's internal state runs one step ahead, so the
                             state getters and sph_push() refer to step k+1.  Meant for loops that
                             only call sph_step()/sph_positions_host() (`-m time`-like).       */
    int32_t stage_tiles;  /* 1: dense CTAs (128 consecutive particles within a few cells of one
                             row) compute density out of neighbour tiles staged in shared memory
                             by bulk TMA (A/B switch; see DESIGN.md 3.4 for the measurement)    */
    int32_t density_sum;  /* how a particle's density terms are added up: 0 = library default,
                             1 = term by term in the reference's loop order (bit-identical to the
                             serial restatement of simulator.cu:163-185), 2 = factored,
                             rho = (m dk) * sum (h^2 - r^2)^3 with two interleaved partial sums (a few
                             ulp away; the reference's own order is its CAS race order)           */
    int32_t sort_algo;    /* how a single GPU sorts particles by cell every step: 0 = library default
                             (SPH_SORT_COUNT unless the environment says SPH_SORT=radix),
                             SPH_SORT_COUNT = counting sort by cell (per-cell counts, one scan of the
                             table, scatter; members of a cell ranked by index), SPH_SORT_RADIX =
                             8-bit onesweep radix passes.  Both give the same order: (key, index).
                             Slabs of a cluster always use the radix passes.                     */
    int32_t reserved[1];
} SphOptions;

/* --- life cycle (ref: Simulator ctor/dtor/setup, simulator.cu:370-460) ------ */
int sph_create(const SphSettings *settings, sph_sim **out);
int sph_create_ex(const SphSettings *settings, const SphOptions *options, sph_sim **out);
void sph_destroy(sph_sim *sim);
/* Allocates device state and initialises particles exactly as the reference
 * does (random: unseeded glibc rand(), 3 draws/particle; grid: 0.9h lattice,
 * x outer / z inner), then uploads.  ref: simulator.cu:411-460 */
int sph_setup(sph_sim *sim);

/* --- per-timestep advance ----------------------------------------------------
 * sph_step: one timestep + blocking device->host copy of all positions, i.e.
 * Simulator::simulate() without the mouse hand-off (ref: simulator.cu:462-497).
 * sph_step_timed: same, accumulating the reference's three wall-clock buckets,
 * i.e. Simulator::simulateAndTime() (ref: simulator.cu:499-546).
 * sph_advance: `steps` timesteps without any host readback (device-resident
 * benchmarking; no reference equivalent). */
int sph_step(sph_sim *sim);
int sph_step_timed(sph_sim *sim, SphTimes *times);
int sph_advance(sph_sim *sim, int steps);
/* sph_advance bracketed by CUDA events on the simulator's own stream; *ms receives
 * the device time of the `steps` timesteps (measurement helper for bench.py). */
int sph_advance_timed(sph_sim *sim, int steps, float *ms);
/* Mouse push applied to the cell grid of the step that just ran, i.e. what
 * simulate() does when display.cpp set mouseClicked (ref: simulator.cu:329-367,
 * 482-489).  x,y are window pixels. */
int sph_push(sph_sim *sim, int x, int y);

/* --- position readback (ref: Simulator::getPosition, simulator.cu:407-409) ---
 * Pinned host buffer owned by the simulator, 3*numParticles floats in original
 * particle order, refreshed by sph_step / sph_step_timed / sph_readback. */
const float *sph_positions_host(sph_sim *sim);
int sph_readback(sph_sim *sim);

/* --- state injection / parity hooks (test plumbing) ---------- */
int sph_set_state(sph_sim *sim, const float *pos, const float *vel /* may be NULL => 0 */);
int sph_get_state(sph_sim *sim, float *pos, float *vel);
/* cell keys of the CURRENT positions, original particle order */
int sph_get_keys(sph_sim *sim, int key_mode, uint32_t *keys);
/* order of the LAST step's sort: ids[s] = original id at sorted slot s,
 * sorted_keys[s] its key (either may be NULL) */
int sph_get_sorted_index(sph_sim *sim, uint32_t *ids, uint32_t *sorted_keys);
/* cell table of the LAST step: start[k] = first sorted slot with key >= k,
 * table_size+1 entries; *table_size receives the number of keys */
int sph_get_cell_start(sph_sim *sim, uint32_t *start, uint32_t *table_size);
/* K[i] = neighbours with !(r^2 > h^2) (self included), C[i] = candidates in the
 * 27-cell stencil, for the CURRENT positions, original order */
int sph_get_neighbor_counts(sph_sim *sim, int32_t *K, int32_t *C);
/* density / pressure / force the LAST step computed (from its pre-step
 * positions), original order; force needs options.record_force */
int sph_get_density_pressure_force(sph_sim *sim, float *rho, float *prs, float *force);
/* kinetic energy 0.5*m*|v|^2 summed, and mean density of the last step */
int sph_get_stats(sph_sim *sim, double *kinetic_energy, double *mean_density);

/* --- slab mode (building block of the multi-GPU cluster below) ------------------------------
 * A simulator created with SphOptions.z_cell_lo < z_cell_hi owns the global cell layers
 * [z_cell_lo, z_cell_hi) along z and the particles in them; its keys are local to the slab (layer 0
 * and the last layer hold the neighbours' ghost particles).  Such a simulator is stepped by the
 * cluster driver (sph_cluster_*), not by sph_step(); this call hands it its particle set. */
/* Replace the owned particle set (host arrays; ids are global particle ids). */
int sph_slab_load(sph_sim *sim, int n, const float *pos, const float *vel, const uint32_t *ids);

/* --- multi-GPU: a cluster of z-slabs, one per GPU -------------------------------------------
 * north_star: "the domain is partitioned across the 8xB200 box by a spatial slab decomposition,
 * with per-step ghost-particle halo exchange and particle migration over NVLink via NCCL or
 * P2P".  There is no reference equivalent (single GPU); the calls mirror the
 * single-GPU ones above (setup / advance / step / positions).
 *
 * The whole per-step protocol runs inside the library and never brings a count to the host:
 * particle counts live in device memory, every kernel is launched over the slab's capacity, and
 * halo / migration messages are fixed-capacity buffers whose header carries the count.  Slabs
 * driven by the same process exchange them with peer-to-peer copies (cudaMemcpyPeerAsync over
 * NVLink), slabs of different processes with ncclSend / ncclRecv (one process per GPU under
 * torchrun: local_count = 1).  csrc/sph_cluster.cu has the protocol. */
typedef struct sph_cluster sph_cluster;
#define SPH_NCCL_ID_BYTES 128
#define SPH_MAX_LOCAL_SLABS 16

typedef struct SphClusterOptions {
    int32_t world;           /* slabs (== GPUs) of the whole job                                  */
    int32_t first_rank;      /* global index of this process's first slab                         */
    int32_t local_count;     /* slabs this process drives: world for a single-process run,
                                1 with one process per GPU                                        */
    int32_t devices[SPH_MAX_LOCAL_SLABS]; /* CUDA device of each local slab (may repeat: several
                                slabs on one GPU, used by the tests)                              */
    int32_t nz_cells;        /* global cell layers along z (0 = numCellsPerDim); the global box is
                                boxDim x boxDim x nz_cells*h                                      */
    int32_t capacity;        /* particles per slab (0 = 1.25 * numParticles / world + 65536)      */
    int32_t ghost_capacity;  /* ghost particles per face (0 = capacity / 8)                       */
    int32_t emig_capacity;   /* emigrants per face and step (0 = capacity / 32)                   */
    int32_t density_sum;     /* as SphOptions.density_sum                                         */
    int32_t rebalance_every; /* sph_cluster_advance moves the slab boundaries towards equal particle
                                counts every this many steps (0 = static slabs)                   */
    int32_t reserved[6];
    uint8_t nccl_id[SPH_NCCL_ID_BYTES]; /* from sph_cluster_nccl_id() on rank 0, broadcast by the
                                caller; only read when local_count < world                        */
} SphClusterOptions;

typedef struct SphSlabStats {
    int32_t rank, device;
    int32_t z_cell_lo, z_cell_hi;   /* owned global cell layers [lo, hi)                          */
    int32_t n_owned;                /* live particles                                             */
    int32_t ghosts_lo, ghosts_hi;   /* ghost particles of the last step                           */
    int32_t steps;
    int64_t migrated_total;         /* particles received from neighbours since the load          */
    int64_t ghosts_total;           /* ghost particles installed since the load                   */
    uint32_t overflow;              /* non-zero: a capacity was exceeded, particles were lost      */
    int32_t rebalances;             /* boundary moves applied so far                              */
    double kinetic_energy;          /* 0.5 m |v|^2 summed over the owned particles                */
    double density_sum;             /* sum of the last step's densities over its live particles   */
    uint32_t debug_flags;           /* violation bits of the self-checking build (sph_debug_flags) */
    int32_t checked_build;          /* 1 in the self-checking build                               */
} SphSlabStats;

/* A fresh NCCL unique id (rank 0 calls this and broadcasts the bytes by any means). */
int sph_cluster_nccl_id(uint8_t id[SPH_NCCL_ID_BYTES]);
int sph_cluster_create(const SphSettings *settings, const SphClusterOptions *options, sph_cluster **out);
void sph_cluster_destroy(sph_cluster *c);
/* The reference's initialisation (ref: simulator.cu:430-453, as sph_setup) of the whole box; every
 * slab keeps the particles of its layers.  Ids are the single-simulator ids. */
int sph_cluster_setup(sph_cluster *c);
/* Explicit particle set of local slab `local_index` (host arrays; ids are global). */
int sph_cluster_load(sph_cluster *c, int local_index, int n, const float *pos, const float *vel,
                     const uint32_t *ids);
/* `steps` timesteps, device-resident; one synchronisation at the end. */
int sph_cluster_advance(sph_cluster *c, int steps);
/* The same between CUDA events on every local slab's stream; *ms = slowest local slab. */
int sph_cluster_advance_timed(sph_cluster *c, int steps, float *ms);
/* One timestep, then every owned particle's record {x, y, z, id} on its way to the slab's pinned
 * host buffer (the copy overlaps the next step; sph_cluster_sync() or the next call completes it):
 * Simulator::simulate()'s "positions on the host after every step", per slab.  A record whose id
 * is 0xffffffff belongs to a particle that has just left the slab. */
int sph_cluster_step(sph_cluster *c);
int sph_cluster_sync(sph_cluster *c);
int sph_cluster_host_records(sph_cluster *c, int local_index, const float **records, int *count);
/* Positions of the local slabs' particles scattered into out[3 * id ...] -- getPosition() for
 * the ids this process holds (all of them in a single-process run). */
int sph_cluster_positions(sph_cluster *c, float *out, int64_t n_global);
int sph_cluster_download(sph_cluster *c, int local_index, uint32_t *ids, float *pos, float *vel, int *n);
int sph_cluster_stats(sph_cluster *c, int local_index, SphSlabStats *out);
/* Moves slab boundaries by at most one layer each towards equal particle counts (collective:
 * every process of the job calls it at the same step). */
int sph_cluster_rebalance(sph_cluster *c);
int64_t sph_cluster_launch_count(sph_cluster *c);

/* Self-checking build (nvcc -DSPH_BOUNDS_CHECK, `python -m cudafluidsimulator_b200.build
 * --checked`): every data-dependent index of the hot kernels is verified on the device and
 * violations are OR-ed into *flags (bit meanings: SPH_DBG_* in csrc/sph_common.cuh).  In the
 * normal build *flags stays 0 and *checked_build is 0. */
int sph_debug_flags(sph_sim *sim, uint32_t *flags, int *checked_build);

/* --- measurement -------------------------------------------------------------
 * Per-kernel CUDA-event times (ms, summed since the last reset) for the stages
 * hash, histogram, sort passes, reorder+cell ranges, density, force+integrate.
 * With the counting sort by cell (sph_sort_info) "histogram" is the stand-alone count kernel (it
 * only runs after a state upload; otherwise the count is part of force+integrate) and "sort
 * passes" are the two table-scan kernels and the scatter.
 * Enabling adds event records around every launch (and disables the graph). */
#define SPH_STAGE_COUNT 8
int sph_profile_enable(sph_sim *sim, int on);
int sph_profile_read(sph_sim *sim, double ms[SPH_STAGE_COUNT], int64_t launches[SPH_STAGE_COUNT],
                     int reset);
const char *sph_stage_name(int stage);
/* total kernel launches issued by this simulator so far */
int64_t sph_launch_count(sph_sim *sim);
int sph_num_particles(sph_sim *sim);
/* How this simulator sorts particles by cell: *algo = SPH_SORT_COUNT or SPH_SORT_RADIX, *kernels =
 * sort kernels launched per step (radix: histogram + passes; count: table scan x2 + scatter, the
 * count itself being fused into the force kernel when *count_fused), *radix_passes = 8-bit passes
 * the radix sort needs for this key range. */
int sph_sort_info(sph_sim *sim, int32_t *algo, int32_t *kernels, int32_t *count_fused, int32_t *radix_passes);

const char *sph_last_error(void);
int sph_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SPH_B200_H */
