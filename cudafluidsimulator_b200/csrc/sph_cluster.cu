// sph_cluster.cu -- the SPH step across several GPUs: z-slab decomposition, per-step ghost halo
// exchange and particle migration (north_star).  There is no reference counterpart
// (the reference is single-GPU); the single-GPU step this builds on replaces
// ref src/simulator.cu:462-497.
//
// One slab-mode simulator (sph_api.cu) per GPU owns the global cell layers [zlo, zhi) along z --
// gravity is -y and the reference's grid init fills x-planes first, so z-slabs start balanced.
// Keys are local to the slab (layer 0 and ncz-1 are the ghost layers).  Per step and slab, all on
// the slab's own stream and WITHOUT any host round trip:
//
//   sort (n_total from device memory) -> reorder (n_live) -> pack the lowest / highest owned layer
//   [halo A]   pos+vel of those layers -> the neighbours' ghost slots
//   ghost install (cell ranges of the ghost layers) -> density of the owned particles
//   [halo B]   {p, a} of the same layers -> the neighbours' ghost slots
//   force + integrate; a particle whose new z cell leaves [zlo, zhi) is appended to the migration
//   message of that side and gets the dead key (the next sort parks it behind the live ones)
//   [migration] emigrants -> appended behind the neighbour's particles, keyed
//
// Every count (particles, boundary layers, ghosts, emigrants) stays in device memory (SlabDyn and
// the message headers); kernels are launched over capacities.  How a message reaches the neighbour:
//   * slabs of this process: cudaMemcpyPeerAsync of the fixed-capacity buffer (NVLink P2P),
//     ordered by events;
//   * slabs of different processes (one process per GPU): the RECEIVER's unpack kernel reads the
//     sender's buffer in place over NVLink -- the buffers are mapped into the neighbour with CUDA
//     IPC at creation -- so exactly `count` entries cross the link and nothing is staged.  Sender
//     and receiver hand-shake on two words of the message header (seq: "round r is complete",
//     ack: "round r has been consumed"), written and polled by one-thread kernels on the slabs'
//     own streams: the transfer is fused into the unpack kernels and the host never waits.
//     NCCL only carries the IPC handles and the rebalancing numbers.  SPH_CLUSTER_NCCL_DATA=1
//     moves the messages with ncclSend / ncclRecv instead (whole buffers; the A/B comparison).
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

#include "sph_internal.cuh"
#include "sph_sort.cuh"

using namespace sph;

extern "C" int sph_internal_fail(int code, const char *fmt, ...);

namespace {

#define CU(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return sph_internal_fail((int)e__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                                     __FILE__, __LINE__);                                          \
    } while (0)

// ---- NCCL, resolved at run time: only a multi-process job needs it ------------------------------
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
    if (g_nccl.lib) return 0;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names)
        if ((h = dlopen(n, RTLD_NOW | RTLD_GLOBAL)) != nullptr) break;
    if (!h) return sph_internal_fail(SPH_E_STATE, "NCCL is needed for a multi-process cluster but libnccl.so.2 was not found: %s", dlerror());
#define SYM(field, name)                                                                        \
    do {                                                                                        \
        *(void **)(&g_nccl.field) = dlsym(h, name);                                             \
        if (!g_nccl.field) return sph_internal_fail(SPH_E_STATE, "libnccl lacks %s", name);     \
    } while (0)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    g_nccl.lib = h;
    return 0;
}

#define NC(expr)                                                                                  \
    do {                                                                                          \
        ncclResult_t r__ = (expr);                                                                \
        if (r__ != ncclSuccess)                                                                   \
            return sph_internal_fail(SPH_E_STATE, "%s failed: %s (%s:%d)", #expr,                 \
                                     g_nccl.GetErrorString(r__), __FILE__, __LINE__);             \
    } while (0)

enum MsgKind { kHaloA = 0, kHaloB = 1, kMigrate = 2, kMsgKinds = 3 };
// phases of a step, for SPH_CLUSTER_TRACE (device time between CUDA events on the slab's stream)
enum Phase { kPhSort = 0, kPhPackA, kPhDensityIn, kPhGhostsIn, kPhDensityBd, kPhPackB, kPhForceIn, kPhPressuresIn,
             kPhForceBd, kPhMigrate, kPhases };
const char *kPhaseNames[kPhases] = {"sort+reorder", "pack halo A", "density (interior)", "wait + ghosts in",
                                    "density (boundary)", "pack halo B", "force (interior)", "wait + pressures in",
                                    "force (boundary)", "wait + immigrants in"};

struct Slab {
    sph_sim *sim = nullptr;
    SlabCore core{};
    int rank = 0, device = 0;
    int zlo = 0, zhi = 0;
    SlabDyn *dyn = nullptr;        // device
    SlabDyn *dyn_host = nullptr;   // pinned mirror, refreshed asynchronously
    // messages: [kind][side]; side 0 = towards / from the slab below, 1 = above
    MsgHeader *send[kMsgKinds][2] = {};
    MsgHeader *recv[kMsgKinds][2] = {};
    MsgHeader *remote[kMsgKinds][2] = {};      // the send buffer of a neighbour in ANOTHER process that faces
                                               // this slab, mapped with CUDA IPC (peer memory)
    cudaEvent_t ev_packed[kMsgKinds] = {};     // this slab's send buffers of that kind are complete
    cudaEvent_t ev_copied[kMsgKinds][2] = {};  // this slab has copied the message of its neighbour on that side
    bool copied_pending[kMsgKinds][2] = {};    // ... and that neighbour has not waited for it yet
    ncclComm_t comm = nullptr;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    // per-step host records
    float4 *out_stage = nullptr;   // device copy of the records, so the next step can overwrite cur_pos
    float4 *host_records = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_out_ready = nullptr, ev_out_done = nullptr;
    bool out_pending = false;
    int out_count = 0;
    int known_total = 0;           // host-side bound on n_total
    int rebalances = 0;
    double *stats_dev = nullptr;   // 2 doubles
    int *info_dev = nullptr;       // 3 x 8 ints: rebalancing numbers of this slab / from below / from above
    bool hashed = false;
    bool counted = false;       // d.cell_count / d.pairs[1] hold the next step's counts (fused into force + append)
    bool count_dirty = false;   // ... or counts that no longer describe the particles: clear before counting
    // SPH_CLUSTER_TRACE=1: CUDA events at the phase boundaries of every step
    std::vector<cudaEvent_t> trace_ev;
    size_t trace_used = 0;
    double phase_ms[kPhases] = {};
};

}  // namespace

struct sph_cluster {
    SphSettings st{};
    SphClusterOptions opt{};
    int nz = 0, nc = 0;
    int cap = 0, cap_g = 0, cap_m = 0;
    size_t msg_bytes[kMsgKinds] = {};
    std::vector<Slab> slabs;            // local slabs, ascending rank
    std::vector<int> zlo, zhi;          // every rank's layer range
    std::vector<int> max_layers;        // every rank's cell-table room, in owned layers
    int64_t launches = 0;
    int steps_since_rebalance = 0;
    bool loaded = false;
    bool trace = false;                 // SPH_CLUSTER_TRACE
    bool fuse_count = false;            // counting sort by cell with the count fused into force + append (SPH_FUSE_COUNT=0: off)
    bool peer_mem = false;              // messages between processes are read in place over NVLink
    uint32_t round[kMsgKinds] = {1, 1, 1};   // exchange rounds so far + 1, per message kind (same on every process)
};

namespace {

bool is_local(const sph_cluster *c, int rank) {
    return rank >= c->opt.first_rank && rank < c->opt.first_rank + c->opt.local_count;
}
Slab *local_slab(sph_cluster *c, int rank) { return &c->slabs[rank - c->opt.first_rank]; }

void split_layers(int nz, int world, std::vector<int> &lo, std::vector<int> &hi) {
    lo.resize(world);
    hi.resize(world);
    const int base = nz / world, extra = nz % world;
    int z = 0;
    for (int r = 0; r < world; ++r) {
        const int n = base + (r < extra ? 1 : 0);
        lo[r] = z;
        hi[r] = z + n;
        z += n;
    }
}

// Layer range -> the slab's kernel parameters (keys are local: layer = z - zoff).
void apply_range(Slab &s, int zlo, int zhi) {
    Params &p = *s.core.p;
    s.zlo = zlo;
    s.zhi = zhi;
    p.zlo = zlo;
    p.zhi = zhi;
    p.zoff = zlo - 1;
    p.ncz = zhi - zlo + 2;
    p.table_size = (uint32_t)p.nc * p.nc * (uint32_t)p.ncz;
    p.dead_key = p.table_size - 1u;
    *s.core.passes = sort_passes_for(p.table_size);
}

int alloc_msg(MsgHeader **out, size_t bytes) {
    CU(cudaMalloc(out, bytes));
    CU(cudaMemset(*out, 0, bytes));
    return 0;
}

// ---- one exchange of one message kind between all neighbouring slabs ---------------------------
// Local neighbour: the RECEIVER's stream waits for the sender's pack event and copies the whole
// buffer peer to peer; the sender's next pack of that kind waits for the "taken" event.
// Remote neighbour: ncclSend / ncclRecv of the whole buffer on the slab's own stream.
int exchange(sph_cluster *c, int kind) {
    const size_t bytes = c->msg_bytes[kind];
    for (Slab &s : c->slabs) {
        CU(cudaSetDevice(s.device));
        bool grouped = false;
        for (int side = 0; side < 2; ++side) {
            const int peer = side ? s.rank + 1 : s.rank - 1;
            if (peer < 0 || peer >= c->opt.world) continue;
            if (is_local(c, peer)) {
                Slab &o = *local_slab(c, peer);
                // my ghosts from `side` are the peer's message towards me: its side (1 - side)
                CU(cudaStreamWaitEvent(s.core.stream, o.ev_packed[kind], 0));
                CU(cudaMemcpyPeerAsync(s.recv[kind][side], s.device, o.send[kind][1 - side], o.device, bytes,
                                       s.core.stream));
                CU(cudaEventRecord(s.ev_copied[kind][side], s.core.stream));
                s.copied_pending[kind][side] = true;
            } else if (!c->peer_mem) {
                if (!grouped) {
                    NC(g_nccl.GroupStart());
                    grouped = true;
                }
                NC(g_nccl.Send(s.send[kind][side], bytes, ncclChar, peer, s.comm, s.core.stream));
                NC(g_nccl.Recv(s.recv[kind][side], bytes, ncclChar, peer, s.comm, s.core.stream));
            }   // (peer memory: nothing to move, the unpack kernels read the neighbour's buffer)
        }
        if (grouped) NC(g_nccl.GroupEnd());
    }
    c->round[kind] += 1;
    return 0;
}

// Before a slab overwrites its send buffers of `kind`: the local neighbours must have copied
// the previous contents out.
int wait_taken(sph_cluster *c, Slab &s, int kind) {
    for (int side = 0; side < 2; ++side) {
        const int peer = side ? s.rank + 1 : s.rank - 1;
        if (peer < 0 || peer >= c->opt.world || !is_local(c, peer)) continue;
        Slab &o = *local_slab(c, peer);
        if (o.copied_pending[kind][1 - side]) {   // the peer copies my message from ITS side (1 - side)
            CU(cudaStreamWaitEvent(s.core.stream, o.ev_copied[kind][1 - side], 0));
            o.copied_pending[kind][1 - side] = false;
        }
    }
    return 0;
}

bool has_peer(const sph_cluster *c, const Slab &s, int side) {
    const int peer = side ? s.rank + 1 : s.rank - 1;
    return peer >= 0 && peer < c->opt.world;
}
bool peer_is_remote_mem(const sph_cluster *c, const Slab &s, int side) {
    const int peer = side ? s.rank + 1 : s.rank - 1;
    return has_peer(c, s, side) && !is_local(c, peer) && c->peer_mem;
}
// send buffer of `kind` towards `side`, or null without a neighbour there
MsgHeader *out_msg(const sph_cluster *c, const Slab &s, int kind, int side) {
    return has_peer(c, s, side) ? s.send[kind][side] : nullptr;
}
// where the unpack kernels find the neighbour's message: the local copy, or the neighbour's own
// buffer (peer memory); null without a neighbour
MsgHeader *in_msg(const sph_cluster *c, const Slab &s, int kind, int side) {
    if (!has_peer(c, s, side)) return nullptr;
    return peer_is_remote_mem(c, s, side) ? s.remote[kind][side] : s.recv[kind][side];
}
// the hand-shake (no-ops unless the neighbour on that side is reached through peer memory)
void before_pack(sph_cluster *c, Slab &s, int kind) {   // the previous round of my buffers has been consumed
    launch_msg_flags(kWaitAck, peer_is_remote_mem(c, s, 0) ? s.send[kind][0] : nullptr,
                     peer_is_remote_mem(c, s, 1) ? s.send[kind][1] : nullptr, c->round[kind] - 1, s.core.stream);
}
void after_pack(sph_cluster *c, Slab &s, int kind) {    // this round of my buffers is complete
    cudaEventRecord(s.ev_packed[kind], s.core.stream);  // (what a neighbour in this process waits for)
    launch_msg_flags(kSetSeq, peer_is_remote_mem(c, s, 0) ? s.send[kind][0] : nullptr,
                     peer_is_remote_mem(c, s, 1) ? s.send[kind][1] : nullptr, c->round[kind], s.core.stream);
}
void before_unpack(sph_cluster *c, Slab &s, int kind, uint32_t round) {   // the neighbours' round is complete
    launch_msg_flags(kWaitSeq, peer_is_remote_mem(c, s, 0) ? s.remote[kind][0] : nullptr,
                     peer_is_remote_mem(c, s, 1) ? s.remote[kind][1] : nullptr, round, s.core.stream);
}
void after_unpack(sph_cluster *c, Slab &s, int kind, uint32_t round) {    // ... and I have consumed it
    launch_msg_flags(kSetAck, peer_is_remote_mem(c, s, 0) ? s.remote[kind][0] : nullptr,
                     peer_is_remote_mem(c, s, 1) ? s.remote[kind][1] : nullptr, round, s.core.stream);
}

// SPH_CLUSTER_TRACE: an event on the slab's stream; consecutive events bracket the phases of a step
int mark(sph_cluster *c, Slab &s) {
    if (!c->trace) return 0;
    if (s.trace_used == s.trace_ev.size()) {
        cudaEvent_t e;
        CU(cudaEventCreate(&e));
        s.trace_ev.push_back(e);
    }
    CU(cudaEventRecord(s.trace_ev[s.trace_used++], s.core.stream));
    return 0;
}
void collect_trace(sph_cluster *c) {   // after the streams are synchronised
    if (!c->trace) return;
    for (Slab &s : c->slabs) {
        cudaSetDevice(s.device);
        for (size_t i = 0; i + kPhases < s.trace_used; i += kPhases + 1)   // kPhases + 1 events per step
            for (int ph = 0; ph < kPhases; ++ph) {
                float ms = 0.f;
                if (cudaEventElapsedTime(&ms, s.trace_ev[i + ph], s.trace_ev[i + ph + 1]) == cudaSuccess) s.phase_ms[ph] += ms;
            }
        s.trace_used = 0;
    }
}

// Launch parameters of a slab: counts come from SlabDyn, the host only knows the capacity.
Params launch_params(const Slab &s, int cap) {
    Params p = *s.core.p;
    p.n = p.n_owned = cap;
    p.dyn = s.dyn;
    p.part = p.part_ctas = p.cta_count = 0;
    p.slot_begin = 0;
    p.slot_end = cap + 2 * s.core.ghost_cap;
    return p;
}

// The neighbour kernels in two parts: the interior CTAs, which read no ghost data, are enqueued
// BEFORE the stream waits for the neighbours' message, the boundary CTAs after it.  By the time
// the stream reaches the wait, the message has long arrived: the exchange costs no time.
Params part_params(const sph_cluster *c, const Slab &s, int part) {
    Params p = launch_params(s, c->cap);
    p.part = part;
    p.part_ctas = (c->cap_g + kBlock - 1) / kBlock + 1;   // CTAs a boundary layer can touch
    if (part == 2) p.cta_count = 2 * p.part_ctas;
    return p;
}

int enqueue_step(sph_cluster *c) {
    const int cap = c->cap;
    // -- build: sort, reorder, pack the boundary layers; density of the interior -------------------
    for (Slab &s : c->slabs) {
        CU(cudaSetDevice(s.device));
        const Params p = launch_params(s, cap);
        DeviceState &d = *s.core.d;
        cudaStream_t st = s.core.stream;
        if (!s.hashed) {   // after a load: keys of everything
            launch_hash_range(p, d, 0, cap, st);   // (entries beyond n_total are never sorted)
            s.hashed = true;
            c->launches += 1;
        }
        mark(c, s);
        if (s.core.cell_sort) {
            // counting sort by cell over the slab's own table; dead entries (key = the table's last cell)
            // land behind the live ones as they do after the radix passes
            if (s.count_dirty) {
                CU(cudaMemsetAsync(d.cell_count, 0, ((size_t)s.core.table_capacity + 1) * sizeof(uint32_t), st));
                s.count_dirty = false;
            }
            cell_sort_async(d.key, d.pairs[0], d.pairs[1], cap, p.table_size + 1u, d.cell_count, d.cell_start,
                            d.sort_scratch, st, nullptr, s.counted, &s.dyn->n_total, (uint32_t)p.slot0);
            *s.core.sorted_buf = 0;
            d.sorted_pairs = d.pairs[0];
            launch_reorder_counted(p, d, 0, st);
            c->launches += (s.counted ? 4 : 5) - (1 + *s.core.passes + 1);   // (the tally below counts the radix launches)
            s.counted = false;
        } else {
            *s.core.sorted_buf = sort_pairs_async(d.key, d.pairs[0], d.pairs[1], cap, *s.core.passes, d.sort_scratch,
                                                  s.core.sm_count, st, nullptr, &s.dyn->n_total);
            d.sorted_pairs = d.pairs[*s.core.sorted_buf];
            launch_reorder(p, d, *s.core.sorted_buf, cap, s.core.sm_count, st);
        }
        mark(c, s);
        int rc = wait_taken(c, s, kHaloA);
        if (rc) return rc;
        before_pack(c, s, kHaloA);
        launch_pack_layer(p, d, false, out_msg(c, s, kHaloA, 0), out_msg(c, s, kHaloA, 1), c->cap_g, s.dyn, st);
        after_pack(c, s, kHaloA);
        mark(c, s);
        launch_density(part_params(c, s, 1), *s.core.th, d, false, st);
        mark(c, s);
        c->launches += 4 + *s.core.passes;
    }
    const uint32_t round_a = c->round[kHaloA];
    int rc = exchange(c, kHaloA);
    if (rc) return rc;
    // -- ghosts in, density of the boundary layers, pack {p, a}; force of the interior -------------
    for (Slab &s : c->slabs) {
        CU(cudaSetDevice(s.device));
        const Params p = launch_params(s, cap);
        DeviceState &d = *s.core.d;
        cudaStream_t st = s.core.stream;
        before_unpack(c, s, kHaloA, round_a);
        for (int side = 0; side < 2; ++side)
            launch_ghost_install(p, d, in_msg(c, s, kHaloA, side), c->cap_g, side, s.dyn, st);
        after_unpack(c, s, kHaloA, round_a);
        mark(c, s);
        launch_density(part_params(c, s, 2), *s.core.th, d, false, st);
        mark(c, s);
        rc = wait_taken(c, s, kHaloB);
        if (rc) return rc;
        before_pack(c, s, kHaloB);
        launch_pack_layer(p, d, true, out_msg(c, s, kHaloB, 0), out_msg(c, s, kHaloB, 1), c->cap_g, s.dyn, st);
        after_pack(c, s, kHaloB);
        mark(c, s);
        // (emigrants of either force launch go straight into the migration messages)
        rc = wait_taken(c, s, kMigrate);
        if (rc) return rc;
        before_pack(c, s, kMigrate);
        for (int side = 0; side < 2; ++side) CU(cudaMemsetAsync(&s.send[kMigrate][side]->count, 0, sizeof(uint32_t), st));
        launch_force_integrate(part_params(c, s, 1), *s.core.th, d, st, c->fuse_count);
        mark(c, s);
        c->launches += 6;
    }
    const uint32_t round_b = c->round[kHaloB];
    rc = exchange(c, kHaloB);
    if (rc) return rc;
    // -- ghost pressures in, force + integrate of the boundary layers ------------------------------
    for (Slab &s : c->slabs) {
        CU(cudaSetDevice(s.device));
        const Params p = launch_params(s, cap);
        DeviceState &d = *s.core.d;
        cudaStream_t st = s.core.stream;
        before_unpack(c, s, kHaloB, round_b);
        launch_ghost_pa(p, d, in_msg(c, s, kHaloB, 0), in_msg(c, s, kHaloB, 1), c->cap_g, st);
        after_unpack(c, s, kHaloB, round_b);
        mark(c, s);
        launch_force_integrate(part_params(c, s, 2), *s.core.th, d, st, c->fuse_count);
        after_pack(c, s, kMigrate);
        mark(c, s);
        c->launches += 2;
    }
    const uint32_t round_m = c->round[kMigrate];
    rc = exchange(c, kMigrate);
    if (rc) return rc;
    // -- immigrants appended, counts of the next step --------------------------------------------
    for (Slab &s : c->slabs) {
        CU(cudaSetDevice(s.device));
        const Params p = launch_params(s, cap);
        before_unpack(c, s, kMigrate, round_m);
        launch_append_immigrants(p, *s.core.d, in_msg(c, s, kMigrate, 0), in_msg(c, s, kMigrate, 1), c->cap_m,
                                 s.send[kMigrate][0], s.send[kMigrate][1], cap, s.dyn, false, s.core.stream,
                                 c->fuse_count);
        s.counted = c->fuse_count;
        after_unpack(c, s, kMigrate, round_m);
        mark(c, s);
        CU(cudaMemcpyAsync(s.dyn_host, s.dyn, sizeof(SlabDyn), cudaMemcpyDeviceToHost, s.core.stream));
        c->launches += 2;
    }
    c->steps_since_rebalance += 1;
    return 0;
}

int sync_all(sph_cluster *c) {
    for (Slab &s : c->slabs) {
        CU(cudaSetDevice(s.device));
        CU(cudaStreamSynchronize(s.core.stream));
        if (s.copy_stream) CU(cudaStreamSynchronize(s.copy_stream));
        CU(cudaGetLastError());
        s.out_pending = false;
        s.known_total = s.dyn_host->n_total;
        if (c->trace && &s == &c->slabs.back()) collect_trace(c);
        if (s.dyn_host->overflow)
            return sph_internal_fail(SPH_E_STATE,
                                     "slab %d: capacity exceeded (flags %u: 1 = particles, 2 = ghost layer, 4 = emigrants "
                                     "per step); particles were lost -- raise SphClusterOptions capacities",
                                     s.rank, s.dyn_host->overflow);
    }
    return 0;
}

// Kinetic energy of the owned particles (state after the last step) and the density sum of that
// step (its live particles, per sorted slot).
__global__ void __launch_bounds__(256)
    k_cluster_stats(const SlabDyn *dyn, const float4 *__restrict__ cur_pos, const float4 *__restrict__ cur_vel,
                    const float *__restrict__ rho, int slot0, double *out) {
    double ke = 0.0, rs = 0.0;
    const int n_total = dyn->n_total, n_prev = dyn->n_prev;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_total; i += gridDim.x * blockDim.x) {
        if (__float_as_uint(cur_pos[i].w) == 0xffffffffu) continue;   // emigrated
        const float4 v = cur_vel[i];
        ke += 0.5 * (double)kMass * ((double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z);
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_prev; i += gridDim.x * blockDim.x)
        rs += (double)rho[slot0 + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ke += __shfl_xor_sync(0xffffffffu, ke, o);
        rs += __shfl_xor_sync(0xffffffffu, rs, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out, ke);
        atomicAdd(out + 1, rs);
    }
}

}  // namespace

extern "C" {

int sph_cluster_nccl_id(uint8_t id[SPH_NCCL_ID_BYTES]) {
    if (!id) return sph_internal_fail(SPH_E_INVALID, "null argument");
    int rc = load_nccl();
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) <= SPH_NCCL_ID_BYTES, "ncclUniqueId grew");
    ncclUniqueId u;
    NC(g_nccl.GetUniqueId(&u));
    memset(id, 0, SPH_NCCL_ID_BYTES);
    memcpy(id, &u, sizeof(u));
    return 0;
}

int sph_cluster_create(const SphSettings *st, const SphClusterOptions *o, sph_cluster **out) {
    if (!st || !o || !out) return sph_internal_fail(SPH_E_INVALID, "null argument");
    *out = nullptr;
    if (o->world < 1 || o->local_count < 1 || o->local_count > SPH_MAX_LOCAL_SLABS || o->first_rank < 0 ||
        o->first_rank + o->local_count > o->world)
        return sph_internal_fail(SPH_E_INVALID, "bad slab layout: world %d, first_rank %d, local_count %d", o->world,
                                 o->first_rank, o->local_count);
    const int nc = (int)st->numCellsPerDim;
    const int nz = o->nz_cells > 0 ? o->nz_cells : nc;
    if (nz < 3 * o->world) return sph_internal_fail(SPH_E_INVALID, "%d cell layers cannot be split into %d slabs of >= 3 layers", nz, o->world);
    sph_cluster *c = new (std::nothrow) sph_cluster();
    if (!c) return sph_internal_fail(SPH_E_NOMEM, "out of host memory");
    c->st = *st;
    c->opt = *o;
    c->nz = nz;
    c->nc = nc;
    const long long share = ((long long)st->numParticles + o->world - 1) / o->world;
    c->cap = o->capacity > 0 ? o->capacity : (int)std::min<long long>(share + share / 4 + 65536, 0x7fffffff);
    c->cap_g = o->ghost_capacity > 0 ? o->ghost_capacity : c->cap / 8 + 1024;
    c->cap_g = (c->cap_g + 1) & ~1;
    c->cap_m = o->emig_capacity > 0 ? o->emig_capacity : c->cap / 32 + 1024;
    c->msg_bytes[kHaloA] = sizeof(MsgHeader) + (size_t)c->cap_g * 2 * sizeof(float4);
    c->msg_bytes[kHaloB] = sizeof(MsgHeader) + (size_t)c->cap_g * sizeof(float2);
    c->msg_bytes[kMigrate] = sizeof(MsgHeader) + (size_t)c->cap_m * 2 * sizeof(float4);
    split_layers(nz, o->world, c->zlo, c->zhi);
    c->max_layers.resize(o->world);
    for (int r = 0; r < o->world; ++r) c->max_layers[r] = c->zhi[r] - c->zlo[r] + kSlabSpareLayers;
    const bool need_nccl = o->local_count < o->world;
    int rc = need_nccl ? load_nccl() : 0;
    c->slabs.resize(o->local_count);
    for (int i = 0; i < o->local_count && rc == 0; ++i) {
        Slab &s = c->slabs[i];
        s.rank = o->first_rank + i;
        s.device = o->devices[i];
        SphSettings ss = *st;
        ss.numParticles = 0;
        SphOptions so;
        memset(&so, 0, sizeof so);
        so.device = s.device;
        so.key_mode = SPH_KEY_FLAT;
        so.use_graph = 2;
        so.capacity = c->cap;
        so.z_cell_lo = c->zlo[s.rank];
        so.z_cell_hi = c->zhi[s.rank];
        so.nz_cells = nz;
        so.ghost_capacity = c->cap_g;
        so.density_sum = o->density_sum;
        rc = sph_create_ex(&ss, &so, &s.sim);
        if (rc == 0) rc = sph_setup(s.sim);
        if (rc == 0) rc = sph_internal_core(s.sim, &s.core);
        if (rc) break;
        s.zlo = so.z_cell_lo;
        s.zhi = so.z_cell_hi;
        auto fail_cuda = [&](cudaError_t e) { if (e != cudaSuccess && rc == 0) rc = sph_internal_fail((int)e, "cluster allocation failed: %s", cudaGetErrorString(e)); };
        fail_cuda(cudaSetDevice(s.device));
        fail_cuda(cudaMalloc(&s.dyn, sizeof(SlabDyn)));
        fail_cuda(cudaMemset(s.dyn, 0, sizeof(SlabDyn)));
        fail_cuda(cudaMallocHost(&s.dyn_host, sizeof(SlabDyn)));
        if (rc == 0) memset(s.dyn_host, 0, sizeof(SlabDyn));
        fail_cuda(cudaMalloc(&s.stats_dev, 2 * sizeof(double)));
        fail_cuda(cudaMalloc(&s.info_dev, 24 * sizeof(int)));
        for (int k = 0; k < kMsgKinds && rc == 0; ++k) {
            for (int side = 0; side < 2 && rc == 0; ++side) {
                rc = alloc_msg(&s.send[k][side], c->msg_bytes[k]);
                if (rc == 0) rc = alloc_msg(&s.recv[k][side], c->msg_bytes[k]);
                fail_cuda(cudaEventCreateWithFlags(&s.ev_copied[k][side], cudaEventDisableTiming));
            }
            fail_cuda(cudaEventCreateWithFlags(&s.ev_packed[k], cudaEventDisableTiming));
        }
        fail_cuda(cudaEventCreate(&s.ev_t0));
        fail_cuda(cudaEventCreate(&s.ev_t1));
        if (rc) break;
        // the force kernel appends emigrants straight into the migration messages
        DeviceState &d = *s.core.d;
        for (int side = 0; side < 2; ++side) {
            d.emig_pos[side] = reinterpret_cast<float4 *>(s.send[kMigrate][side] + 1);
            d.emig_vel[side] = d.emig_pos[side] + c->cap_m;
            d.emig_count[side] = &s.send[kMigrate][side]->count;
        }
        d.emig_capacity = c->cap_m;
        if (need_nccl) {
            ncclUniqueId u;
            memcpy(&u, o->nccl_id, sizeof(u));
            ncclResult_t r = g_nccl.CommInitRank(&s.comm, o->world, u, s.rank);
            if (r != ncclSuccess) rc = sph_internal_fail(SPH_E_STATE, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
        }
    }
    // Messages between processes: map the neighbours' send buffers (CUDA IPC) so that the unpack
    // kernels read them in place.  The handles travel over the NCCL communicator.
    const char *tr = getenv("SPH_CLUSTER_TRACE");
    c->trace = tr && atoi(tr) != 0;
    {
        const char *f = getenv("SPH_FUSE_COUNT");
        c->fuse_count = !c->slabs.empty() && c->slabs[0].core.cell_sort && !(f && !strcmp(f, "0"));
    }
    const char *force_nccl = getenv("SPH_CLUSTER_NCCL_DATA");
    c->peer_mem = need_nccl && !(force_nccl && atoi(force_nccl) != 0);
    for (size_t i = 0; i < c->slabs.size() && rc == 0 && c->peer_mem; ++i) {
        Slab &s = c->slabs[i];
        struct Handles { cudaIpcMemHandle_t h[kMsgKinds]; };
        Handles mine[2], theirs[2];
        Handles *dev = nullptr;
        auto fail_cuda = [&](cudaError_t e) { if (e != cudaSuccess && rc == 0) rc = sph_internal_fail((int)e, "CUDA IPC set-up failed: %s", cudaGetErrorString(e)); };
        fail_cuda(cudaSetDevice(s.device));
        fail_cuda(cudaMalloc(&dev, 4 * sizeof(Handles)));
        bool any = false;
        for (int side = 0; side < 2 && rc == 0; ++side) {
            const int peer = side ? s.rank + 1 : s.rank - 1;
            if (peer < 0 || peer >= o->world || is_local(c, peer)) continue;
            for (int k = 0; k < kMsgKinds; ++k) fail_cuda(cudaIpcGetMemHandle(&mine[side].h[k], s.send[k][side]));
            any = true;
        }
        if (rc == 0 && any) {
            fail_cuda(cudaMemcpy(dev, mine, 2 * sizeof(Handles), cudaMemcpyHostToDevice));
            ncclResult_t r = g_nccl.GroupStart();
            for (int side = 0; side < 2 && r == ncclSuccess; ++side) {
                const int peer = side ? s.rank + 1 : s.rank - 1;
                if (peer < 0 || peer >= o->world || is_local(c, peer)) continue;
                r = g_nccl.Send(dev + side, sizeof(Handles), ncclChar, peer, s.comm, s.core.stream);
                if (r == ncclSuccess) r = g_nccl.Recv(dev + 2 + side, sizeof(Handles), ncclChar, peer, s.comm, s.core.stream);
            }
            if (r == ncclSuccess) r = g_nccl.GroupEnd();
            if (r != ncclSuccess) rc = sph_internal_fail(SPH_E_STATE, "exchange of the IPC handles failed: %s", g_nccl.GetErrorString(r));
            fail_cuda(cudaStreamSynchronize(s.core.stream));
            fail_cuda(cudaMemcpy(theirs, dev + 2, 2 * sizeof(Handles), cudaMemcpyDeviceToHost));
            for (int side = 0; side < 2 && rc == 0; ++side) {
                const int peer = side ? s.rank + 1 : s.rank - 1;
                if (peer < 0 || peer >= o->world || is_local(c, peer)) continue;
                for (int k = 0; k < kMsgKinds; ++k)
                    fail_cuda(cudaIpcOpenMemHandle((void **)&s.remote[k][side], theirs[side].h[k], cudaIpcMemLazyEnablePeerAccess));
            }
        }
        cudaFree(dev);
    }
    // peer access between the devices of local neighbours (NVLink P2P); a failure only means the
    // copies are staged by the driver
    for (size_t i = 0; i + 1 < c->slabs.size() && rc == 0; ++i) {
        const int a = c->slabs[i].device, b = c->slabs[i + 1].device;
        if (a == b) continue;
        int ok = 0;
        if (cudaDeviceCanAccessPeer(&ok, a, b) == cudaSuccess && ok) {
            cudaSetDevice(a);
            if (cudaDeviceEnablePeerAccess(b, 0) != cudaSuccess) cudaGetLastError();
            cudaSetDevice(b);
            if (cudaDeviceEnablePeerAccess(a, 0) != cudaSuccess) cudaGetLastError();
        }
    }
    if (rc) {
        sph_cluster_destroy(c);
        return rc;
    }
    *out = c;
    return 0;
}

void sph_cluster_destroy(sph_cluster *c) {
    if (!c) return;
    // Peer memory: a neighbour in another process may still be reading this slab's buffers (and this
    // slab the neighbour's).  Drain the own streams, then swap one word with every remote neighbour
    // -- it sends only after ITS streams have drained -- before anything is unmapped or freed.
    if (c->peer_mem) {
        for (Slab &s : c->slabs) {
            cudaSetDevice(s.device);
            if (s.core.stream) cudaStreamSynchronize(s.core.stream);
        }
        for (Slab &s : c->slabs) {
            if (!s.comm || !s.info_dev) continue;
            cudaSetDevice(s.device);
            bool grouped = false;
            for (int side = 0; side < 2; ++side) {
                const int peer = side ? s.rank + 1 : s.rank - 1;
                if (peer < 0 || peer >= c->opt.world || is_local(c, peer)) continue;
                if (!grouped) { g_nccl.GroupStart(); grouped = true; }
                g_nccl.Send(s.info_dev, 4, ncclChar, peer, s.comm, s.core.stream);
                g_nccl.Recv(s.info_dev + 8 * (1 + side), 4, ncclChar, peer, s.comm, s.core.stream);
            }
            if (grouped) {
                g_nccl.GroupEnd();
                cudaStreamSynchronize(s.core.stream);
            }
        }
    }
    if (c->trace)
        for (Slab &s : c->slabs) {
            const int steps = std::max(1, s.dyn_host ? s.dyn_host->steps : 1);
            fprintf(stderr, "[sph cluster trace] slab %d, ms per step over %d steps:", s.rank, steps);
            for (int ph = 0; ph < kPhases; ++ph) fprintf(stderr, " %s %.3f;", kPhaseNames[ph], s.phase_ms[ph] / steps);
            fprintf(stderr, "\n");
            for (cudaEvent_t e : s.trace_ev) cudaEventDestroy(e);
        }
    for (Slab &s : c->slabs) {
        cudaSetDevice(s.device);
        if (s.core.stream) cudaStreamSynchronize(s.core.stream);
        if (s.copy_stream) {
            cudaStreamSynchronize(s.copy_stream);
            cudaStreamDestroy(s.copy_stream);
        }
        if (s.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(s.comm);
        if (s.sim) {
            // the emigrant pointers alias the migration messages, which are freed below
            DeviceState *d = s.core.d;
            if (d) {
                for (int side = 0; side < 2; ++side) d->emig_pos[side] = d->emig_vel[side] = nullptr;
                d->emig_count[0] = d->emig_count[1] = nullptr;
            }
        }
        for (int k = 0; k < kMsgKinds; ++k) {
            for (int side = 0; side < 2; ++side) {
                if (s.remote[k][side]) cudaIpcCloseMemHandle(s.remote[k][side]);
                cudaFree(s.send[k][side]);
                cudaFree(s.recv[k][side]);
                if (s.ev_copied[k][side]) cudaEventDestroy(s.ev_copied[k][side]);
            }
            if (s.ev_packed[k]) cudaEventDestroy(s.ev_packed[k]);
        }
        if (s.ev_t0) cudaEventDestroy(s.ev_t0);
        if (s.ev_t1) cudaEventDestroy(s.ev_t1);
        if (s.ev_out_ready) cudaEventDestroy(s.ev_out_ready);
        if (s.ev_out_done) cudaEventDestroy(s.ev_out_done);
        cudaFree(s.dyn);
        cudaFree(s.stats_dev);
        cudaFree(s.info_dev);
        cudaFree(s.out_stage);
        if (s.dyn_host) cudaFreeHost(s.dyn_host);
        if (s.host_records) cudaFreeHost(s.host_records);
        if (s.sim) sph_destroy(s.sim);
    }
    delete c;
}

int sph_cluster_load(sph_cluster *c, int li, int n, const float *pos, const float *vel, const uint32_t *ids) {
    if (!c || li < 0 || li >= (int)c->slabs.size()) return sph_internal_fail(SPH_E_INVALID, "bad slab index");
    Slab &s = c->slabs[li];
    int rc = sph_slab_load(s.sim, n, pos, vel, ids);
    if (rc) return rc;
    CU(cudaSetDevice(s.device));
    SlabDyn h;
    memset(&h, 0, sizeof h);
    h.n_total = h.n_live = n;
    CU(cudaMemcpy(s.dyn, &h, sizeof h, cudaMemcpyHostToDevice));
    *s.dyn_host = h;
    s.known_total = n;
    s.hashed = false;
    if (s.counted) { s.counted = false; s.count_dirty = true; }
    for (int k = 0; k < kMsgKinds; ++k)
        for (int side = 0; side < 2; ++side) {   // (counts only: seq / ack follow the cluster's rounds)
            CU(cudaMemset(&s.send[k][side]->count, 0, sizeof(uint32_t)));
            CU(cudaMemset(&s.recv[k][side]->count, 0, sizeof(uint32_t)));
        }
    c->loaded = true;
    return 0;
}

// ref: simulator.cu:430-453 -- the same particle set a single simulator starts from, split by layer
int sph_cluster_setup(sph_cluster *c) {
    if (!c) return sph_internal_fail(SPH_E_INVALID, "null cluster");
    const SphSettings &st = c->st;
    const int n = st.numParticles;
    if (c->nz != c->nc) return sph_internal_fail(SPH_E_INVALID, "sph_cluster_setup() initialises the reference's cubic box; use sph_cluster_load() for nz_cells != numCellsPerDim");
    std::vector<float> pos((size_t)3 * std::max(n, 1));
    if (st.randomInit) {
        for (int i = 0; i < n; ++i)
            for (int a = 0; a < 3; ++a) pos[3 * (size_t)i + a] = rand() / (float)RAND_MAX * (st.boxDim - 2.f) + 1.f;
    } else {
        const float spacing = 0.9f * st.h;
        const int nx = (int)(floorf((st.boxDim - 2 * st.h) / spacing) + 1);
        if ((long long)n > (long long)nx * nx * nx)
            return sph_internal_fail(SPH_E_INVALID, "grid init: the %d^3 lattice of a boxDim=%g box cannot hold %d particles", nx, (double)st.boxDim, n);
        int count = 0;
        for (int x = 0; x < nx && count < n; ++x)
            for (int y = 0; y < nx && count < n; ++y)
                for (int z = 0; z < nx && count < n; ++z) {
                    pos[3 * (size_t)count] = st.h + spacing * x;
                    pos[3 * (size_t)count + 1] = st.h + spacing * y;
                    pos[3 * (size_t)count + 2] = st.h + spacing * z;
                    ++count;
                }
    }
    for (size_t li = 0; li < c->slabs.size(); ++li) {
        Slab &s = c->slabs[li];
        std::vector<float> mp;
        std::vector<uint32_t> mi;
        for (int i = 0; i < n; ++i) {
            int cz = (int)(pos[3 * (size_t)i + 2] / st.h);   // IEEE divide, truncate (ref: simulator.cu:69)
            cz = std::min(std::max(cz, 0), c->nz - 1);
            if (cz >= s.zlo && cz < s.zhi) {
                mp.insert(mp.end(), &pos[3 * (size_t)i], &pos[3 * (size_t)i] + 3);
                mi.push_back((uint32_t)i);
            }
        }
        if ((int)mi.size() > c->cap)
            return sph_internal_fail(SPH_E_INVALID, "slab %d would own %zu particles, capacity %d", s.rank, mi.size(), c->cap);
        int rc = sph_cluster_load(c, (int)li, (int)mi.size(), mp.data(), nullptr, mi.data());
        if (rc) return rc;
    }
    return 0;
}

int sph_cluster_sync(sph_cluster *c) {
    if (!c) return sph_internal_fail(SPH_E_INVALID, "null cluster");
    return sync_all(c);
}

int sph_cluster_advance(sph_cluster *c, int steps) {
    if (!c || steps < 0) return sph_internal_fail(SPH_E_INVALID, "bad argument");
    if (!c->loaded) return sph_internal_fail(SPH_E_STATE, "no particles: call sph_cluster_setup() or sph_cluster_load()");
    for (int k = 0; k < steps; ++k) {
        if (c->opt.rebalance_every > 0 && c->steps_since_rebalance >= c->opt.rebalance_every) {
            int rc = sph_cluster_rebalance(c);
            if (rc) return rc;
        }
        int rc = enqueue_step(c);
        if (rc) return rc;
    }
    return sync_all(c);
}

int sph_cluster_advance_timed(sph_cluster *c, int steps, float *ms) {
    if (!c || steps < 0 || !ms) return sph_internal_fail(SPH_E_INVALID, "bad argument");
    if (!c->loaded) return sph_internal_fail(SPH_E_STATE, "no particles: call sph_cluster_setup() or sph_cluster_load()");
    for (Slab &s : c->slabs) {
        CU(cudaSetDevice(s.device));
        CU(cudaEventRecord(s.ev_t0, s.core.stream));
    }
    for (int k = 0; k < steps; ++k) {
        if (c->opt.rebalance_every > 0 && c->steps_since_rebalance >= c->opt.rebalance_every) {
            int rc = sph_cluster_rebalance(c);
            if (rc) return rc;
        }
        int rc = enqueue_step(c);
        if (rc) return rc;
    }
    for (Slab &s : c->slabs) {
        CU(cudaSetDevice(s.device));
        CU(cudaEventRecord(s.ev_t1, s.core.stream));
    }
    int rc = sync_all(c);
    if (rc) return rc;
    *ms = 0.f;
    for (Slab &s : c->slabs) {
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, s.ev_t0, s.ev_t1));
        *ms = std::max(*ms, t);
    }
    return 0;
}

int sph_cluster_step(sph_cluster *c) {
    if (!c) return sph_internal_fail(SPH_E_INVALID, "null cluster");
    if (!c->loaded) return sph_internal_fail(SPH_E_STATE, "no particles: call sph_cluster_setup() or sph_cluster_load()");
    int rc = enqueue_step(c);
    if (rc) return rc;
    for (Slab &s : c->slabs) {
        CU(cudaSetDevice(s.device));
        if (!s.out_stage) {
            CU(cudaMalloc(&s.out_stage, (size_t)c->cap * sizeof(float4)));
            if (!s.host_records) CU(cudaMallocHost(&s.host_records, (size_t)c->cap * sizeof(float4)));
            CU(cudaStreamCreateWithFlags(&s.copy_stream, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&s.ev_out_ready, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&s.ev_out_done, cudaEventDisableTiming));
        }
        // the host's bound on the slab's particle count: the last count it has seen plus what one
        // step can bring in; the records beyond the true count are stale and carry no meaning
        const int seen = std::max(s.known_total, s.dyn_host->n_total);
        const int bound = std::min(c->cap, seen + 2 * c->cap_m);
        if (s.out_pending) CU(cudaStreamWaitEvent(s.core.stream, s.ev_out_done, 0));   // staging buffer free again
        CU(cudaMemcpyAsync(s.out_stage, s.core.d->cur_pos, (size_t)bound * sizeof(float4), cudaMemcpyDeviceToDevice,
                           s.core.stream));
        CU(cudaEventRecord(s.ev_out_ready, s.core.stream));
        CU(cudaStreamWaitEvent(s.copy_stream, s.ev_out_ready, 0));
        CU(cudaMemcpyAsync(s.host_records, s.out_stage, (size_t)bound * sizeof(float4), cudaMemcpyDeviceToHost,
                           s.copy_stream));
        CU(cudaEventRecord(s.ev_out_done, s.copy_stream));
        s.out_pending = true;
        s.out_count = bound;
    }
    return 0;
}

int sph_cluster_host_records(sph_cluster *c, int li, const float **records, int *count) {
    if (!c || li < 0 || li >= (int)c->slabs.size() || !records || !count) return sph_internal_fail(SPH_E_INVALID, "bad argument");
    Slab &s = c->slabs[li];
    *records = reinterpret_cast<const float *>(s.host_records);
    *count = std::min(s.out_count, std::max(s.dyn_host->n_total, 0));
    return 0;
}

int sph_cluster_download(sph_cluster *c, int li, uint32_t *ids, float *pos, float *vel, int *n_out) {
    if (!c || li < 0 || li >= (int)c->slabs.size()) return sph_internal_fail(SPH_E_INVALID, "bad slab index");
    int rc = sync_all(c);
    if (rc) return rc;
    Slab &s = c->slabs[li];
    CU(cudaSetDevice(s.device));
    const int n = s.dyn_host->n_total;
    std::vector<float4> hp((size_t)std::max(n, 1)), hv((size_t)std::max(n, 1));
    if (n) {
        CU(cudaMemcpy(hp.data(), s.core.d->cur_pos, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(hv.data(), s.core.d->cur_vel, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost));
    }
    int m = 0;
    for (int i = 0; i < n; ++i) {
        uint32_t id;
        memcpy(&id, &hp[i].w, 4);
        if (id == 0xffffffffu) continue;   // emigrated
        if (ids) ids[m] = id;
        if (pos) { pos[3 * m] = hp[i].x; pos[3 * m + 1] = hp[i].y; pos[3 * m + 2] = hp[i].z; }
        if (vel) { vel[3 * m] = hv[i].x; vel[3 * m + 1] = hv[i].y; vel[3 * m + 2] = hv[i].z; }
        ++m;
    }
    if (n_out) *n_out = m;
    return 0;
}

// getPosition() of the multi-GPU path: every local slab's records {x, y, z, id} come to its pinned
// host buffer (one asynchronous copy per slab, all in flight together), then one host thread per
// slab scatters them into out[3 * id ...] (ids are disjoint between slabs).
int sph_cluster_positions(sph_cluster *c, float *out, int64_t n_global) {
    if (!c || !out) return sph_internal_fail(SPH_E_INVALID, "bad argument");
    int rc = sync_all(c);
    if (rc) return rc;
    for (Slab &s : c->slabs) {
        CU(cudaSetDevice(s.device));
        if (!s.host_records) CU(cudaMallocHost(&s.host_records, (size_t)c->cap * sizeof(float4)));
        const int n = s.dyn_host->n_total;
        if (n) CU(cudaMemcpyAsync(s.host_records, s.core.d->cur_pos, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost, s.core.stream));
    }
    std::vector<std::thread> workers;
    std::vector<long long> bad(c->slabs.size(), -1);
    for (size_t i = 0; i < c->slabs.size(); ++i) {
        Slab &s = c->slabs[i];
        CU(cudaSetDevice(s.device));
        CU(cudaStreamSynchronize(s.core.stream));
        workers.emplace_back([&s, &bad, i, out, n_global]() {
            const int n = s.dyn_host->n_total;
            const float4 *hp = s.host_records;
            for (int k = 0; k < n; ++k) {
                uint32_t id;
                memcpy(&id, &hp[k].w, 4);
                if (id == 0xffffffffu) continue;   // emigrated
                if ((int64_t)id >= n_global) { bad[i] = id; return; }
                out[3 * (size_t)id] = hp[k].x;
                out[3 * (size_t)id + 1] = hp[k].y;
                out[3 * (size_t)id + 2] = hp[k].z;
            }
        });
    }
    for (auto &w : workers) w.join();
    for (size_t i = 0; i < bad.size(); ++i)
        if (bad[i] >= 0) return sph_internal_fail(SPH_E_STATE, "slab %d holds id %lld >= %lld", c->slabs[i].rank, bad[i], (long long)n_global);
    return 0;
}

int sph_cluster_stats(sph_cluster *c, int li, SphSlabStats *out) {
    if (!c || li < 0 || li >= (int)c->slabs.size() || !out) return sph_internal_fail(SPH_E_INVALID, "bad argument");
    int rc = sync_all(c);
    if (rc) return rc;
    Slab &s = c->slabs[li];
    CU(cudaSetDevice(s.device));
    memset(out, 0, sizeof *out);
    CU(cudaMemcpy(s.dyn_host, s.dyn, sizeof(SlabDyn), cudaMemcpyDeviceToHost));
    const SlabDyn &h = *s.dyn_host;
    out->rank = s.rank;
    out->device = s.device;
    out->z_cell_lo = s.zlo;
    out->z_cell_hi = s.zhi;
    out->n_owned = h.n_live;
    out->ghosts_lo = h.g_lo;
    out->ghosts_hi = h.g_hi;
    out->steps = h.steps;
    out->migrated_total = (int64_t)h.migrated;
    out->ghosts_total = (int64_t)h.ghosts;
    out->overflow = h.overflow;
    out->rebalances = s.rebalances;
    double hs[2] = {0, 0};
    CU(cudaMemsetAsync(s.stats_dev, 0, 2 * sizeof(double), s.core.stream));
    k_cluster_stats<<<148 * 4, 256, 0, s.core.stream>>>(s.dyn, s.core.d->cur_pos, s.core.d->cur_vel, s.core.d->rho,
                                                        s.core.p->slot0, s.stats_dev);
    CU(cudaMemcpyAsync(hs, s.stats_dev, sizeof hs, cudaMemcpyDeviceToHost, s.core.stream));
    CU(cudaStreamSynchronize(s.core.stream));
    out->kinetic_energy = hs[0];
    out->density_sum = hs[1];
    return sph_debug_flags(s.sim, &out->debug_flags, &out->checked_build);
}

int64_t sph_cluster_launch_count(sph_cluster *c) { return c ? c->launches : 0; }

// ---- load rebalancing ---------------------------------------------------------------------------
// Every boundary between two slabs moves by at most one layer towards the lighter slab, when that
// reduces the difference of their particle counts.  A decision needs the two slabs' numbers only,
// so neighbours swap five integers (the same point-to-point channels as the halos; no collective)
// and both sides reach the same verdict.  The particles of a layer that changes owner are sent
// through the ordinary migration messages: k_rekey_emigrate marks everything outside the new range
// as emigrated and re-keys the rest for the new local layer numbering.
struct SlabInfo5 {
    int n_live, lo_count, hi_count, layers, max_layers;
};

static int boundary_move(const SlabInfo5 &below, const SlabInfo5 &above) {
    const long long a = below.n_live, b = above.n_live;
    // -1: the slab below hands its top layer up; +1: the slab above hands its bottom layer down.
    // A slab keeps >= 3 layers and stays inside its cell table even if both its faces move.
    if (a - b > below.hi_count && below.hi_count > 0 && below.layers >= 5 && above.layers <= above.max_layers - 2) return -1;
    if (b - a > above.lo_count && above.lo_count > 0 && above.layers >= 5 && below.layers <= below.max_layers - 2) return +1;
    return 0;
}

int sph_cluster_rebalance(sph_cluster *c) {
    if (!c) return sph_internal_fail(SPH_E_INVALID, "null cluster");
    c->steps_since_rebalance = 0;
    const int W = c->opt.world;
    if (W == 1) return 0;
    int rc = sync_all(c);   // the pinned SlabDyn mirrors are those of the last step now
    if (rc) return rc;
    const size_t L = c->slabs.size();
    std::vector<SlabInfo5> mine(L), nb_lo(L), nb_hi(L);
    for (size_t i = 0; i < L; ++i) {
        Slab &s = c->slabs[i];
        const SlabDyn &h = *s.dyn_host;
        mine[i] = SlabInfo5{h.n_live, h.steps ? h.lo_count : 0, h.steps ? h.hi_count : 0, s.zhi - s.zlo,
                            c->max_layers[s.rank]};
    }
    // neighbours' numbers: local ones directly, remote ones over send / recv
    for (size_t i = 0; i < L; ++i) {
        Slab &s = c->slabs[i];
        CU(cudaSetDevice(s.device));
        bool grouped = false;
        for (int side = 0; side < 2; ++side) {
            const int peer = side ? s.rank + 1 : s.rank - 1;
            if (peer < 0 || peer >= W) continue;
            if (is_local(c, peer)) {
                (side ? nb_hi : nb_lo)[i] = mine[peer - c->opt.first_rank];
            } else {
                if (!grouped) {
                    CU(cudaMemcpyAsync(s.info_dev, &mine[i], sizeof(SlabInfo5), cudaMemcpyHostToDevice, s.core.stream));
                    NC(g_nccl.GroupStart());
                    grouped = true;
                }
                NC(g_nccl.Send(s.info_dev, sizeof(SlabInfo5), ncclChar, peer, s.comm, s.core.stream));
                NC(g_nccl.Recv(s.info_dev + 8 * (1 + side), sizeof(SlabInfo5), ncclChar, peer, s.comm, s.core.stream));
            }
        }
        if (grouped) {
            NC(g_nccl.GroupEnd());
            int host[24];
            CU(cudaMemcpyAsync(host, s.info_dev, sizeof host, cudaMemcpyDeviceToHost, s.core.stream));
            CU(cudaStreamSynchronize(s.core.stream));
            if (s.rank > 0 && !is_local(c, s.rank - 1)) memcpy(&nb_lo[i], host + 8, sizeof(SlabInfo5));
            if (s.rank + 1 < W && !is_local(c, s.rank + 1)) memcpy(&nb_hi[i], host + 16, sizeof(SlabInfo5));
        }
    }
    bool any = false;
    std::vector<int> new_lo(L), new_hi(L);
    for (size_t i = 0; i < L; ++i) {
        Slab &s = c->slabs[i];
        new_lo[i] = s.zlo + (s.rank > 0 ? boundary_move(nb_lo[i], mine[i]) : 0);
        new_hi[i] = s.zhi + (s.rank + 1 < W ? boundary_move(mine[i], nb_hi[i]) : 0);
        any = any || new_lo[i] != s.zlo || new_hi[i] != s.zhi;
    }
    // Whether ANY boundary of the job moved is not known here (a process only sees its own faces),
    // so the migration round below always runs when slabs are remote; it is skipped only when every
    // slab is local and nothing moved.
    if (!any && c->opt.local_count == W) return 0;
    for (size_t i = 0; i < L; ++i) {
        Slab &s = c->slabs[i];
        CU(cudaSetDevice(s.device));
        const bool moved = new_lo[i] != s.zlo || new_hi[i] != s.zhi;
        if (moved) {
            apply_range(s, new_lo[i], new_hi[i]);
            c->zlo[s.rank] = new_lo[i];
            c->zhi[s.rank] = new_hi[i];
            s.rebalances += 1;
        }
        if (s.counted) {   // keys change / particles arrive outside a step: the fused counts are void
            s.counted = false;
            s.count_dirty = true;
        }
        rc = wait_taken(c, s, kMigrate);
        if (rc) return rc;
        before_pack(c, s, kMigrate);
        for (int side = 0; side < 2; ++side) CU(cudaMemsetAsync(&s.send[kMigrate][side]->count, 0, sizeof(uint32_t), s.core.stream));
        if (moved) {
            const Params p = launch_params(s, c->cap);
            launch_rekey_emigrate(p, *s.core.d, s.core.stream);
            c->launches += 1;
        }
        after_pack(c, s, kMigrate);
    }
    const uint32_t round_m = c->round[kMigrate];
    rc = exchange(c, kMigrate);
    if (rc) return rc;
    for (Slab &s : c->slabs) {
        CU(cudaSetDevice(s.device));
        const Params p = launch_params(s, c->cap);
        // (rebalance = true: append behind n_total and keep the earlier dead entries counted)
        before_unpack(c, s, kMigrate, round_m);
        launch_append_immigrants(p, *s.core.d, in_msg(c, s, kMigrate, 0), in_msg(c, s, kMigrate, 1), c->cap_m,
                                 s.send[kMigrate][0], s.send[kMigrate][1], c->cap, s.dyn, true, s.core.stream);
        after_unpack(c, s, kMigrate, round_m);
        CU(cudaMemcpyAsync(s.dyn_host, s.dyn, sizeof(SlabDyn), cudaMemcpyDeviceToHost, s.core.stream));
        c->launches += 2;
    }
    return sync_all(c);
}

}  // extern "C"
