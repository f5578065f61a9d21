// sph_api.cu -- host step driver and C ABI (include/sph_b200.h) of libsph_b200.so.
//
// Replaces the host half of the reference's simulator.cu (ref: src/simulator.cu:370-546:
// ctor/dtor, setup(), getPosition(), simulate(), simulateAndTime()).  One stream, the
// whole step replayed as a CUDA graph, one synchronisation per step (the reference
// synchronises the device three to four times per step and clears its grid with 10^6
// one-thread blocks).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/sph_b200.h"
#include "sph_internal.cuh"
#include "sph_kernels.cuh"
#include "sph_sort.cuh"

using namespace sph;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(expr)                                                                          \
    do {                                                                                  \
        cudaError_t e__ = (expr);                                                         \
        if (e__ != cudaSuccess)                                                           \
            return fail((int)e__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                              \
    } while (0)

constexpr int kDefaultDensityExact = 0;   // SphOptions.density_sum == 0: the factored sum

enum Stage { kStHash = 0, kStHist, kStSort, kStReorder, kStDensity, kStForce, kStPush, kStOther };
const char *kStageNames[SPH_STAGE_COUNT] = {"hash",    "histogram",       "sort_passes", "reorder_cellstart",
                                            "density", "force_integrate", "push",        "other"};

struct EventPair {
    cudaEvent_t a, b;
    int stage;
};

}  // namespace

struct sph_sim {
    SphSettings settings;
    SphOptions opt;
    Params p{};   // zero: no slab, no CTA gap (particle_cta() is the identity)
    Thresholds th;
    DeviceState d;
    int sm_count = 148;
    int capacity = 0;
    int passes = 0;
    bool cell_sort = false;     // counting sort by cell instead of the radix passes (slabs too: SlabCore)
    bool fuse_count = false;    // single GPU: ... with the count produced by the previous step's force kernel
    bool counted = false;       // d.cell_count / d.pairs[1] describe the current state (fused count ran);
                                // the count table is non-zero exactly while this is set
    int sorted_buf = 0;
    cudaStream_t stream = nullptr;
    float *host_pos = nullptr;  // pinned, 3*n floats, original order
    bool is_setup = false;
    bool own_stream = true;
    // slab mode
    int ghost_cap = 0;          // == p.slot0
    uint32_t table_capacity = 0;   // cell_start entries allocated (slab: room for the layer range to grow)
    bool keys_valid = false;    // d.key matches d.cur_pos
    bool step_valid = false;    // srt_*/cell_start/rho/pa describe the last step
    cudaGraphExec_t graph = nullptr;       // step graph writing out_buf[0]
    cudaGraphExec_t graph_alt = nullptr;   // same step writing out_buf[1] (pipelined readback)
    cudaGraphExec_t graph_noout = nullptr; // same step without the id-ordered position copy (sph_advance)
    bool out_stale = false;                // out_buf[0] is older than the state (see k_unpermute)
    cudaGraphExec_t graph_build = nullptr, graph_update = nullptr;   // the two halves, for the timed step
    float *out_buf[2] = {nullptr, nullptr};
    int out_parity = 0;
    bool spec_inflight = false;            // a step is enqueued whose positions were not handed out yet
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_step = nullptr, ev_copy = nullptr;
    bool profiling = false;
    std::vector<EventPair> events;
    size_t events_used = 0;
    double stage_ms[SPH_STAGE_COUNT] = {0};
    int64_t stage_launches[SPH_STAGE_COUNT] = {0};
    int64_t launches = 0;
};

namespace {

// ---- instrumentation --------------------------------------------------------------
void stage_begin(sph_sim *s, int stage) {
    s->launches++;
    s->stage_launches[stage]++;
    if (!s->profiling) return;
    if (s->events_used == s->events.size()) {
        EventPair ep;
        cudaEventCreate(&ep.a);
        cudaEventCreate(&ep.b);
        s->events.push_back(ep);
    }
    EventPair &ep = s->events[s->events_used];
    ep.stage = stage;
    cudaEventRecord(ep.a, s->stream);
}
void stage_end(sph_sim *s) {
    if (!s->profiling) return;
    cudaEventRecord(s->events[s->events_used].b, s->stream);
    s->events_used++;
}
void profile_collect(sph_sim *s) {  // call after the stream is synchronised
    for (size_t i = 0; i < s->events_used; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s->events[i].a, s->events[i].b) == cudaSuccess)
            s->stage_ms[s->events[i].stage] += ms;
    }
    s->events_used = 0;
}
void sort_before(void *ctx, int st) {
    stage_begin((sph_sim *)ctx, st == kSortStageHistogram ? kStHist : kStSort);
}
void sort_after(void *ctx, int) { stage_end((sph_sim *)ctx); }

// ---- step pieces --------------------------------------------------------------------
// "build": everything the reference's "Grid construction" bucket covers.
void enqueue_build(sph_sim *s) {
    if (!s->keys_valid) {
        stage_begin(s, kStHash);
        launch_hash(s->p, s->d, s->stream);
        stage_end(s);
        s->keys_valid = true;
        if (s->counted) {   // counts of a state that was replaced
            cudaMemsetAsync(s->d.cell_count, 0, ((size_t)s->p.table_size + 1) * sizeof(uint32_t), s->stream);
            s->counted = false;
        }
    }
    SortHooks hooks{s, sort_before, sort_after};
    if (s->cell_sort) {
        cell_sort_async(s->d.key, s->d.pairs[0], s->d.pairs[1], s->p.n, s->p.table_size + 1u, s->d.cell_count,
                        s->d.cell_start, s->d.sort_scratch, s->stream, &hooks, s->counted);
        s->counted = false;
        s->sorted_buf = 0;
        s->d.sorted_pairs = s->d.pairs[0];
        stage_begin(s, kStReorder);
        launch_reorder_counted(s->p, s->d, 0, s->stream);
        stage_end(s);
        return;
    }
    s->sorted_buf = sort_pairs_async(s->d.key, s->d.pairs[0], s->d.pairs[1], s->p.n, s->passes,
                                     s->d.sort_scratch, s->sm_count, s->stream, &hooks);
    s->d.sorted_pairs = s->d.pairs[s->sorted_buf];
    stage_begin(s, kStReorder);
    launch_reorder(s->p, s->d, s->sorted_buf, s->p.n, s->sm_count, s->stream);
    stage_end(s);
}
// "update": the reference's "SPH update" bucket.
void enqueue_update(sph_sim *s) {
    stage_begin(s, kStDensity);
    launch_density(s->p, s->th, s->d, false, s->stream);
    stage_end(s);
    stage_begin(s, kStForce);
    launch_force_integrate(s->p, s->th, s->d, s->stream, s->fuse_count);
    stage_end(s);
    s->counted = s->fuse_count;
}

int sort_launches(const sph_sim *s) { return s->cell_sort ? 3 /*scan sums, scan, scatter*/ : s->passes; }
int hist_launches(const sph_sim *s) { return s->fuse_count ? 0 : 1; }   // histogram / count kernel
int graph_launches_per_step(const sph_sim *s) { return hist_launches(s) + sort_launches(s) + 3; }

// One timestep on the stream; graph replay when possible.
int enqueue_step(sph_sim *s) {
    const bool want_graph = s->opt.use_graph != 2 && !s->profiling;
    if (want_graph && s->keys_valid && s->counted == s->fuse_count) {
        cudaGraphExec_t &graph_slot = !s->d.out_pos ? s->graph_noout : (s->out_parity ? s->graph_alt : s->graph);
        if (!graph_slot) {
            cudaGraph_t g = nullptr;
            const int64_t l0 = s->launches;
            int64_t sl0[SPH_STAGE_COUNT];
            memcpy(sl0, s->stage_launches, sizeof(sl0));
            CU(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
            enqueue_build(s);
            enqueue_update(s);
            CU(cudaStreamEndCapture(s->stream, &g));
            s->launches = l0;  // capture launches nothing
            memcpy(s->stage_launches, sl0, sizeof(sl0));
            CU(cudaGraphInstantiate(&graph_slot, g, 0));
            cudaGraphDestroy(g);
        }
        CU(cudaGraphLaunch(graph_slot, s->stream));
        s->launches += graph_launches_per_step(s);
        s->stage_launches[kStHist] += hist_launches(s);
        s->stage_launches[kStSort] += sort_launches(s);
        s->stage_launches[kStReorder] += 1;
        s->stage_launches[kStDensity] += 1;
        s->stage_launches[kStForce] += 1;
    } else {
        enqueue_build(s);
        enqueue_update(s);
    }
    s->step_valid = true;
    s->out_stale = s->d.out_pos == nullptr;
    return 0;
}

// One half of the step ("Grid construction" or "SPH update" bucket) as a graph replay: at small N
// the step is launch-bound (hist + 3 passes + reorder = 6 launches for ~30 us of work).
int enqueue_half(sph_sim *s, bool build) {
    const bool want_graph = s->opt.use_graph != 2 && !s->profiling && s->keys_valid &&
                            (build ? s->counted == s->fuse_count : true);
    if (!want_graph) {
        if (build) enqueue_build(s); else enqueue_update(s);
        return 0;
    }
    cudaGraphExec_t &slot = build ? s->graph_build : s->graph_update;
    const int64_t l0 = s->launches;
    int64_t sl0[SPH_STAGE_COUNT];
    memcpy(sl0, s->stage_launches, sizeof(sl0));
    if (!slot) {
        cudaGraph_t g = nullptr;
        CU(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
        if (build) enqueue_build(s); else enqueue_update(s);
        CU(cudaStreamEndCapture(s->stream, &g));
        CU(cudaGraphInstantiate(&slot, g, 0));
        cudaGraphDestroy(g);
        s->launches = l0;
        memcpy(s->stage_launches, sl0, sizeof(sl0));
    }
    CU(cudaGraphLaunch(slot, s->stream));
    s->counted = build ? false : s->fuse_count;   // (what enqueue_build / enqueue_update record when launched plainly)
    if (build) {
        s->launches += 1 + hist_launches(s) + sort_launches(s);
        s->stage_launches[kStHist] += hist_launches(s); s->stage_launches[kStSort] += sort_launches(s); s->stage_launches[kStReorder] += 1;
    } else {
        s->launches += 2;
        s->stage_launches[kStDensity] += 1; s->stage_launches[kStForce] += 1;
    }
    return 0;
}

int sync_stream(sph_sim *s) {
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    if (s->profiling) profile_collect(s);
    return 0;
}

void drop_graph(sph_sim *s) {
    if (s->graph) cudaGraphExecDestroy(s->graph);
    if (s->graph_alt) cudaGraphExecDestroy(s->graph_alt);
    if (s->graph_build) cudaGraphExecDestroy(s->graph_build);
    if (s->graph_update) cudaGraphExecDestroy(s->graph_update);
    if (s->graph_noout) cudaGraphExecDestroy(s->graph_noout);
    s->graph = s->graph_alt = s->graph_build = s->graph_update = s->graph_noout = nullptr;
}

float bisect_sqrt_threshold(float target, bool smallest_ge) {
    // floats are ordered like their bit patterns for positive values
    uint32_t lo = 0, hi = 0x7f7fffffu;
    if (smallest_ge) {  // smallest t with sqrtf(t) >= target
        while (lo < hi) {
            uint32_t mid = lo + (hi - lo) / 2;
            float t;
            memcpy(&t, &mid, 4);
            if (sqrtf(t) >= target) hi = mid; else lo = mid + 1;
        }
    } else {            // largest t with sqrtf(t) <= target
        while (lo < hi) {
            uint32_t mid = lo + (hi - lo + 1) / 2;
            float t;
            memcpy(&t, &mid, 4);
            if (sqrtf(t) <= target) lo = mid; else hi = mid - 1;
        }
    }
    float r;
    memcpy(&r, &lo, 4);
    return r;
}

}  // namespace

// Internal: error reporting for the other translation units of the library (sph_cluster.cu).
extern "C" int sph_internal_fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

namespace {

int free_device(sph_sim *s) {
    DeviceState &d = s->d;
    cudaFree(d.cur_pos); cudaFree(d.cur_vel); cudaFree(d.srt_pos); cudaFree(d.srt_vel);
    cudaFree(d.key); cudaFree(d.pairs[0]); cudaFree(d.pairs[1]); cudaFree(d.cell_start);
    cudaFree(d.pa); cudaFree(d.rho); cudaFree(d.force); cudaFree(s->out_buf[0]); cudaFree(s->out_buf[1]);
    s->out_buf[0] = s->out_buf[1] = nullptr;
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    if (s->ev_step) cudaEventDestroy(s->ev_step);
    if (s->ev_copy) cudaEventDestroy(s->ev_copy);
    s->copy_stream = nullptr;
    s->ev_step = s->ev_copy = nullptr;
    cudaFree(d.sort_scratch); cudaFree(d.cell_count); cudaFree(d.stats); cudaFree(d.counts); cudaFree(d.masks.words); cudaFree(d.masks.base); cudaFree(d.masks.cursor); cudaFree(d.pair_xy); cudaFree(d.pair_z);
    cudaFree(s->p.dbg);
    s->p.dbg = nullptr;
    memset(&d, 0, sizeof(d));
    if (s->host_pos) cudaFreeHost(s->host_pos);
    s->host_pos = nullptr;
    return 0;
}

// Upload a full state given in ORIGINAL particle order.
int upload_state(sph_sim *s, const float *pos, const float *vel) {
    const int n = s->p.n;
    std::vector<float4> hp((size_t)n), hv((size_t)n);
    const float box = s->settings.boxDim;
    for (int i = 0; i < n; ++i) {
        const float x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
        if (!(x >= 0.f && x < box && y >= 0.f && y < box && z >= 0.f && z < box))
            return fail(SPH_E_OUT_OF_BOX, "particle %d at (%g, %g, %g) is outside [0, %g)^3", i, x, y,
                        z, box);
        uint32_t id = (uint32_t)i;
        float w;
        memcpy(&w, &id, 4);
        hp[i] = make_float4(x, y, z, w);
        hv[i] = vel ? make_float4(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2], 0.f)
                    : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    CU(cudaMemcpyAsync(s->d.cur_pos, hp.data(), sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, s->stream));
    CU(cudaMemcpyAsync(s->d.cur_vel, hv.data(), sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    s->keys_valid = false;
    s->step_valid = false;
    s->spec_inflight = false;
    s->out_stale = true;
    return 0;
}

// ids[slot] for the current storage order (from pos.w of `src`)
int download_ids(sph_sim *s, const float4 *src, std::vector<float4> &buf) {
    buf.resize((size_t)s->p.n);
    CU(cudaMemcpyAsync(buf.data(), src, sizeof(float4) * (size_t)s->p.n, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return 0;
}
inline uint32_t id_of(const float4 &v) {
    uint32_t id;
    memcpy(&id, &v.w, 4);
    return id;
}

#define REQUIRE_SETUP(s)                                                  \
    do {                                                                  \
        if (!(s)) return fail(SPH_E_INVALID, "null simulator handle");    \
        if (!(s)->is_setup) return fail(SPH_E_STATE, "sph_setup() has not been called"); \
        CU(cudaSetDevice((s)->opt.device));                               \
    } while (0)

}  // namespace

extern "C" {

int sph_abi_version(void) { return SPH_B200_ABI_VERSION; }
const char *sph_last_error(void) { return g_err; }
const char *sph_stage_name(int stage) {
    return (stage >= 0 && stage < SPH_STAGE_COUNT) ? kStageNames[stage] : "?";
}

int sph_create(const SphSettings *settings, sph_sim **out) {
    return sph_create_ex(settings, nullptr, out);
}

// ref: simulator.cu:370-375 -- the constructor only records the settings.
int sph_create_ex(const SphSettings *st, const SphOptions *options, sph_sim **out) {
    if (!st || !out) return fail(SPH_E_INVALID, "null argument");
    *out = nullptr;
    const int nc = (int)st->numCellsPerDim;
    if (st->numParticles < 0) return fail(SPH_E_INVALID, "numParticles < 0");
    if (!(st->h > 0.f) || !(st->boxDim > 0.f) || !(st->timestep > 0.f))
        return fail(SPH_E_INVALID, "h, boxDim and timestep must be positive");
    if (nc < 1 || nc > 1024 || (float)nc != st->numCellsPerDim)
        return fail(SPH_E_INVALID, "numCellsPerDim must be an integer in [1, 1024], got %g",
                    (double)st->numCellsPerDim);
    sph_sim *s = new (std::nothrow) sph_sim();
    if (!s) return fail(SPH_E_NOMEM, "out of host memory");
    s->settings = *st;
    memset(&s->opt, 0, sizeof(s->opt));
    if (options) s->opt = *options;
    if (s->opt.key_mode != SPH_KEY_FLAT && s->opt.key_mode != SPH_KEY_MORTON) {
        const int bad = s->opt.key_mode;
        delete s;
        return fail(SPH_E_INVALID, "unknown key_mode %d", bad);
    }
    if (s->opt.density_sum < 0 || s->opt.density_sum > 2) {
        const int bad = s->opt.density_sum;
        delete s;
        return fail(SPH_E_INVALID, "unknown density_sum %d", bad);
    }
    memset(&s->d, 0, sizeof(s->d));
    Params &p = s->p;
    p.n = p.n_owned = st->numParticles;
    p.nc = nc;
    p.h = st->h;
    p.h2 = st->h * st->h;
    p.vk = st->v_kernel_coeff;
    p.dk = st->d_kernel_coeff;
    p.box = st->boxDim;
    p.hi = st->boxDim - st->h;
    p.dt = st->timestep;
    p.key_mode = s->opt.key_mode;
    p.slab = 0; p.slot0 = 0; p.zoff = 0; p.ncz = nc; p.zlo = 0; p.zhi = nc; p.nz = nc;
    p.hi_z = p.hi; p.dead_key = 0xffffffffu;
    p.slot_begin = 0; p.slot_end = p.n; p.dbg = nullptr;
    if (s->opt.z_cell_hi > s->opt.z_cell_lo) {   // slab mode
        const int nz = s->opt.nz_cells > 0 ? s->opt.nz_cells : nc;
        if (p.key_mode != SPH_KEY_FLAT || s->opt.z_cell_lo < 0 || s->opt.z_cell_hi > nz) {
            delete s;
            return fail(SPH_E_INVALID, "slab mode needs flat keys and 0 <= z_cell_lo < z_cell_hi <= nz_cells");
        }
        p.slab = 1;
        p.zlo = s->opt.z_cell_lo; p.zhi = s->opt.z_cell_hi; p.nz = nz;
        p.zoff = p.zlo - 1;
        p.ncz = p.zhi - p.zlo + 2;
        p.hi_z = (float)nz * st->h - st->h;
        if ((double)nc * nc * (p.ncz + kSlabSpareLayers) >= (double)(1u << 30)) {
            delete s;
            return fail(SPH_E_INVALID, "slab too thick: nc^2 * (layers + 2) must stay below 2^30");
        }
    }
    if (p.key_mode == SPH_KEY_FLAT) {
        p.table_size = (uint32_t)nc * nc * (uint32_t)p.ncz;
        if (p.slab) p.dead_key = p.table_size - 1u;
    } else {
        int bits = 0;
        while ((1 << bits) < nc) ++bits;
        p.table_size = 1u << (3 * bits);
    }
    s->passes = sort_passes_for(p.table_size);
    {   // the step's sort: counting sort by cell on a single GPU unless the radix passes are asked for
        const char *e = getenv("SPH_SORT");
        const bool radix = s->opt.sort_algo == SPH_SORT_RADIX || (s->opt.sort_algo == 0 && e && !strcmp(e, "radix"));
        s->cell_sort = !radix;
        const char *f = getenv("SPH_FUSE_COUNT");
        s->fuse_count = s->cell_sort && !p.slab && p.key_mode == SPH_KEY_FLAT && !(f && !strcmp(f, "0"));
    }
    s->th.r2_eps = bisect_sqrt_threshold(kEps, true);
    s->th.r2_h = bisect_sqrt_threshold(st->h, false);
    s->capacity = s->opt.capacity > p.n ? s->opt.capacity : p.n;
    *out = s;
    return 0;
}

void sph_destroy(sph_sim *s) {
    if (!s) return;
    if (s->is_setup) {
        cudaSetDevice(s->opt.device);
        if (s->stream) cudaStreamSynchronize(s->stream);
        drop_graph(s);
        for (auto &ep : s->events) {
            cudaEventDestroy(ep.a);
            cudaEventDestroy(ep.b);
        }
        free_device(s);
        if (s->stream && s->own_stream) cudaStreamDestroy(s->stream);
    }
    delete s;
}

// Device half of sph_setup(); on failure the caller releases whatever was allocated.
static int setup_device(sph_sim *s, int n, const std::vector<float> &pos) {
    const SphSettings &st = s->settings;
    (void)st;
    // -- device ---------------------------------------------------------------
    CU(cudaSetDevice(s->opt.device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, s->opt.device));
    if (prop.major < 10)
        return fail(SPH_E_INVALID, "device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                    s->opt.device, prop.major, prop.minor);
    s->sm_count = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    const size_t cap = (size_t)(s->capacity > 0 ? s->capacity : 1);
    DeviceState &d = s->d;
    size_t scap = cap;   // sorted-slot capacity: owned + ghost room on both sides in slab mode
    if (s->p.slab) {
        int g = s->opt.ghost_capacity > 0 ? s->opt.ghost_capacity : (int)(cap / 4) + 1024;
        g = (g + 1) & ~1;   // slot0 must be even (pair-interleaved copy)
        s->ghost_cap = g;
        s->p.slot0 = g;
        scap = cap + 2 * (size_t)g;
        // (emigrants go straight into the cluster's migration messages: sph_cluster.cu points
        // d.emig_* at them)
    }
    CU(cudaMalloc(&d.cur_pos, cap * sizeof(float4)));
    CU(cudaMalloc(&d.cur_vel, cap * sizeof(float4)));
    CU(cudaMalloc(&d.srt_pos, scap * sizeof(float4)));
    CU(cudaMalloc(&d.srt_vel, scap * sizeof(float4)));
    CU(cudaMalloc(&d.key, cap * sizeof(uint32_t)));
    CU(cudaMalloc(&d.pairs[0], cap * sizeof(uint64_t)));
    CU(cudaMalloc(&d.pairs[1], cap * sizeof(uint64_t)));
    // slab mode: the owned layer range may grow by a few layers when the cluster rebalances
    s->table_capacity = s->p.table_size + (s->p.slab ? (uint32_t)s->p.nc * s->p.nc * (uint32_t)kSlabSpareLayers : 0u);
    CU(cudaMalloc(&d.cell_start, ((size_t)s->table_capacity + 1) * sizeof(uint32_t)));
    CU(cudaMalloc(&d.pa, scap * sizeof(float2)));
    CU(cudaMalloc(&d.rho, scap * sizeof(float)));
    if (!s->p.slab) {
        CU(cudaMalloc(&s->out_buf[0], cap * 3 * sizeof(float)));
        d.out_pos = s->out_buf[0];
        if (s->opt.pipeline_readback) {
            CU(cudaMalloc(&s->out_buf[1], cap * 3 * sizeof(float)));
            CU(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&s->ev_step, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&s->ev_copy, cudaEventDisableTiming));
        }
    }
    {
        size_t words = sort_scratch_words((int)cap);
        if (s->cell_sort) {
            words = std::max(words, cell_sort_scratch_words(s->table_capacity + 1u));
            const size_t entries = (size_t)s->table_capacity + 1 + 3;   // the scan moves whole uint4
            CU(cudaMalloc(&d.cell_count, entries * sizeof(uint32_t)));
            CU(cudaMemsetAsync(d.cell_count, 0, entries * sizeof(uint32_t), s->stream));
        }
        CU(cudaMalloc(&d.sort_scratch, words * sizeof(uint32_t)));
    }
    CU(cudaMalloc(&d.stats, 2 * sizeof(double)));
    CU(cudaMalloc(&s->p.dbg, sizeof(uint32_t)));
    CU(cudaMemsetAsync(s->p.dbg, 0, sizeof(uint32_t), s->stream));
    if (s->opt.record_force) CU(cudaMalloc(&d.force, scap * sizeof(float4)));
    if (s->p.key_mode == SPH_KEY_FLAT) {
        // + 2 records: the staged tiles copy even-aligned pair ranges
        CU(cudaMalloc(&d.pair_xy, ((scap + 1) / 2 + 2) * sizeof(float4)));
        CU(cudaMalloc(&d.pair_z, ((scap + 1) / 2 + 2) * sizeof(float2)));
        CU(cudaMemsetAsync(d.pair_xy, 0, ((scap + 1) / 2 + 2) * sizeof(float4), s->stream));
        CU(cudaMemsetAsync(d.pair_z, 0, ((scap + 1) / 2 + 2) * sizeof(float2), s->stream));
        d.stage_tiles = s->opt.stage_tiles ? 1 : 0;
        d.density_exact = s->opt.density_sum == 2 ? 0 : (s->opt.density_sum == 1 ? 1 : kDefaultDensityExact);
    }
    if (s->p.key_mode == SPH_KEY_FLAT) {
        const size_t warps = (cap + kBlock - 1) / kBlock * (kBlock / 32);
        CU(cudaMalloc(&d.masks.base, warps * sizeof(uint32_t)));
        CU(cudaMemsetAsync(d.masks.base, 0xff, warps * sizeof(uint32_t), s->stream));
        CU(cudaMalloc(&d.masks.cursor, kMaskPools * 32 * sizeof(uint32_t)));
        if (!s->opt.no_mask_handoff) {
            d.masks.rows = (uint32_t)((warps * kMaskRowsPerWarp + kMaskPools - 1) / kMaskPools + 64);
            CU(cudaMalloc(&d.masks.words, (size_t)d.masks.rows * kMaskPools * 32 * sizeof(uint32_t)));
        }
    }
    CU(cudaMemsetAsync(d.cell_start, 0, ((size_t)s->p.table_size + 1) * sizeof(uint32_t), s->stream));
    CU(cudaMemsetAsync(d.rho, 0, scap * sizeof(float), s->stream));
    if (!s->p.slab) {
        CU(cudaMallocHost(&s->host_pos, cap * 3 * sizeof(float)));
        memset(s->host_pos, 0, cap * 3 * sizeof(float));
    } else {
        s->p.n = 0;
    }
    // every clear above ran on the simulator's own stream (it is non-blocking: the legacy default
    // stream would not be ordered against it)
    CU(cudaStreamSynchronize(s->stream));
    if (n > 0) {
        int rc = upload_state(s, pos.data(), nullptr);
        if (rc) return rc;
    }
    return 0;
}


// ref: simulator.cu:411-460
int sph_setup(sph_sim *s) {
    if (!s) return fail(SPH_E_INVALID, "null simulator handle");
    if (s->is_setup) return fail(SPH_E_STATE, "sph_setup() called twice");
    const SphSettings &st = s->settings;
    const int n = s->p.slab ? 0 : s->p.n;   // slab mode: particles come from sph_slab_load()

    // -- initial positions, on the host exactly as the reference computes them --
    std::vector<float> pos((size_t)3 * (n > 0 ? n : 1));
    if (s->p.slab) {
        // nothing: the caller decides which particles this slab owns
    } else if (st.randomInit) {
        // unseeded glibc rand(): three draws per particle in x, y, z order (ref: 430-437)
        for (int i = 0; i < n; ++i) {
            const float x = rand() / (float)RAND_MAX * (st.boxDim - 2.f) + 1.f;
            const float y = rand() / (float)RAND_MAX * (st.boxDim - 2.f) + 1.f;
            const float z = rand() / (float)RAND_MAX * (st.boxDim - 2.f) + 1.f;
            pos[3 * (size_t)i] = x;
            pos[3 * (size_t)i + 1] = y;
            pos[3 * (size_t)i + 2] = z;
        }
    } else {
        // 0.9h lattice from (h,h,h), x outer / z inner (ref: 438-453)
        const float spacing = 0.9f * st.h;
        const int nx = (int)(floorf((st.boxDim - 2 * st.h) / spacing) + 1);
        if ((long long)n > (long long)nx * nx * nx)
            return fail(SPH_E_INVALID,
                        "grid init: the %d^3 lattice of a boxDim=%g box holds %lld particles, %d "
                        "requested (the reference leaves the rest uninitialised); scale boxDim / "
                        "numCellsPerDim",
                        nx, (double)st.boxDim, (long long)nx * nx * nx, n);
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

    int rc = setup_device(s, n, pos);
    if (rc) {   // leave nothing behind: a retried sph_setup() starts from scratch
        free_device(s);
        if (s->stream && s->own_stream) cudaStreamDestroy(s->stream);
        s->stream = nullptr;
        return rc;
    }
    s->is_setup = true;
    return 0;
}

#define NOT_IN_SLAB_MODE(s) \
    do { if ((s)->p.slab) return fail(SPH_E_STATE, "not available in slab mode (use the sph_slab_* calls)"); } while (0)

// sph_step() with SphOptions.pipeline_readback: the copy of step k's positions (copy stream)
// overlaps the computation of step k+1 (compute stream), which is started before returning.
static int step_pipelined(sph_sim *s) {
    const size_t bytes = sizeof(float) * 3 * (size_t)s->p.n;
    int rc = 0;
    if (!s->spec_inflight) {   // first call, or the state was replaced: compute the step asked for
        s->d.out_pos = s->out_buf[s->out_parity];
        rc = enqueue_step(s);
        if (rc) return rc;
    }
    CU(cudaEventRecord(s->ev_step, s->stream));
    CU(cudaStreamWaitEvent(s->copy_stream, s->ev_step, 0));
    CU(cudaMemcpyAsync(s->host_pos, s->out_buf[s->out_parity], bytes, cudaMemcpyDeviceToHost, s->copy_stream));
    CU(cudaEventRecord(s->ev_copy, s->copy_stream));
    // speculative next step into the other output buffer
    s->out_parity ^= 1;
    s->d.out_pos = s->out_buf[s->out_parity];
    rc = enqueue_step(s);
    if (rc) return rc;
    s->spec_inflight = true;
    CU(cudaEventSynchronize(s->ev_copy));
    CU(cudaGetLastError());
    return 0;
}

// A pipelined sph_step() left one more step enqueued than it handed out (its positions are in
// out_buf[out_parity]).  The other stepping calls count that step as their first one, so that
// mixing them with pipelined sph_step() advances the state by exactly the steps asked for.
// Returns 1 if such a step was consumed.
static int take_speculative_step(sph_sim *s) {
    if (!s->spec_inflight) return 0;
    s->spec_inflight = false;
    return 1;
}

int sph_step(sph_sim *s) {
    REQUIRE_SETUP(s);
    NOT_IN_SLAB_MODE(s);
    if (s->p.n == 0) return 0;
    if (s->opt.pipeline_readback) return step_pipelined(s);
    s->d.out_pos = s->out_buf[0];
    s->out_parity = 0;
    int rc = enqueue_step(s);
    if (rc) return rc;
    // ref: simulator.cu:478-480 -- blocking copy of every position to the host
    CU(cudaMemcpyAsync(s->host_pos, s->d.out_pos, sizeof(float) * 3 * (size_t)s->p.n,
                       cudaMemcpyDeviceToHost, s->stream));
    return sync_stream(s);
}

// ref: simulator.cu:499-546 -- same wall-clock buckets, launch + sync inside each
int sph_step_timed(sph_sim *s, SphTimes *times) {
    REQUIRE_SETUP(s);
    NOT_IN_SLAB_MODE(s);
    if (!times) return fail(SPH_E_INVALID, "null times");
    using clk = std::chrono::steady_clock;
    auto secs = [](clk::time_point a) {
        return std::chrono::duration_cast<std::chrono::duration<double>>(clk::now() - a).count();
    };
    if (s->p.n > 0) {
        int rc = 0;
        if (take_speculative_step(s)) {
            // the step is already on the stream (enqueued by a pipelined sph_step): its device time
            // cannot be split into the two buckets any more, it is charged to "SPH update"
            auto t1 = clk::now();
            rc = sync_stream(s);
            if (rc) return rc;
            times->sphUpdate += secs(t1);
        } else {
            s->d.out_pos = s->out_buf[0];
            s->out_parity = 0;
            auto t0 = clk::now();
            rc = enqueue_half(s, true);
            if (rc == 0) rc = sync_stream(s);
            if (rc) return rc;
            times->buildGrid += secs(t0);

            auto t1 = clk::now();
            rc = enqueue_half(s, false);
            if (rc == 0) rc = sync_stream(s);
            if (rc) return rc;
            times->sphUpdate += secs(t1);
            s->step_valid = true;
            s->out_stale = false;
        }

        auto t2 = clk::now();
        CU(cudaMemcpyAsync(s->host_pos, s->out_buf[s->out_parity], sizeof(float) * 3 * (size_t)s->p.n,
                           cudaMemcpyDeviceToHost, s->stream));
        rc = sync_stream(s);
        if (rc) return rc;
        times->memcpy += secs(t2);
    }
    times->iters += 1;
    return 0;
}

int sph_advance(sph_sim *s, int steps) {
    REQUIRE_SETUP(s);
    NOT_IN_SLAB_MODE(s);
    if (steps < 0) return fail(SPH_E_INVALID, "steps < 0");
    if (s->p.n == 0) return 0;
    if (steps > 0) steps -= take_speculative_step(s);
    s->d.out_pos = nullptr;   // positions in id order are produced on demand (sph_readback)
    for (int k = 0; k < steps; ++k) {
        int rc = enqueue_step(s);
        if (rc) return rc;
    }
    return sync_stream(s);
}

int sph_advance_timed(sph_sim *s, int steps, float *ms) {
    REQUIRE_SETUP(s);
    NOT_IN_SLAB_MODE(s);
    if (steps < 0 || !ms) return fail(SPH_E_INVALID, "bad argument");
    if (take_speculative_step(s)) {   // timing starts from a quiet stream, with the pending step counted
        int rc = sync_stream(s);
        if (rc) return rc;
        if (steps > 0) --steps;
    }
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a));
    CU(cudaEventCreate(&b));
    CU(cudaEventRecord(a, s->stream));
    int rc = 0;
    s->d.out_pos = nullptr;
    for (int k = 0; k < steps && rc == 0 && s->p.n > 0; ++k) rc = enqueue_step(s);
    cudaEventRecord(b, s->stream);
    if (rc == 0) rc = sync_stream(s);
    *ms = 0.f;
    if (rc == 0) cudaEventElapsedTime(ms, a, b);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    return rc;
}

int sph_push(sph_sim *s, int x, int y) {
    REQUIRE_SETUP(s);
    NOT_IN_SLAB_MODE(s);
    if (s->p.n == 0) return 0;
    if (!s->step_valid) return fail(SPH_E_STATE, "sph_push() needs the cell table of a step that just ran");
    stage_begin(s, kStPush);
    launch_push(s->p, s->d, x, y, s->stream);
    stage_end(s);
    return sync_stream(s);
}

const float *sph_positions_host(sph_sim *s) { return s ? s->host_pos : nullptr; }

int sph_readback(sph_sim *s) {
    REQUIRE_SETUP(s);
    NOT_IN_SLAB_MODE(s);
    if (s->p.n == 0) return 0;
    take_speculative_step(s);   // the host buffer now shows the internal (latest) step
    float *src = s->out_buf[s->out_parity];
    if (s->out_stale) {         // the last steps ran without the id-ordered copy: make it now
        src = s->out_buf[0];
        s->out_parity = 0;
        stage_begin(s, kStOther);
        launch_unpermute(s->p, s->d, src, s->stream);
        stage_end(s);
        s->out_stale = false;
    }
    CU(cudaMemcpyAsync(s->host_pos, src, sizeof(float) * 3 * (size_t)s->p.n, cudaMemcpyDeviceToHost,
                       s->stream));
    return sync_stream(s);
}

int sph_set_state(sph_sim *s, const float *pos, const float *vel) {
    REQUIRE_SETUP(s);
    NOT_IN_SLAB_MODE(s);
    if (!pos) return fail(SPH_E_INVALID, "null positions");
    if (s->p.n == 0) return 0;
    return upload_state(s, pos, vel);
}

int sph_get_state(sph_sim *s, float *pos, float *vel) {
    REQUIRE_SETUP(s);
    NOT_IN_SLAB_MODE(s);
    const int n = s->p.n;
    if (n == 0) return 0;
    std::vector<float4> hp, hv;
    int rc = download_ids(s, s->d.cur_pos, hp);
    if (rc) return rc;
    if (vel) {
        rc = download_ids(s, s->d.cur_vel, hv);
        if (rc) return rc;
    }
    for (int i = 0; i < n; ++i) {
        const uint32_t id = id_of(hp[i]);
        if (id >= (uint32_t)n) return fail(SPH_E_STATE, "corrupt particle id %u at slot %d", id, i);
        if (pos) { pos[3 * id] = hp[i].x; pos[3 * id + 1] = hp[i].y; pos[3 * id + 2] = hp[i].z; }
        if (vel) { vel[3 * id] = hv[i].x; vel[3 * id + 1] = hv[i].y; vel[3 * id + 2] = hv[i].z; }
    }
    return 0;
}

int sph_get_keys(sph_sim *s, int key_mode, uint32_t *keys) {
    REQUIRE_SETUP(s);
    NOT_IN_SLAB_MODE(s);
    if (!keys) return fail(SPH_E_INVALID, "null keys");
    if (key_mode != SPH_KEY_FLAT && key_mode != SPH_KEY_MORTON)
        return fail(SPH_E_INVALID, "unknown key_mode %d", key_mode);
    const int n = s->p.n;
    if (n == 0) return 0;
    // hash with the requested key form into a temporary, without touching the step's keys
    uint32_t *tmp = nullptr;
    CU(cudaMalloc(&tmp, sizeof(uint32_t) * (size_t)n));
    Params p = s->p;
    p.key_mode = key_mode;
    DeviceState d = s->d;
    d.key = tmp;
    stage_begin(s, kStHash);
    launch_hash(p, d, s->stream);
    stage_end(s);
    std::vector<uint32_t> hk((size_t)n);
    cudaError_t e = cudaMemcpyAsync(hk.data(), tmp, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost, s->stream);
    int rc = e == cudaSuccess ? sync_stream(s) : fail((int)e, "copy failed: %s", cudaGetErrorString(e));
    cudaFree(tmp);
    if (rc) return rc;
    std::vector<float4> hp;
    rc = download_ids(s, s->d.cur_pos, hp);
    if (rc) return rc;
    for (int i = 0; i < n; ++i) keys[id_of(hp[i])] = hk[i];
    return 0;
}

int sph_get_sorted_index(sph_sim *s, uint32_t *ids, uint32_t *sorted_keys) {
    REQUIRE_SETUP(s);
    NOT_IN_SLAB_MODE(s);
    const int n = s->p.n;
    if (n == 0) return 0;
    if (!s->step_valid) return fail(SPH_E_STATE, "no step has run since the state was set");
    if (ids) {
        std::vector<float4> hp;
        int rc = download_ids(s, s->d.srt_pos, hp);
        if (rc) return rc;
        for (int i = 0; i < n; ++i) ids[i] = id_of(hp[i]);
    }
    if (sorted_keys) {
        std::vector<uint64_t> pr((size_t)n);
        CU(cudaMemcpyAsync(pr.data(), s->d.pairs[s->sorted_buf], sizeof(uint64_t) * (size_t)n,
                           cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
        for (int i = 0; i < n; ++i) sorted_keys[i] = (uint32_t)(pr[i] >> 32);
    }
    return 0;
}

int sph_get_cell_start(sph_sim *s, uint32_t *start, uint32_t *table_size) {
    REQUIRE_SETUP(s);
    if (table_size) *table_size = s->p.table_size;
    if (!start) return 0;
    if (!s->step_valid) return fail(SPH_E_STATE, "no step has run since the state was set");
    CU(cudaMemcpyAsync(start, s->d.cell_start, sizeof(uint32_t) * ((size_t)s->p.table_size + 1),
                       cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return 0;
}

int sph_get_neighbor_counts(sph_sim *s, int32_t *K, int32_t *C) {
    REQUIRE_SETUP(s);
    NOT_IN_SLAB_MODE(s);
    const int n = s->p.n;
    if (n == 0) return 0;
    if (!s->d.counts) CU(cudaMalloc(&s->d.counts, sizeof(int32_t) * 2 * (size_t)s->capacity));
    enqueue_build(s);  // grid of the CURRENT positions; the state itself is untouched
    stage_begin(s, kStDensity);
    launch_density(s->p, s->th, s->d, true, s->stream);
    stage_end(s);
    s->step_valid = false;  // rho/pa/force no longer line up with the sorted slots
    std::vector<int32_t> h((size_t)2 * n);
    CU(cudaMemcpyAsync(h.data(), s->d.counts, sizeof(int32_t) * 2 * (size_t)n, cudaMemcpyDeviceToHost, s->stream));
    int rc = sync_stream(s);
    if (rc) return rc;
    std::vector<float4> hp;
    rc = download_ids(s, s->d.srt_pos, hp);
    if (rc) return rc;
    for (int i = 0; i < n; ++i) {
        const uint32_t id = id_of(hp[i]);
        if (K) K[id] = h[i];
        if (C) C[id] = h[(size_t)n + i];
    }
    return 0;
}

int sph_get_density_pressure_force(sph_sim *s, float *rho, float *prs, float *force) {
    REQUIRE_SETUP(s);
    NOT_IN_SLAB_MODE(s);
    const int n = s->p.n;
    if (n == 0) return 0;
    if (!s->step_valid) return fail(SPH_E_STATE, "no step has run since the state was set");
    if (force && !s->d.force)
        return fail(SPH_E_STATE, "forces are only kept when SphOptions.record_force is set");
    std::vector<float4> hp;
    int rc = download_ids(s, s->d.srt_pos, hp);
    if (rc) return rc;
    std::vector<float> hr((size_t)n);
    std::vector<float2> ha((size_t)n);
    std::vector<float4> hf;
    CU(cudaMemcpyAsync(hr.data(), s->d.rho, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaMemcpyAsync(ha.data(), s->d.pa, sizeof(float2) * (size_t)n, cudaMemcpyDeviceToHost, s->stream));
    if (force) {
        hf.resize((size_t)n);
        CU(cudaMemcpyAsync(hf.data(), s->d.force, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost, s->stream));
    }
    CU(cudaStreamSynchronize(s->stream));
    for (int i = 0; i < n; ++i) {
        const uint32_t id = id_of(hp[i]);
        if (rho) rho[id] = hr[i];
        if (prs) prs[id] = ha[i].x;
        if (force) { force[3 * id] = hf[i].x; force[3 * id + 1] = hf[i].y; force[3 * id + 2] = hf[i].z; }
    }
    return 0;
}

int sph_get_stats(sph_sim *s, double *ke, double *mean_rho) {
    REQUIRE_SETUP(s);
    double h[2] = {0, 0};
    if (s->p.n > 0) {
        stage_begin(s, kStOther);
        launch_stats(s->p, s->d, s->stream);
        stage_end(s);
        CU(cudaMemcpyAsync(h, s->d.stats, sizeof(h), cudaMemcpyDeviceToHost, s->stream));
        int rc = sync_stream(s);
        if (rc) return rc;
    }
    if (ke) *ke = h[0];
    if (mean_rho) *mean_rho = s->p.n ? h[1] / s->p.n : 0.0;
    return 0;
}

// ---- slab decomposition ---------------------------------------------------------------
#define REQUIRE_SLAB(s)                                                               \
    do {                                                                              \
        REQUIRE_SETUP(s);                                                             \
        if (!(s)->p.slab) return fail(SPH_E_STATE, "simulator was not created in slab mode"); \
    } while (0)

int sph_slab_load(sph_sim *s, int n, const float *pos, const float *vel, const uint32_t *ids) {
    REQUIRE_SLAB(s);
    if (n < 0 || n > s->capacity) return fail(SPH_E_INVALID, "n = %d exceeds the capacity %d", n, s->capacity);
    if (n && (!pos || !ids)) return fail(SPH_E_INVALID, "null argument");
    std::vector<float4> hp((size_t)n), hv((size_t)n);
    for (int i = 0; i < n; ++i) {
        float w;
        memcpy(&w, &ids[i], 4);
        hp[i] = make_float4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], w);
        hv[i] = vel ? make_float4(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2], 0.f) : make_float4(0, 0, 0, 0);
    }
    if (n) {
        CU(cudaMemcpyAsync(s->d.cur_pos, hp.data(), sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, s->stream));
        CU(cudaMemcpyAsync(s->d.cur_vel, hv.data(), sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, s->stream));
    }
    CU(cudaStreamSynchronize(s->stream));
    s->p.n = n; s->step_valid = false;
    return 0;
}

// Internal (declared in sph_internal.cuh, not part of the public header): what the slab cluster
// driver (sph_cluster.cu) needs from a simulator created in slab mode.
int sph_internal_core(sph_sim *s, sph::SlabCore *out) {
    REQUIRE_SLAB(s);
    out->p = &s->p;
    out->d = &s->d;
    out->th = &s->th;
    out->stream = s->stream;
    out->capacity = s->capacity;
    out->ghost_cap = s->ghost_cap;
    out->passes = &s->passes;
    out->sm_count = s->sm_count;
    out->sorted_buf = &s->sorted_buf;
    out->device = s->opt.device;
    out->table_capacity = s->table_capacity;
    out->cell_sort = s->cell_sort;
    return 0;
}

int sph_debug_flags(sph_sim *s, uint32_t *flags, int *checked_build) {
    REQUIRE_SETUP(s);
    if (!flags) return fail(SPH_E_INVALID, "null argument");
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaMemcpy(flags, s->p.dbg, sizeof(uint32_t), cudaMemcpyDeviceToHost));
#ifdef SPH_BOUNDS_CHECK
    if (checked_build) *checked_build = 1;
#else
    if (checked_build) *checked_build = 0;
#endif
    return 0;
}

int sph_profile_enable(sph_sim *s, int on) {
    REQUIRE_SETUP(s);
    CU(cudaStreamSynchronize(s->stream));
    s->profiling = on != 0;
    s->events_used = 0;
    return 0;
}

int sph_profile_read(sph_sim *s, double ms[SPH_STAGE_COUNT], int64_t launches[SPH_STAGE_COUNT], int reset) {
    REQUIRE_SETUP(s);
    CU(cudaStreamSynchronize(s->stream));
    if (s->profiling) profile_collect(s);
    for (int i = 0; i < SPH_STAGE_COUNT; ++i) {
        if (ms) ms[i] = s->stage_ms[i];
        if (launches) launches[i] = s->stage_launches[i];
        if (reset) {
            s->stage_ms[i] = 0;
            s->stage_launches[i] = 0;
        }
    }
    return 0;
}

int64_t sph_launch_count(sph_sim *s) { return s ? s->launches : 0; }
int sph_num_particles(sph_sim *s) { return s ? s->p.n : 0; }
int sph_sort_info(sph_sim *s, int32_t *algo, int32_t *kernels, int32_t *count_fused, int32_t *radix_passes) {
    if (!s) return fail(SPH_E_INVALID, "null simulator");
    if (algo) *algo = s->cell_sort ? SPH_SORT_COUNT : SPH_SORT_RADIX;
    if (kernels) *kernels = hist_launches(s) + sort_launches(s);
    if (count_fused) *count_fused = s->fuse_count ? 1 : 0;
    if (radix_passes) *radix_passes = s->passes;
    return 0;
}

}  // extern "C"
