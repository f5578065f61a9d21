// simulator.cpp -- `class Simulator` of include/simulator.h as a thin forwarder to the
// C ABI of libsph_b200.so.  Host-side replacement for ref src/simulator.cu:370-546.
#include "simulator.h"

#include <chrono>
#include <cstdlib>
#include <cstring>

#include "sph_b200.h"

// The reference passes mouse clicks through two process globals defined in
// display.cpp:19-20 and declared extern in simulator.cu:16-17.  Weak definitions keep
// that hand-off working when a display.cpp is linked and let headless programs link
// without one.
__attribute__((weak)) bool mouseClicked = false;
__attribute__((weak)) int2 clickCoords = {0, 0};

static_assert(sizeof(Settings) == sizeof(SphSettings), "Settings layout drifted from the C ABI");
static_assert(sizeof(Times) == sizeof(SphTimes), "Times layout drifted from the C ABI");

namespace {

SphOptions options_from_env() {
    SphOptions o;
    std::memset(&o, 0, sizeof o);
    if (const char *k = std::getenv("SPH_KEY_MODE"))
        o.key_mode = (std::strcmp(k, "morton") == 0) ? SPH_KEY_MORTON : SPH_KEY_FLAT;
    if (const char *d = std::getenv("SPH_DEVICE")) o.device = std::atoi(d);
    // opt-in: simulate() overlaps the host copy of step k with the computation of step k+1
    // (same positions; mouse pushes then land one step later, see SphOptions.pipeline_readback)
    if (const char *r = std::getenv("SPH_PIPELINE_READBACK")) o.pipeline_readback = std::atoi(r) != 0;
    return o;
}

}  // namespace

Simulator::Simulator(Settings *settings) : settings(settings), impl(NULL) {}

Simulator::~Simulator() {
    sph_destroy(impl);
    impl = NULL;
    sph_cluster_destroy(cluster);
    cluster = NULL;
    free(clusterPositions);
    clusterPositions = NULL;
}

void Simulator::note(int rc, const char *what) {
    lastStatus = rc;
    if (rc != 0) fprintf(stderr, "sph: %s failed (%d): %s\n", what, rc, sph_last_error());
}

void Simulator::setup() {
    // Settings is read here, not in the constructor, exactly like the reference
    // (which dereferences the pointer in setup() and on every step).
    SphSettings s;
    std::memset(&s, 0, sizeof s);
    s.randomInit = settings->randomInit ? 1 : 0;
    s.numParticles = settings->numParticles;
    s.h = settings->h;
    s.v_kernel_coeff = settings->v_kernel_coeff;
    s.d_kernel_coeff = settings->d_kernel_coeff;
    s.boxDim = settings->boxDim;
    s.numCellsPerDim = settings->numCellsPerDim;
    s.timestep = settings->timestep;
    SphOptions o = options_from_env();
    const char *g = std::getenv("SPH_GPUS");
    const int gpus = g ? std::atoi(g) : 1;
    if (gpus > 1) {
        // one process, one z-slab per GPU; the reference's particle set, the reference's ids
        SphClusterOptions co;
        std::memset(&co, 0, sizeof co);
        co.world = co.local_count = gpus;
        // SPH_GPUS_SAME_DEVICE=1: every slab on one GPU (tests on a single-GPU machine)
        const char *same = std::getenv("SPH_GPUS_SAME_DEVICE");
        const bool one_device = same && std::atoi(same) != 0;
        for (int i = 0; i < gpus && i < SPH_MAX_LOCAL_SLABS; ++i) co.devices[i] = one_device ? o.device : o.device + i;
        if (const char *r = std::getenv("SPH_REBALANCE_EVERY")) co.rebalance_every = std::atoi(r);
        int rc = gpus <= SPH_MAX_LOCAL_SLABS ? sph_cluster_create(&s, &co, &cluster) : SPH_E_INVALID;
        note(rc, "sph_cluster_create");
        if (rc == 0) note(sph_cluster_setup(cluster), "sph_cluster_setup");
        clusterPositions = static_cast<float *>(std::calloc((size_t)3 * (s.numParticles > 0 ? s.numParticles : 1), sizeof(float)));
        return;
    }
    int rc = sph_create_ex(&s, &o, &impl);
    note(rc, "sph_create_ex");
    if (rc == 0) note(sph_setup(impl), "sph_setup");
}

const float3 *Simulator::getPosition() {
    if (cluster) return reinterpret_cast<const float3 *>(clusterPositions);
    return reinterpret_cast<const float3 *>(sph_positions_host(impl));
}

void Simulator::simulate() {
    if (cluster) {   // (the mouse push needs the single-GPU cell table: not offered across slabs)
        note(sph_cluster_advance(cluster, 1), "sph_cluster_advance");
        if (lastStatus == 0)
            note(sph_cluster_positions(cluster, clusterPositions, settings->numParticles), "sph_cluster_positions");
        mouseClicked = false;
        return;
    }
    if (!impl) return;
    note(sph_step(impl), "sph_step");
    if (mouseClicked) {  // ref: simulator.cu:482-489
        note(sph_push(impl, clickCoords.x, clickCoords.y), "sph_push");
        mouseClicked = false;
    }
}

void Simulator::simulateAndTime(Times *times) {
    if (cluster) {
        // the slab step is enqueued as a whole (sort ... migration, no host round trip), so its two
        // compute buckets cannot be told apart from the host: all of it is charged to "SPH update"
        using clk = std::chrono::steady_clock;
        const auto t0 = clk::now();
        note(sph_cluster_advance(cluster, 1), "sph_cluster_advance");
        const auto t1 = clk::now();
        if (lastStatus == 0)
            note(sph_cluster_positions(cluster, clusterPositions, settings->numParticles), "sph_cluster_positions");
        const auto t2 = clk::now();
        times->sphUpdate += std::chrono::duration<double>(t1 - t0).count();
        times->memcpy += std::chrono::duration<double>(t2 - t1).count();
        times->iters += 1;
        return;
    }
    if (!impl) return;
    note(sph_step_timed(impl, reinterpret_cast<SphTimes *>(times)), "sph_step_timed");
}

void Simulator::moveParticles(int2 mouse_pos) {
    if (!impl) return;
    note(sph_push(impl, mouse_pos.x, mouse_pos.y), "sph_push");
}
