// sph_internal.cuh -- what sph_cluster.cu needs from sph_api.cu beyond the public C ABI.
#pragma once

#include "../../include/sph_b200.h"
#include "sph_kernels.cuh"

namespace sph {

constexpr int kSlabSpareLayers = 8;   // cell-table room for a slab's layer range to grow (rebalancing)

// Handles into a slab-mode simulator: the cluster driver launches the step kernels itself, with
// counts that live in device memory (SlabDyn), on the simulator's own stream.
struct SlabCore {
    Params *p;
    DeviceState *d;
    Thresholds *th;
    cudaStream_t stream;
    int capacity, ghost_cap;
    int *passes;
    int sm_count;
    int *sorted_buf;
    int device;
    uint32_t table_capacity;
    bool cell_sort;   // counting sort by cell (d->cell_count allocated) instead of the radix passes
};

}  // namespace sph

extern "C" int sph_internal_core(sph_sim *s, sph::SlabCore *out);
