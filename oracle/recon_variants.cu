/*
 * recon_variants.cu -- BENCHMARK COMPARATORS, TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * RECONSTRUCTIONS of the reference's two sorted neighbour-search variants, whose source is
 * NOT in /root/reference.  Everything known about them is one sentence of the
 * README (README.md:5): `index_sort` "uses an array of particle indices sorted by flattened grid
 * index", `z_index_sort` "... sorted by the Morton-encoded grid index".  The least-assumption
 * reconstruction is implemented here:
 *   - the reference's 56-byte AoS `Particle` array stays UNSORTED (only an index array is sorted);
 *   - key_i from the reference's own getGridCell + flattenGridCoord (or a Morton interleave of the
 *     same int3);
 *   - (key, index) pairs sorted with CUB DeviceRadixSort (a library sort was the 15-418/618 norm;
 *     that it was CUB is a guess);
 *   - cellStart/cellEnd by boundary detection on the sorted keys;
 *   - the reference's neighbour loops with `while (neighbor != NULL)` replaced by
 *     `for k in [start, end): neighbor = &particles[index[k]]`, calling the reference's OWN device
 *     functions densityKernel / pressureKernel / viscosityKernel, then its own
 *     kernelUpdatePositions.
 * Every number produced by this file must be labelled "reconstructed from README, not reference
 * source; parity unpinned".  The unmodified reference TU is #included from where it lies.
 */
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <stdio.h>
#include <string>

#include <cub/cub.cuh>

#define private public
#include REF_SIM_CU
#undef private

bool mouseClicked = false;
int2 clickCoords = {0, 0};

namespace {

__device__ __forceinline__ unsigned spread3(unsigned v) {
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__device__ __forceinline__ unsigned cellKey(int3 c, int morton) {
    if (morton) return spread3(c.x) | (spread3(c.y) << 1) | (spread3(c.z) << 2);
    return (unsigned)flattenGridCoord(c);   /* reference code */
}

__global__ void reconKeys(const Particle *particles, int n, int morton, unsigned *keys, unsigned *idx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = cellKey(getGridCell(particles[i].position), morton);   /* reference code */
    idx[i] = i;
}

__global__ void reconRanges(const unsigned *keys, int n, unsigned *start, unsigned *end) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned k = keys[i];
    if (i == 0 || keys[i - 1] != k) start[k] = i;
    if (i == n - 1 || keys[i + 1] != k) end[k] = i + 1;
}

__global__ void reconDensity(Particle *particles, const unsigned *idx, const unsigned *start,
                             const unsigned *end, int morton) {
    int pIdx = blockIdx.x * blockDim.x + threadIdx.x;
    if (pIdx >= deviceSettings.numParticles) return;
    Particle *particle = &particles[pIdx];
    int3 cell = getGridCell(particle->position);
    float density = 0.f;
    for (int dz = -1; dz < 2; dz++) {
        int sz = cell.z + dz;
        if (sz < 0 || sz >= deviceSettings.numCellsPerDim) continue;
        for (int dy = -1; dy < 2; dy++) {
            int sy = cell.y + dy;
            if (sy < 0 || sy >= deviceSettings.numCellsPerDim) continue;
            for (int dx = -1; dx < 2; dx++) {
                int sx = cell.x + dx;
                if (sx < 0 || sx >= deviceSettings.numCellsPerDim) continue;
                unsigned c = cellKey(make_int3(sx, sy, sz), morton);
                for (unsigned k = start[c]; k < end[c]; k++)
                    density += MASS * densityKernel(particle, &particles[idx[k]]);
            }
        }
    }
    particle->density = fmaxf(density, EPS_F);
    particle->pressure = fmaxf(0.f, GAS_CONSTANT * (particle->density - REST_DENSITY));
}

__global__ void reconForces(Particle *particles, const unsigned *idx, const unsigned *start,
                            const unsigned *end, int morton) {
    int pIdx = blockIdx.x * blockDim.x + threadIdx.x;
    if (pIdx >= deviceSettings.numParticles) return;
    Particle *particle = &particles[pIdx];
    int3 cell = getGridCell(particle->position);
    float3 f = make_float3(0.f, 0.f, 0.f);
    for (int dz = -1; dz < 2; dz++) {
        int sz = cell.z + dz;
        if (sz < 0 || sz >= deviceSettings.numCellsPerDim) continue;
        for (int dy = -1; dy < 2; dy++) {
            int sy = cell.y + dy;
            if (sy < 0 || sy >= deviceSettings.numCellsPerDim) continue;
            for (int dx = -1; dx < 2; dx++) {
                int sx = cell.x + dx;
                if (sx < 0 || sx >= deviceSettings.numCellsPerDim) continue;
                unsigned c = cellKey(make_int3(sx, sy, sz), morton);
                for (unsigned k = start[c]; k < end[c]; k++) {
                    Particle *neighbor = &particles[idx[k]];
                    float fPressure = -MASS * (particle->pressure + neighbor->pressure) / (2.f * neighbor->density);
                    float3 kern1 = pressureKernel(particle, neighbor);
                    f.x += kern1.x * fPressure; f.y += kern1.y * fPressure; f.z += kern1.z * fPressure;
                    float fViscosity = VISCOSITY * MASS * viscosityKernel(particle, neighbor) / neighbor->density;
                    f.x += (neighbor->velocity.x - particle->velocity.x) * fViscosity;
                    f.y += (neighbor->velocity.y - particle->velocity.y) * fViscosity;
                    f.z += (neighbor->velocity.z - particle->velocity.z) * fViscosity;
                }
            }
        }
    }
    particle->force = f;
}

struct Recon {
    Settings settings;
    Simulator *sim;
    int morton;
    unsigned table;
    unsigned *keys, *idx, *keys2, *idx2, *start, *end;
    void *tmp;
    size_t tmpBytes;
};

}  // namespace

extern "C" {

void *recon_create(int morton, int randomInit, int n, float h, float vk, float dk, float boxDim,
                   float numCellsPerDim, float timestep) {
    Recon *r = new Recon;
    r->settings = Settings{randomInit != 0, n, h, vk, dk, boxDim, numCellsPerDim, timestep};
    r->morton = morton;
    r->sim = new Simulator(&r->settings);
    srand(1);
    r->sim->setup();   /* the reference's own allocation + init + deviceSettings upload */
    int nc = (int)numCellsPerDim, bits = 0;
    while ((1 << bits) < nc) ++bits;
    r->table = morton ? (1u << (3 * bits)) : (unsigned)nc * nc * nc;
    cudaMalloc(&r->keys, 4 * (size_t)n); cudaMalloc(&r->idx, 4 * (size_t)n);
    cudaMalloc(&r->keys2, 4 * (size_t)n); cudaMalloc(&r->idx2, 4 * (size_t)n);
    cudaMalloc(&r->start, 4 * (size_t)r->table); cudaMalloc(&r->end, 4 * (size_t)r->table);
    r->tmp = nullptr; r->tmpBytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, r->tmpBytes, r->keys, r->keys2, r->idx, r->idx2, n);
    cudaMalloc(&r->tmp, r->tmpBytes);
    return cudaDeviceSynchronize() == cudaSuccess ? r : nullptr;
}

/* One step, timed like simulateAndTime (ref: simulator.cu:499-546): seconds accumulated into
 * buckets[0] (key + sort + ranges), [1] (density, forces, positions), [2] (D2H of positions). */
int recon_step_timed(void *handle, double *buckets) {
    Recon *r = (Recon *)handle;
    int n = r->settings.numParticles;
    dim3 block(MAX_THREADS_PER_BLOCK), grid((n + MAX_THREADS_PER_BLOCK - 1) / MAX_THREADS_PER_BLOCK);
    using clk = std::chrono::steady_clock;
    auto t0 = clk::now();
    int keyBits = 0;
    while ((1ull << keyBits) < r->table) ++keyBits;
    reconKeys<<<grid, block>>>(r->sim->particles, n, r->morton, r->keys, r->idx);
    cub::DeviceRadixSort::SortPairs(r->tmp, r->tmpBytes, r->keys, r->keys2, r->idx, r->idx2, n, 0, keyBits);
    cudaMemset(r->start, 0, 4 * (size_t)r->table);
    cudaMemset(r->end, 0, 4 * (size_t)r->table);
    reconRanges<<<grid, block>>>(r->keys2, n, r->start, r->end);
    cudaDeviceSynchronize();
    auto t1 = clk::now();
    reconDensity<<<grid, block>>>(r->sim->particles, r->idx2, r->start, r->end, r->morton);
    reconForces<<<grid, block>>>(r->sim->particles, r->idx2, r->start, r->end, r->morton);
    kernelUpdatePositions<<<grid, block>>>(r->sim->particles, r->sim->devicePosition);   /* reference kernel */
    cudaDeviceSynchronize();
    auto t2 = clk::now();
    cudaMemcpy(r->sim->position, r->sim->devicePosition, sizeof(float3) * (size_t)n, cudaMemcpyDeviceToHost);
    auto t3 = clk::now();
    buckets[0] += std::chrono::duration<double>(t1 - t0).count();
    buckets[1] += std::chrono::duration<double>(t2 - t1).count();
    buckets[2] += std::chrono::duration<double>(t3 - t2).count();
    cudaError_t e = cudaGetLastError();
    return (int)e;
}

const float *recon_positions(void *handle) { return (const float *)((Recon *)handle)->sim->position; }

void recon_destroy(void *handle) {
    Recon *r = (Recon *)handle;
    if (!r) return;
    cudaFree(r->keys); cudaFree(r->idx); cudaFree(r->keys2); cudaFree(r->idx2);
    cudaFree(r->start); cudaFree(r->end); cudaFree(r->tmp);
    cudaFree(r->sim->neighborGrid); cudaFree(r->sim->particles); cudaFree(r->sim->devicePosition);
    free(r->sim->position);
    delete r;
}

}  // extern "C"
