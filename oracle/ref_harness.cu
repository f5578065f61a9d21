/*
 * ref_harness.cu -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Headless C-ABI harness around the UNMODIFIED reference CUDA translation unit
 * (/root/reference/src/simulator.cu).  The reference source is #included from
 * where it lies (path passed as -DREF_SIM_CU=...); nothing of it is copied into
 * this repository.  The result, oracle/_ref/libsph_ref.so, is git-ignored but
 * travels to the GPU box, where it is
 *   - the ground truth of tests/test_gpu_reference.py (my CUDA path and the C
 *     restatement in sph_oracle.c are both compared with it), and
 *   - "the reference's own CUDA build on one B200" of bench.py --impl reference.
 *
 * `#define private public` only widens access to Simulator's four pointers so
 * that identical states can be injected / dumped; it does not change layout or
 * code generation of the reference.  Build flags are the reference's
 * (Makefile:28: -O3 -m64, no fast-math) with the arch switched to sm_100a.
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
#include <vector>

#define private public
#include REF_SIM_CU
#undef private

/* defined by display.cpp in the reference (display.cpp:19-20) */
bool mouseClicked = false;
int2 clickCoords = {0, 0};

namespace {

struct RefSim {
    Settings settings;
    Simulator *sim;
};

int sync_status() {
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaGetLastError();
    return (int)e;
}

__global__ void harnessKeys(const Particle *particles, int n, int *cells, int *flat) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int3 c = getGridCell(particles[i].position);      /* reference code */
    cells[3 * i] = c.x;
    cells[3 * i + 1] = c.y;
    cells[3 * i + 2] = c.z;
    flat[i] = flattenGridCoord(c);                    /* reference code */
}

/* Walk every list once and record which list each particle sits in. */
__global__ void harnessMembership(Particle **grid, const Particle *base, int ncell,
                                  int *listOf) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    for (Particle *p = grid[c]; p != NULL; p = p->next) listOf[p - base] = c;
}

/* Same 27-cell walk as ref simulator.cu:163-185, counting instead of summing:
 * C = candidates, K = candidates the reference's `dist2 > h2` test keeps,
 * Knz = candidates for which the reference's densityKernel() returned > 0. */
__global__ void harnessCounts(Particle *particles, Particle **grid, int n, int *K,
                              int *C, int *Knz) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Particle *particle = &particles[i];
    int3 cell = getGridCell(particle->position);
    int k = 0, c = 0, knz = 0;
    float h2 = deviceSettings.h * deviceSettings.h;
    for (int dz = -1; dz < 2; dz++) {
        int sz = cell.z + dz;
        if (sz < 0 || sz >= deviceSettings.numCellsPerDim) continue;
        for (int dy = -1; dy < 2; dy++) {
            int sy = cell.y + dy;
            if (sy < 0 || sy >= deviceSettings.numCellsPerDim) continue;
            for (int dx = -1; dx < 2; dx++) {
                int sx = cell.x + dx;
                if (sx < 0 || sx >= deviceSettings.numCellsPerDim) continue;
                Particle *nb = grid[flattenGridCoord(make_int3(sx, sy, sz))];
                while (nb != NULL) {
                    float ddx = particle->position.x - nb->position.x;
                    float ddy = particle->position.y - nb->position.y;
                    float ddz = particle->position.z - nb->position.z;
                    float d2 = ddx * ddx + ddy * ddy + ddz * ddz;
                    c++;
                    if (!(d2 > h2)) k++;
                    if (densityKernel(particle, nb) > 0.f) knz++; /* reference code */
                    nb = nb->next;
                }
            }
        }
    }
    K[i] = k;
    C[i] = c;
    Knz[i] = knz;
}

void launch_dims(int n, dim3 &grid, dim3 &block) {
    block = dim3(MAX_THREADS_PER_BLOCK);
    grid = dim3((n + MAX_THREADS_PER_BLOCK - 1) / MAX_THREADS_PER_BLOCK);
}

void reset_grid(RefSim *r) {
    int nc = (int)r->settings.numCellsPerDim;
    kernelResetGrid<<<dim3(nc, nc, nc), dim3(1)>>>(r->sim->neighborGrid);
}

}  // namespace

extern "C" {

int ref_abi_version() { return 3; }

int ref_sizeof_particle() { return (int)sizeof(Particle); }
int ref_sizeof_settings() { return (int)sizeof(Settings); }

/* Settings + Simulator + setup() exactly as main.cpp:62-66 does. */
void *ref_create(int randomInit, int n, float h, float vk, float dk, float boxDim,
                 float numCellsPerDim, float timestep) {
    RefSim *r = new RefSim;
    r->settings = Settings{randomInit != 0, n, h, vk, dk, boxDim, numCellsPerDim, timestep};
    r->sim = new Simulator(&r->settings);
    srand(1); /* keep -i random reproducible across repeated creates in one process */
    r->sim->setup();
    if (sync_status() != 0) return NULL;
    return r;
}

/* The reference destructor is unusable; free by hand. */
void ref_destroy(void *handle) {
    RefSim *r = (RefSim *)handle;
    if (!r) return;
    cudaFree(r->sim->neighborGrid);
    cudaFree(r->sim->particles);
    cudaFree(r->sim->devicePosition);
    free(r->sim->position);
    r->sim->position = NULL;
    r->sim->neighborGrid = NULL;
    r->sim->particles = NULL;
    /* Simulator object itself intentionally leaked: its dtor walks device memory */
    delete r;
}

int ref_set_state(void *handle, const float *pos, const float *vel) {
    RefSim *r = (RefSim *)handle;
    int n = r->settings.numParticles;
    Particle *tmp = (Particle *)calloc((size_t)n, sizeof(Particle));
    for (int i = 0; i < n; i++) {
        tmp[i].position = make_float3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        if (vel) tmp[i].velocity = make_float3(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]);
    }
    cudaMemcpy(r->sim->particles, tmp, sizeof(Particle) * (size_t)n, cudaMemcpyHostToDevice);
    free(tmp);
    return sync_status();
}

/* Any output may be NULL.  force/rho/prs are what the LAST step computed from
 * the pre-step positions; pos/vel are post-step. */
int ref_get_state(void *handle, float *pos, float *vel, float *force, float *rho,
                  float *prs) {
    RefSim *r = (RefSim *)handle;
    int n = r->settings.numParticles;
    Particle *tmp = (Particle *)malloc(sizeof(Particle) * (size_t)n);
    cudaMemcpy(tmp, r->sim->particles, sizeof(Particle) * (size_t)n, cudaMemcpyDeviceToHost);
    int rc = sync_status();
    for (int i = 0; i < n && rc == 0; i++) {
        if (pos) { pos[3*i] = tmp[i].position.x; pos[3*i+1] = tmp[i].position.y; pos[3*i+2] = tmp[i].position.z; }
        if (vel) { vel[3*i] = tmp[i].velocity.x; vel[3*i+1] = tmp[i].velocity.y; vel[3*i+2] = tmp[i].velocity.z; }
        if (force) { force[3*i] = tmp[i].force.x; force[3*i+1] = tmp[i].force.y; force[3*i+2] = tmp[i].force.z; }
        if (rho) rho[i] = tmp[i].density;
        if (prs) prs[i] = tmp[i].pressure;
    }
    free(tmp);
    return rc;
}

int ref_step(void *handle) {
    RefSim *r = (RefSim *)handle;
    mouseClicked = false;
    r->sim->simulate();
    return sync_status();
}

int ref_step_click(void *handle, int x, int y) {
    RefSim *r = (RefSim *)handle;
    mouseClicked = true;
    clickCoords = make_int2(x, y);
    r->sim->simulate();
    return sync_status();
}

/* buckets: buildGrid, sphUpdate, memcpy seconds accumulated (times.h:5-10) */
int ref_step_timed(void *handle, double *buckets, int *iters) {
    RefSim *r = (RefSim *)handle;
    Times t;
    t.buildGrid = buckets[0];
    t.sphUpdate = buckets[1];
    t.memcpy = buckets[2];
    t.iters = *iters;
    r->sim->simulateAndTime(&t);
    buckets[0] = t.buildGrid;
    buckets[1] = t.sphUpdate;
    buckets[2] = t.memcpy;
    *iters = t.iters;
    return sync_status();
}

/* The reference's host position buffer (original particle order), N*3 floats. */
const float *ref_positions(void *handle) {
    RefSim *r = (RefSim *)handle;
    return (const float *)r->sim->getPosition();
}

int ref_keys(void *handle, int *cells, int *flat) {
    RefSim *r = (RefSim *)handle;
    int n = r->settings.numParticles;
    int *dCells, *dFlat;
    cudaMalloc(&dCells, sizeof(int) * 3 * (size_t)n);
    cudaMalloc(&dFlat, sizeof(int) * (size_t)n);
    dim3 g, b;
    launch_dims(n, g, b);
    harnessKeys<<<g, b>>>(r->sim->particles, n, dCells, dFlat);
    cudaMemcpy(cells, dCells, sizeof(int) * 3 * (size_t)n, cudaMemcpyDeviceToHost);
    cudaMemcpy(flat, dFlat, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost);
    cudaFree(dCells);
    cudaFree(dFlat);
    return sync_status();
}

/* Builds the reference's lists with its own kernelBuildGrid, extracts list
 * membership and neighbour / candidate counts, then resets the heads with its
 * own kernelResetGrid -- state is left as before the call. */
int ref_neighbor_counts(void *handle, int *listOf, int *K, int *C, int *Knz) {
    RefSim *r = (RefSim *)handle;
    int n = r->settings.numParticles;
    int nc = (int)r->settings.numCellsPerDim;
    int ncell = nc * nc * nc;
    int *d;
    cudaMalloc(&d, sizeof(int) * 4 * (size_t)n);
    cudaMemset(d, 0xff, sizeof(int) * 4 * (size_t)n);
    dim3 g, b;
    launch_dims(n, g, b);
    kernelBuildGrid<<<g, b>>>(r->sim->particles, r->sim->neighborGrid);
    harnessMembership<<<(ncell + 127) / 128, 128>>>(r->sim->neighborGrid, r->sim->particles,
                                                    ncell, d);
    harnessCounts<<<g, b>>>(r->sim->particles, r->sim->neighborGrid, n, d + n, d + 2 * (size_t)n,
                            d + 3 * (size_t)n);
    reset_grid(r);
    int rc = sync_status();
    if (listOf) cudaMemcpy(listOf, d, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost);
    if (K) cudaMemcpy(K, d + n, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost);
    if (C) cudaMemcpy(C, d + 2 * (size_t)n, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost);
    if (Knz) cudaMemcpy(Knz, d + 3 * (size_t)n, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost);
    cudaFree(d);
    return rc ? rc : sync_status();
}

}  // extern "C"
