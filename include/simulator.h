// simulator.h -- source-compatible replacement for the reference's public header
// (ref: src/simulator.h:1-74), backed by the B200-native library libsph_b200.so.
//
// A caller written against the reference -- src/main.cpp (time mode) and
// src/display.cpp (free mode) -- compiles against this header unchanged: same constant
// names, same positional `Settings` aggregate, same `class Simulator` members with
// the same meaning:
//
//   Simulator(Settings*)   records the pointer only; the caller keeps ownership and the
//                          object must outlive the simulator (ref: simulator.cu:370-375)
//   setup()                allocate + initialise + upload            (ref: 411-460)
//   simulate()             one timestep, host positions refreshed, consumes the
//                          mouseClicked / clickCoords hand-off        (ref: 462-497)
//   simulateAndTime(t)     one timestep with the three wall-clock buckets (ref: 499-546)
//   getPosition()          N float3 in ORIGINAL particle order, simulator-owned,
//                          valid after a step returns                 (ref: 407-409)
//   moveParticles(int2)    declared but never defined in the reference (simulator.h:73);
//                          here: the mouse push of simulator.cu:329-367
//
// What is deliberately absent: `struct Particle` (ref: simulator.h:33-51).  It is the
// reference's private AoS record with an intrusive list pointer; no caller outside
// simulator.cu touches it and this implementation stores particles as sorted SoA
// float4 arrays on the device instead.
#pragma once

#include <cuda_runtime.h>  // float3, int2 (the reference header pulls them in the same way)
#include <stdio.h>

#include "times.h"

// Physics constants under the reference's names (ref: simulator.h:6-12; typed constants
// instead of macros -- callers only use them in expressions, e.g. main.cpp:60-61).  The
// device code keeps its own copies in csrc/sph_common.cuh.
constexpr float PI = 3.14159265f;
constexpr float MASS = 0.02f;
constexpr float REST_DENSITY = 1000.f;
constexpr float GAS_CONSTANT = 1.f;
constexpr float VISCOSITY = 1.f;
constexpr float GRAVITY = -9.8f;
constexpr float ELASTICITY = 0.5f;

// Window rectangle, in pixels, that maps onto the box for mouse pushes
// (ref: simulator.h:14-17, used by display.cpp:24-25).
constexpr int BOX_MIN_X = 200, BOX_MAX_X = 600;
constexpr int BOX_MIN_Y = 150, BOX_MAX_Y = 450;

// Field order and types are the reference's (ref: simulator.h:19-31): callers
// aggregate-initialise it positionally (main.cpp:62-63).  numCellsPerDim really is a
// float there.  sizeof == 32; include/sph_b200.h's SphSettings has the same layout.
struct Settings {
    bool randomInit;        // -i random | grid
    int numParticles;       // -n
    float h;                // smoothing length == cell edge
    float v_kernel_coeff;   // 45 / (pi h^6)
    float d_kernel_coeff;   // 315 / (64 pi h^9)
    float boxDim;           // box edge
    float numCellsPerDim;   // boxDim / h
    float timestep;
};

struct sph_sim;      // opaque handles of the C ABI (include/sph_b200.h)
struct sph_cluster;

class Simulator {
  public:
    const Settings *settings;

    Simulator(Settings *settings);
    virtual ~Simulator();

    void setup();
    const float3 *getPosition();
    void simulate();
    void simulateAndTime(Times *times);
    void moveParticles(int2 mouse_pos);

    // additive: last error of the underlying C ABI (0 == ok); the reference has no
    // error reporting at all, so ignoring this reproduces its behaviour
    int status() const { return lastStatus; }
    // additive: the C-ABI handle, for callers that want the extra entry points of sph_b200.h
    // (state I/O, parity hooks); null when the simulator runs on several GPUs
    sph_sim *handle() const { return impl; }
    // additive: with SPH_GPUS=N (./sph -g N) the box is split into N z-slabs, one per GPU, behind
    // the same five methods; this is the handle of that cluster (sph_cluster_* of sph_b200.h)
    sph_cluster *clusterHandle() const { return cluster; }

  private:
    sph_sim *impl;
    sph_cluster *cluster = nullptr;
    float *clusterPositions = nullptr;   // getPosition() buffer of the multi-GPU path
    int lastStatus = 0;
    void note(int rc, const char *what);
};
