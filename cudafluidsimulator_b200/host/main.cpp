// main.cpp -- the `./sph -n <N> -i <random|grid> -m <free|time>` driver
// (ref: src/main.cpp:12-83), headless-capable.
//
// Flags, defaults and validation are the reference's: -n (default 1000), -i grid,
// -m time, -? usage, exit 1 on a bad value; time mode = 100 x simulateAndTime() then
// displayTimes().  Additive flags (reference behaviour when absent):
//   -b <boxDim>  -c <cellsPerDim>   scale the domain beyond the hard-coded 10 / 100 box
//                                   (ref: main.cpp:62-63), needed above 109^3 particles
//   -k <flat|morton>                cell-key form of the sort
//   -s <steps>                      iterations in time mode (default 100, main.cpp:69)
//   -f <frames>                     frames to run in headless free mode (default 600)
//   -l <file> / -d <file>           load the initial state from / dump the final state to a file
//                                   (state I/O; the reference has none).  Format: "SPHB200\0",
//                                   int32 N, N x 3 float32 positions, N x 3 float32 velocities
// Free mode needs GLUT/OpenGL (ref: display.cpp), which this image does not have; when
// built without SPH_WITH_GLUT it runs the same per-frame sequence display() does
// (simulate() + getPosition()) without drawing and reports frames per second.
#include <unistd.h>

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <string>

#include "simulator.h"
#include "sph_b200.h"

#include <cstring>
#include <vector>

#ifdef SPH_WITH_GLUT
void startVisualization(Simulator *simulator);  // display.cpp of the caller
#endif

static void usage() {
    printf("Program Options:\n");
    printf("  -n  <NUM_PARTICLES>    Number of particles to simulate\n");
    printf("  -i  <random/grid>      Initialization mode: random or grid\n");
    printf("  -m  <free/time>        Execution mode: free or timed\n");
    printf("  -b  <BOX_DIM>          (extension) box edge length, default 10\n");
    printf("  -c  <CELLS_PER_DIM>    (extension) grid cells per dimension, default 100\n");
    printf("  -k  <flat/morton>      (extension) cell key used by the sort, default flat\n");
    printf("  -s  <STEPS>            (extension) timed iterations, default 100\n");
    printf("  -?                     This message\n");
}

static bool one_of(const std::string &v, const char *a, const char *b) { return v == a || v == b; }

static const char kMagic[8] = {'S', 'P', 'H', 'B', '2', '0', '0', 0};

static bool load_state(const std::string &path, int n, std::vector<float> &pos, std::vector<float> &vel) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    char magic[8];
    int32_t m = 0;
    bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, kMagic, 8) == 0 && fread(&m, 4, 1, f) == 1 && m == n;
    pos.resize((size_t)3 * n);
    vel.resize((size_t)3 * n);
    ok = ok && fread(pos.data(), 4, pos.size(), f) == pos.size() && fread(vel.data(), 4, vel.size(), f) == vel.size();
    fclose(f);
    return ok;
}

static bool dump_state(const std::string &path, int n, const std::vector<float> &pos, const std::vector<float> &vel) {
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return false;
    int32_t m = n;
    bool ok = fwrite(kMagic, 1, 8, f) == 8 && fwrite(&m, 4, 1, f) == 1 &&
              fwrite(pos.data(), 4, pos.size(), f) == pos.size() && fwrite(vel.data(), 4, vel.size(), f) == vel.size();
    return fclose(f) == 0 && ok;
}

int main(int argc, char **argv) {
    int numParticles = 1000;
    bool randomInit = false;
    bool benchmark = true;
    float boxDim = 10.f;
    float cells = 100;
    int numIters = 100;
    int frames = 600;
    std::string loadPath, dumpPath;
    int opt;

    while ((opt = getopt(argc, argv, "n:i:m:b:c:k:s:f:l:d:?")) != -1) {
        const std::string arg = optarg ? optarg : "";
        switch (opt) {
        case 'n':
            numParticles = std::stoi(arg);
            break;
        case 'i':
            if (!one_of(arg, "random", "grid")) {
                std::cout << "Invalid argument for option -i: " << arg << std::endl;
                usage();
                return 1;
            }
            randomInit = (arg == "random");
            break;
        case 'm':
            if (!one_of(arg, "time", "free")) {
                std::cout << "Invalid argument for option -m: " << arg << std::endl;
                usage();
                return 1;
            }
            benchmark = (arg == "time");
            break;
        case 'b':
            boxDim = std::stof(arg);
            break;
        case 'c':
            cells = (float)std::stoi(arg);
            break;
        case 'k':
            if (!one_of(arg, "flat", "morton")) {
                std::cout << "Invalid argument for option -k: " << arg << std::endl;
                usage();
                return 1;
            }
            setenv("SPH_KEY_MODE", arg.c_str(), 1);
            break;
        case 's':
            numIters = std::stoi(arg);
            break;
        case 'f':
            frames = std::stoi(arg);
            break;
        case 'l':
            loadPath = arg;
            break;
        case 'd':
            dumpPath = arg;
            break;
        case '?':
            usage();
            return 1;
        }
    }

    // ref: main.cpp:57-63
    float h = .1f;
    float h_pow_6 = pow(h, 6);
    float h_pow_9 = pow(h, 9);
    float v_kernel_coeff = 45.f / (PI * h_pow_6);
    float d_kernel_coeff = 315.f / (64.f * PI * h_pow_9);
    Settings settings = {randomInit,     numParticles, h,     v_kernel_coeff,
                         d_kernel_coeff, boxDim,       cells, .01};

    Simulator *simulator = new Simulator(&settings);
    simulator->setup();
    if (simulator->status() != 0) return 2;  // the reference would carry on silently
    if (!loadPath.empty()) {
        std::vector<float> pos, vel;
        if (!load_state(loadPath, numParticles, pos, vel) ||
            sph_set_state(simulator->handle(), pos.data(), vel.data()) != 0) {
            fprintf(stderr, "sph: cannot load %d particles from %s: %s\n", numParticles, loadPath.c_str(), sph_last_error());
            return 2;
        }
    }

    if (benchmark) {
        Times times;
        for (int i = 0; i < numIters; i++) {
            simulator->simulateAndTime(&times);
            if (simulator->status() != 0) return 2;
        }
        displayTimes(&times);
    } else {
#ifdef SPH_WITH_GLUT
        glutInit(&argc, argv);
        startVisualization(simulator);
#else
        auto t0 = std::chrono::steady_clock::now();
        double checksum = 0.0;
        for (int f = 0; f < frames; f++) {
            simulator->simulate();  // what display() does per frame (ref: display.cpp:36-37)
            if (simulator->status() != 0) return 2;
            const float3 *p = simulator->getPosition();
            checksum += p[f % (numParticles > 0 ? numParticles : 1)].y;
        }
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("free mode (headless build, no GLUT): %d frames in %.3f s = %.1f frames/s (checksum %.6f)\n",
               frames, dt, frames / dt, checksum);
#endif
    }
    if (!dumpPath.empty()) {
        std::vector<float> pos((size_t)3 * numParticles), vel((size_t)3 * numParticles);
        if (sph_get_state(simulator->handle(), pos.data(), vel.data()) != 0 ||
            !dump_state(dumpPath, numParticles, pos, vel)) {
            fprintf(stderr, "sph: cannot dump the state to %s: %s\n", dumpPath.c_str(), sph_last_error());
            return 2;
        }
    }
    return 0;
}
