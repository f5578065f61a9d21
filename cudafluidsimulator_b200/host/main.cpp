// main.cpp -- the `./sph -n <N> -i <random|grid> -m <free|time>` driver
// (ref: src/main.cpp:12-83), headless-capable.
//
// Flags, defaults, messages and exit codes are the reference's: -n (default 1000), -i grid,
// -m time, -? usage, exit 1 on a bad value; time mode = 100 x simulateAndTime() then
// displayTimes().  Additive flags (reference behaviour when absent):
//   -b <boxDim>  -c <cellsPerDim>   scale the domain beyond the hard-coded 10 / 100 box
//                                   (ref: main.cpp:62-63), needed above 109^3 particles
//   -g <gpus>                       split the box into z-slabs over this many GPUs (one process, peer-to-peer
//                                   halo exchange and migration inside the library; Simulator picks it up)
//   -k <flat|morton>                cell-key form of the sort
//   -s <steps>                      iterations in time mode (default 100, main.cpp:69)
//   -f <frames>                     frames to run in headless free mode (default 600)
//   -o <prefix>                     headless free mode: write every frame as <prefix>_NNNN.ppm (the box
//                                   outline in white, particles as blue points, x right / y up -- the
//                                   headless stand-in for display.cpp's GL_POINTS view)
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
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "simulator.h"
#include "sph_b200.h"

#ifdef SPH_WITH_GLUT
void startVisualization(Simulator *simulator);  // display.cpp of the caller
#endif

namespace {

struct Options {
    int particles = 1000;       // ref: main.cpp:21
    bool random = false;        // ref: main.cpp:22 (grid)
    bool timed = true;          // ref: main.cpp:23 (time)
    float box = 10.f;           // ref: main.cpp:63
    float cells = 100.f;        // ref: main.cpp:63
    int steps = 100;            // ref: main.cpp:69
    int frames = 600;
    std::string load, dump, frames_prefix;
};

void usage() {
    static const char *lines[] = {
        "Program Options:",
        "  -n  <NUM_PARTICLES>    Number of particles to simulate",
        "  -i  <random/grid>      Initialization mode: random or grid",
        "  -m  <free/time>        Execution mode: free or timed",
        "  -b  <BOX_DIM>          (extension) box edge length, default 10",
        "  -c  <CELLS_PER_DIM>    (extension) grid cells per dimension, default 100",
        "  -g  <GPUS>             (extension) z-slab decomposition over this many GPUs, default 1",
        "  -k  <flat/morton>      (extension) cell key used by the sort, default flat",
        "  -s  <STEPS>            (extension) timed iterations, default 100",
        "  -l/-d <FILE>           (extension) load initial / dump final state",
        "  -f/-o <N>/<PREFIX>     (extension) headless free mode: frames to run / PPM frame prefix",
        "  -?                     This message",
    };
    for (const char *l : lines) printf("%s\n", l);
}

// An option that takes one of two words; anything else is the reference's error path
// (message on stdout, usage, exit status 1 -- ref: main.cpp:31-37, 41-47).
bool choice(char flag, const std::string &value, const char *yes, const char *no, bool &out) {
    if (value != yes && value != no) {
        std::cout << "Invalid argument for option -" << flag << ": " << value << std::endl;
        usage();
        return false;
    }
    out = (value == yes);
    return true;
}

// returns -1 to continue, otherwise the exit status
int parse(int argc, char **argv, Options &o) {
    for (int c; (c = getopt(argc, argv, "n:i:m:b:c:k:s:f:l:d:o:g:?")) != -1;) {
        const std::string v = optarg ? optarg : "";
        bool morton = false;
        switch (c) {
        case 'n': o.particles = std::stoi(v); break;
        case 'i': if (!choice('i', v, "random", "grid", o.random)) return 1; break;
        case 'm': if (!choice('m', v, "time", "free", o.timed)) return 1; break;
        case 'k':
            if (!choice('k', v, "morton", "flat", morton)) return 1;
            setenv("SPH_KEY_MODE", morton ? "morton" : "flat", 1);
            break;
        case 'g': setenv("SPH_GPUS", v.c_str(), 1); break;
        case 'b': o.box = std::stof(v); break;
        case 'c': o.cells = (float)std::stoi(v); break;
        case 's': o.steps = std::stoi(v); break;
        case 'f': o.frames = std::stoi(v); break;
        case 'l': o.load = v; break;
        case 'd': o.dump = v; break;
        case 'o': o.frames_prefix = v; break;
        default: usage(); return 1;   // '?' and unknown flags (ref: main.cpp:51-53)
        }
    }
    return -1;
}

const char kMagic[8] = {'S', 'P', 'H', 'B', '2', '0', '0', 0};

bool load_state(const std::string &path, int n, std::vector<float> &pos, std::vector<float> &vel) {
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

bool dump_state(const std::string &path, int n, const std::vector<float> &pos, const std::vector<float> &vel) {
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return false;
    int32_t m = n;
    bool ok = fwrite(kMagic, 1, 8, f) == 8 && fwrite(&m, 4, 1, f) == 1 &&
              fwrite(pos.data(), 4, pos.size(), f) == pos.size() && fwrite(vel.data(), 4, vel.size(), f) == vel.size();
    return fclose(f) == 0 && ok;
}

// One frame as a binary PPM: orthographic view along -z (x right, y up), box outline white,
// particles blue -- what display.cpp draws with GL_LINES / GL_POINTS, minus the perspective.
bool write_frame(const std::string &prefix, int frame, const float3 *p, int n, float box) {
    const int W = 480, H = 480, margin = 20;
    std::vector<unsigned char> img((size_t)W * H * 3, 0);
    const float scale = (W - 2 * margin) / box;
    auto put = [&](int x, int y, unsigned char r, unsigned char g, unsigned char b) {
        if (x < 0 || x >= W || y < 0 || y >= H) return;
        unsigned char *q = &img[((size_t)(H - 1 - y) * W + x) * 3];
        q[0] = r; q[1] = g; q[2] = b;
    };
    for (int t = 0; t <= W - 2 * margin; ++t) {
        put(margin + t, margin, 255, 255, 255); put(margin + t, H - margin, 255, 255, 255);
        put(margin, margin + t, 255, 255, 255); put(W - margin, margin + t, 255, 255, 255);
    }
    for (int i = 0; i < n; ++i)
        put(margin + (int)(p[i].x * scale), margin + (int)(p[i].y * scale), 40, 90, 255);
    char name[512];
    snprintf(name, sizeof name, "%s_%04d.ppm", prefix.c_str(), frame);
    FILE *f = fopen(name, "wb");
    if (!f) return false;
    fprintf(f, "P6\n%d %d\n255\n", W, H);
    const bool ok = fwrite(img.data(), 1, img.size(), f) == img.size();
    return fclose(f) == 0 && ok;
}

}  // namespace

int main(int argc, char **argv) {
    Options o;
    const int early = parse(argc, argv, o);
    if (early >= 0) return early;

    // kernel coefficients exactly as the reference computes them (ref: main.cpp:57-61):
    // pow(float, int) is the double overload, rounded back to float
    const float h = .1f;
    const float h6 = pow(h, 6), h9 = pow(h, 9);
    Settings settings = {o.random, o.particles, h, 45.f / (PI * h6), 315.f / (64.f * PI * h9),
                         o.box,    o.cells,     .01};

    Simulator *simulator = new Simulator(&settings);
    simulator->setup();
    if (simulator->status() != 0) return 2;  // the reference would carry on silently
    if ((!o.load.empty() || !o.dump.empty()) && !simulator->handle()) {
        fprintf(stderr, "sph: -l / -d need the single-GPU simulator (drop -g)\n");
        return 2;
    }
    if (!o.load.empty()) {
        std::vector<float> pos, vel;
        if (!load_state(o.load, o.particles, pos, vel) ||
            sph_set_state(simulator->handle(), pos.data(), vel.data()) != 0) {
            fprintf(stderr, "sph: cannot load %d particles from %s: %s\n", o.particles, o.load.c_str(), sph_last_error());
            return 2;
        }
    }

    if (o.timed) {
        Times times;
        for (int i = 0; i < o.steps; i++) {
            simulator->simulateAndTime(&times);
            if (simulator->status() != 0) return 2;
        }
        displayTimes(&times);
    } else {
#ifdef SPH_WITH_GLUT
        glutInit(&argc, argv);
        startVisualization(simulator);
#else
        const auto t0 = std::chrono::steady_clock::now();
        double checksum = 0.0;
        for (int f = 0; f < o.frames; f++) {
            simulator->simulate();  // what display() does per frame (ref: display.cpp:36-37)
            if (simulator->status() != 0) return 2;
            const float3 *p = simulator->getPosition();
            checksum += p[f % (o.particles > 0 ? o.particles : 1)].y;
            if (!o.frames_prefix.empty() && !write_frame(o.frames_prefix, f, p, o.particles, o.box)) {
                fprintf(stderr, "sph: cannot write frame %d to %s_*.ppm\n", f, o.frames_prefix.c_str());
                return 2;
            }
        }
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("free mode (headless build, no GLUT): %d frames in %.3f s = %.1f frames/s (checksum %.6f)\n",
               o.frames, dt, o.frames / dt, checksum);
#endif
    }
    if (!o.dump.empty()) {
        std::vector<float> pos((size_t)3 * o.particles), vel((size_t)3 * o.particles);
        if (sph_get_state(simulator->handle(), pos.data(), vel.data()) != 0 ||
            !dump_state(o.dump, o.particles, pos, vel)) {
            fprintf(stderr, "sph: cannot dump the state to %s: %s\n", o.dump.c_str(), sph_last_error());
            return 2;
        }
    }
    return 0;
}
