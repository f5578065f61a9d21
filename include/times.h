// times.h -- timing record of `-m time` and its table printer.
//
// Drop-in for the reference header of the same name (ref: src/times.h:5-36): same
// `Times` layout and the same five output lines, byte for byte, so scripts that
// parse the reference's table keep working.  tests/test_host_cli.py compares the
// output with the reference's own displayTimes() where /root/reference is mounted
// and with a committed golden text elsewhere.
#pragma once

#include <cstdio>
#include <iostream>

struct Times {
    double buildGrid = 0.f;  // "Grid construction": hash + sort + cell ranges + reorder
    double sphUpdate = 0.f;  // "SPH update": density/pressure + force/integrate
    double memcpy = 0.f;     // "Data transfer": device -> host positions
    int iters = 0;
};

inline void displayTimes(Times *times) {
    // per-frame values are quotients, as in the reference (ref: times.h:13-15)
    const double avgBuildGrid = times->iters ? times->buildGrid / times->iters : 0.0;
    const double avgSphUpdate = times->iters ? times->sphUpdate / times->iters : 0.0;
    const double avgMemcpy = times->iters ? times->memcpy / times->iters : 0.0;
    char line[5][96];
    // column widths of the reference's iomanip sequence (ref: times.h:19-35)
    std::snprintf(line[0], sizeof line[0], "%-12s%18s%12s", "Operation", "Per frame", "Total");
    std::snprintf(line[1], sizeof line[1], "%s", "---------------------------------------------");
    std::snprintf(line[2], sizeof line[2], "%-11s%11.5f%15.5f", "Grid construction",
                  avgBuildGrid, times->buildGrid);
    std::snprintf(line[3], sizeof line[3], "%-12s%16.5f%15.5f", "SPH update",
                  avgSphUpdate, times->sphUpdate);
    std::snprintf(line[4], sizeof line[4], "%-12s%15.5f%15.5f", "Data transfer",
                  avgMemcpy, times->memcpy);
    std::cout.flush();
    for (auto &l : line) std::printf("%s\n", l);
    std::fflush(stdout);
}
