#!/bin/bash
# The user-facing CLI (the reference's `./sph -n -i -m time`, 100 timed iterations) on one B200: the same
# table the reference prints (times.h), for BASELINE configs[0..2] and the library's switches.
S=./cudafluidsimulator_b200/sph
run() { echo "== $*"; env "${ENVV[@]}" $S "$@" 2>&1 | tail -5; }
ENVV=(A=1); run -n 10000 -i grid -m time
ENVV=(A=1); run -n 1000000 -i random -m time
ENVV=(A=1); run -n 16000000 -i grid -b 25.6 -c 256 -m time
ENVV=(SPH_PIPELINE_READBACK=1); echo "(SPH_PIPELINE_READBACK=1)"; run -n 16000000 -i grid -b 25.6 -c 256 -m time
ENVV=(SPH_SORT=radix); echo "(SPH_SORT=radix)"; run -n 16000000 -i grid -b 25.6 -c 256 -m time
