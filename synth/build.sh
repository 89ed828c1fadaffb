#!/bin/sh
# Builds the host copy of the synthetic LiDAR raycaster (test/bench input generator).
set -e
cd "$(dirname "$0")"
g++ -std=c++17 -O2 -ffp-contract=off -fPIC -shared -pthread -o libsynth.so synth_host.cpp
echo "built synth/libsynth.so"
