#!/bin/sh
# Builds the CPU oracle (test infrastructure) with the reference's release flags
# (-O3 -DNDEBUG, no -march: slam_viz/CMakeLists.txt:13); FP contraction off.
# The reference itself cannot be compiled here (needs Eigen), so there is no
# oracle/_ref build: see DESIGN.md "Oracle".
set -e
cd "$(dirname "$0")"
g++ -std=c++17 -O3 -DNDEBUG -ffp-contract=off -fPIC -shared -pthread -o liboracle.so slam_oracle.cpp
echo "built oracle/liboracle.so"
