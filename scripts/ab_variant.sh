#!/bin/bash
# Build a variant of libebvo_b200.so with extra -D flags for match.cu into ab/<name>.so (A/B runs on the GPU box copy it over the in-tree library).
# usage: scripts/ab_variant.sh name -DFOO=1 ...
set -e
name=$1; shift
cd "$(dirname "$0")/../edge_based_visual_odometry_b200/csrc"
mkdir -p ../../ab
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v "$@" -c match.cu -o /tmp/match_$name.o 2> /tmp/match_$name.log
grep -A2 "gn_lerp64" /tmp/match_$name.log | grep -E "spill|registers" || true
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../ab/$name.so toed.o /tmp/match_$name.o sift.o undistort.o capi.o -lcudart
