#!/bin/bash
# Builds the library with -DEBVO_CHECKED (device-side bounds assertions on tile / list / pool indices: the stand-in for
# compute-sanitizer memcheck, which is closed on the GPU pool) into ab/checked.so.  On the GPU box (second form) the checked
# library replaces the in-tree one for the whole -m gpu suite and the sanitize workload: any violated assertion fails the call.
# usage (here): scripts/checked_run.sh build      usage (GPU box): scripts/checked_run.sh run
set -e
cd "$(dirname "$0")/.."
if [ "$1" = "build" ]; then
  mkdir -p ab /tmp/chk
  for f in toed match sift undistort capi; do
    extra=""; [ $f = sift -o $f = undistort ] && extra="-fmad=false"
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DEBVO_CHECKED $extra -c edge_based_visual_odometry_b200/csrc/$f.cu -o /tmp/chk/$f.o &
  done
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ab/checked.so /tmp/chk/toed.o /tmp/chk/match.o /tmp/chk/sift.o /tmp/chk/undistort.o /tmp/chk/capi.o -lcudart
  ls -la ab/checked.so
else
  lib=edge_based_visual_odometry_b200/libebvo_b200.so
  cp $lib /tmp/lib_orig.so; cp ab/checked.so $lib
  python scripts/sanitize_run.py 2>&1 | tail -4
  python -m pytest tests -m gpu -q 2>&1 | tail -4
  cp /tmp/lib_orig.so $lib
fi
