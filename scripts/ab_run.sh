#!/bin/bash
# On the GPU box: run scripts/ab2.py for every variant library in ab/ (or the names given), swapping each over the in-tree library.
# usage: scripts/ab_run.sh out.jsonl [frames distinct reps parity] -- name1 name2 ...
out=$1; shift
args=()
while [ $# -gt 0 ] && [ "$1" != "--" ]; do args+=("$1"); shift; done
shift
names=("$@")
[ ${#names[@]} -eq 0 ] && names=($(ls ab/*.so | xargs -n1 basename | sed 's/\.so$//'))
lib=edge_based_visual_odometry_b200/libebvo_b200.so
cp $lib /tmp/lib_orig.so
: > $out
for n in "${names[@]}"; do
  cp ab/$n.so $lib
  timeout 300 python scripts/ab2.py $n "${args[@]}" >> $out 2> gpurun_out/ab2_$n.err || echo "{\"tag\": \"$n\", \"failed\": true}" >> $out
done
cp /tmp/lib_orig.so $lib
cat $out
