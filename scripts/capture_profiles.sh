#!/bin/bash
# On the GPU box: the ncu evidence of a round (run AFTER the same commands exited 0 without ncu).  Outputs under gpurun_out/.
# usage: scripts/capture_profiles.sh r02
tag=${1:-r02}
set -x
python scripts/prof_run.py 8 2 0 > /dev/null || exit 1
# launch list of a short bench run (per-launch durations are cold-cache and serialised: shares, not absolutes, are comparable)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench_b16.csv \
    python bench.py --steps 2 --warmup 3 --batch 16 --no-cpu-baseline --strong-frames 0 > gpurun_out/${tag}_launches_bench.log 2>&1
# full capture of the dominant kernel and of the secondary kernels (second pass of an 8-frame batch)
ncu --set full --clock-control none --import-source on -k regex:gn_lerp64 -s 1 -c 1 -o gpurun_out/${tag}_gn_lerp64 -f python scripts/prof_run.py 8 2 0 > gpurun_out/${tag}_ncu_gn.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:toed_grad_nms|toed_refine|gate_kernel|patch_kernel|ncc_bnb|cluster8|cluster_kernel|ncc2_best" -s 9 -c 9 \
    -o gpurun_out/${tag}_secondary -f python scripts/prof_run.py 8 2 0 > gpurun_out/${tag}_ncu_secondary.log 2>&1
# the SIFT-on kernels (descriptors, gate) and the NCC kernel in its SIFT-on form
python scripts/prof_run.py 8 2 0 1 > /dev/null || exit 1
ncu --set full --clock-control none --import-source on -k "regex:sift_desc|sift_gate|sift_blur" -s 3 -c 3 \
    -o gpurun_out/${tag}_sift -f python scripts/prof_run.py 8 2 0 1 > gpurun_out/${tag}_ncu_sift.log 2>&1
ls -la gpurun_out/${tag}_*.ncu-rep
