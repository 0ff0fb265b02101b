# GPU box: the measurements of the committed round-2 state (one GPU).  Outputs under gpurun_out/.
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -c 200 gpurun_out/r02_bench_default.err
python bench.py --sift --steps 5 --warmup 3 --no-cpu-baseline --strong-frames 0 > gpurun_out/r02_bench_sift.json 2>/dev/null
python scripts/workloads_sweep.py r02 2>&1 | tail -7
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
bash scripts/capture_profiles.sh r02 2>&1 | tail -4
