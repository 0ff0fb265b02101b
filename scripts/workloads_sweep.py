"""Runs bench.py on the other BASELINE.json configurations (EuRoC shape, KITTI SIFT-on, 4K density sweep) and writes
profiles/<tag>_workloads.jsonl + .md.  usage (GPU box): python scripts/workloads_sweep.py r01"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
runs = [["--workload", "euroc", "--steps", "5"], ["--workload", "kitti", "--sift", "--steps", "5"]] + \
       [["--workload", "4k", "--density", str(d), "--steps", "3"] for d in (0.25, 0.5, 1.0, 2.0)]
lines = []
for r in runs:
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline", "--warmup", "3", *r], capture_output=True, text=True)
    if out.returncode != 0:
        print("FAILED", r, out.stderr[-400:], file=sys.stderr)
        continue
    lines.append(json.loads(out.stdout.strip().splitlines()[-1]))
    print(lines[-1]["config"]["workload"][:80], round(lines[-1]["value"], 1), flush=True)
dst = os.path.join(ROOT, "gpurun_out")
os.makedirs(dst, exist_ok=True)
with open(os.path.join(dst, f"{tag}_workloads.jsonl"), "w") as f:
    for d in lines:
        f.write(json.dumps(d) + "\n")
md = [f"# Round {tag[1:].lstrip('0')}: the other BASELINE.json configurations (one B200, `python bench.py --workload ... [--density d] [--sift]`)", "",
      f"Full JSON lines: `{tag}_workloads.jsonl` (scripts/workloads_sweep.py).  These are extra measurements; the contract line (`{tag}_bench_default.json`) is the",
      "KITTI workload.  `toed` = fraction of the FP32 FMA peak (algorithmic 1636 flop/px), `gn` = fraction of the FP64 FMA peak",
      "(algorithmic 5.3 kflop per Gauss-Newton iteration).", "",
      "| workload | frames/step | edges/image | mates/frame | frames/s (resident) | frames/s (e2e) | toed | gn |", "|---|---|---|---|---|---|---|---|"]
for d in lines:
    c, ps = d["config"], d["roofline"]["per_stage"]
    md.append(f"| {c['workload'][:110]} | {c['frames_per_gpu_per_step']} | {c['edges_per_image']:.0f} | {c['mates_per_frame']:.0f} | {d['value']:.1f} | {d['e2e']['value']:.1f} | {ps['toed']['frac']:.2f} | {ps['gauss_newton']['frac']:.2f} |")
open(os.path.join(dst, f"{tag}_workloads.md"), "w").write("\n".join(md) + "\n")
