import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print(d["workload"], round(d["ms_per_pair"], 1), d["kernels_ms"].get("tq_gn"), d["kernel_ms_total"], d["quads"])
