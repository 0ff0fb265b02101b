import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
B=d["config"]["frames_per_gpu_per_step"]
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"toed frac",round(d["roofline"]["frac"],3))
print(" ".join(f"{k}={v['ms_per_step']/B*1000:.1f}" for k,v in d["kernels"].items()))
