"""Measures the FP64 pipe peaks (DFMA and DMMA m8n8k4) that the product roofline divides by.
Writes gpurun_out/fp64_peaks.json; copy into profiles/ when re-measured."""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genfer_b200

ctx = genfer_b200.Context(0)
res = {}
for kind, name in ((0, "dfma"), (1, "dmma_m8n8k4")):
    best = 0.0
    for iters in (2048, 16384, 65536):
        fl, ms = ctx.fp64_peak_probe(kind, iters)
        best = max(best, fl)
        res.setdefault(name + "_runs", []).append({"iters": iters, "tflops": fl / 1e12, "ms": ms})
    res[name + "_tflops"] = best / 1e12
try:
    q = subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.sm,clocks.max.sm,power.draw", "--format=csv,noheader"],
                       capture_output=True, text=True).stdout.strip()
    res["nvidia_smi"] = q
except Exception as e:  # pragma: no cover
    res["nvidia_smi"] = str(e)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/fp64_peaks.json", "w"), indent=1)
print(json.dumps(res))
ctx.close()
