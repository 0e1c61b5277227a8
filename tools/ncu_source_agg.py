"""Aggregate the source page of an .ncu-rep per SASS opcode: instruction mix, stall samples, shared-memory wavefronts."""
import csv
import subprocess
import sys

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[1]
H = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
f = lambda r, k: float(r[H[k]] or 0) if k in H else 0.0
tot_samples = sum(f(r, "# Samples") for r in body)
agg = {}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
stall_tot = {h: 0.0 for h in stall_cols}
for r in body:
    toks = r[H["Source"]].split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "LDG", "RED", "STS", "LDGSTS")) else op.split(".")[0]
    a = agg.setdefault(op, [0, 0, 0, 0, 0])
    a[0] += f(r, "Instructions Executed")
    a[1] += f(r, "# Samples")
    a[2] += f(r, "L1 Wavefronts Shared")
    a[3] += f(r, "L1 Wavefronts Shared Ideal")
    a[4] += f(r, "L1 Wavefronts Shared Excessive")
    for h in stall_cols:
        stall_tot[h] += f(r, h)
print(rows[0][1], "total samples", tot_samples)
tot_inst = sum(a[0] for a in agg.values())
for op, a in sorted(agg.items(), key=lambda t: -t[1][0])[:16]:
    print(f"  {op:12s} inst {a[0]:.3e} ({100 * a[0] / tot_inst:5.1f}%) samples {100 * a[1] / tot_samples:5.1f}%  smem wavefronts {a[2]:.3e} ideal {a[3]:.3e} excessive {a[4]:.3e}")
s = sum(stall_tot.values())
print("  stall samples:", {k: round(100 * v / s, 1) for k, v in sorted(stall_tot.items(), key=lambda t: -t[1])[:9]})
