"""Top stall sites of a kernel from an ncu report: ncu_top.py report.ncu-rep kernel_regex [n]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kern = None
hdr = None
data = []
for r in rows:
    if r and r[0] == "Kernel Name":
        if data: break
        kern = r[1]; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        data.append(r)
i_s = hdr.index("# Samples"); i_src = hdr.index("Source"); i_ex = hdr.index("Instructions Executed")
tot = sum(int(r[i_s]) for r in data)
print(kern[:100], "samples", tot, "instr", sum(int(r[i_ex]) for r in data))
# stall reason columns
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") or h.lower().startswith("warp stall")]
order = sorted(range(len(data)), key=lambda k: -int(data[k][i_s]))[:n]
for k in sorted(order):
    r = data[k]
    print(f"{k:5d} {int(r[i_s]):7d} {100.0*int(r[i_s])/max(tot,1):5.1f}% ex={r[i_ex]:>9} {r[i_src].strip()[:90]}")
