"""Summarise .ncu-rep captures into small, reviewable files under profiles/:
    python tools/ncu_summary.py gpurun_out/r01_gemm_b1024.ncu-rep profiles/r01_gemm_b1024
writes <out>_metrics.csv (one row per captured kernel, the metrics the roofline arithmetic uses)
and <out>_hot.txt (hottest SASS lines by stall samples)."""
import csv
import io
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [w for w in WANT if w in idx]
    with open(out + "_metrics.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols)
        w.writerow([units[idx[c]] for c in cols])
        for r in rows[2:]:
            w.writerow([r[idx[c]][:120] for c in cols])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    lines = list(csv.reader(io.StringIO(src)))
    with open(out + "_hot.txt", "w") as f:
        kern, h, data = None, None, []

        def flush():
            if not data:
                return
            i_s, i_src, i_ex = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
            tot = sum(int(r[i_s]) for r in data)
            f.write(f"== {kern[:110]}\n   stall samples {tot}, warp instructions {sum(int(r[i_ex]) for r in data)}\n")
            top = sorted(range(len(data)), key=lambda k: -int(data[k][i_s]))[:18]
            for k in sorted(top):
                r = data[k]
                f.write(f"   {k:5d} {100.0 * int(r[i_s]) / max(tot, 1):5.1f}%  ex={r[i_ex]:>10}  {r[i_src].strip()[:84]}\n")
        for r in lines:
            if r and r[0] == "Kernel Name":
                flush()
                kern, h, data = r[1], None, []
            elif r and r[0] == "Address":
                h = r
            elif h and len(r) == len(h):
                data.append(r)
        flush()


if __name__ == "__main__":
    main()
