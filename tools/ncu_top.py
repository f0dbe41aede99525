"""Summarise an ncu report: key raw metrics and the most-sampled SASS instructions of the first kernel."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum"]
for i, h in enumerate(hdr):
    if any(h == k or (k in h and "pipe_tensor" in k) for k in keys):
        print(f"{h} [{units[i]}]: {[r[i] for r in data]}")
for i, h in enumerate(hdr):
    if "issue_stalled" in h and "per_issue_active" in h:
        v = float(data[0][i]) if data[0][i] else 0
        if v > 0.3:
            print(f"  stall {h.split('stalled_')[1].split('_per')[0]}: {v:.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = list(csv.reader(src.splitlines()))
# first kernel block only
start = next(i for i, l in enumerate(lines) if l and l[0] == "Address")
h = lines[start]
isamp, isrc, iexec = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
body = []
for l in lines[start + 1:]:
    if not l or l[0] in ("Kernel Name", "Address"):
        break
    body.append(l)
tot = sum(int(l[isamp] or 0) for l in body)
print("total samples", tot, "instructions", len(body))
ranked = sorted(enumerate(body), key=lambda t: -int(t[1][isamp] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for idx, l in sorted(ranked):
    print(f"{idx:5d} {int(l[isamp] or 0):6d} {100.0 * int(l[isamp] or 0) / max(tot, 1):5.1f}%  x{l[iexec]:>9}  {l[isrc].strip()[:90]}")
