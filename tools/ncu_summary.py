"""Summarise an .ncu-rep (read here, no GPU needed): per-kernel headline metrics + opcode mix + top stall reasons.
  python tools/ncu_summary.py gpurun_out/x.ncu-rep [kernel-regex-for-opcode-mix]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
M = [("dur_us", "gpu__time_duration.sum", 1e3), ("regs", "launch__registers_per_thread", 1), ("occ%", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
     ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1), ("Minst", "smsp__inst_executed.sum", 1e-6), ("L1hit%", "l1tex__t_sector_hit_rate.pct", 1),
     ("L2hit%", "lts__t_sector_hit_rate.pct", 1), ("dramR_MB", "dram__bytes_read.sum", 1), ("dramW_MB", "dram__bytes_write.sum", 1),
     ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1), ("fma%", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", 1),
     ("alu%", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", 1), ("xu%", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1),
     ("lsu%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1)]
print("kernel".ljust(44) + "".join(n.rjust(10) for n, _, _ in M))
for r in data:
    name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "")[:43]
    vals = []
    for n, m, s in M:
        try:
            v = float(r[idx[m]].replace(",", "")) * s
            if units[idx[m]] == "byte": v /= 1e6
            elif units[idx[m]] == "Kbyte": v /= 1e3
            elif units[idx[m]] == "Gbyte": v *= 1e3
            if m == "gpu__time_duration.sum":
                u = units[idx[m]]
                v = float(r[idx[m]]) * {"ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(u, 1.0)
            vals.append(f"{v:10.1f}")
        except Exception:
            vals.append("n/a".rjust(10))
    print(name.ljust(44) + "".join(vals))
STALLS = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")] or \
         [h for h in hdr if "warp_issue_stalled" in h and h.endswith(".pct")]
for r in data:
    name = re.sub(r"\(.*", "", r[idx["Kernel Name"]])[:40]
    st = []
    for h in STALLS:
        try:
            st.append((float(r[idx[h]]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        except Exception:
            pass
    st.sort(reverse=True)
    print(name.ljust(42), " ".join(f"{n}={v:.2f}" for v, n in st[:6]))
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + sys.argv[2]], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]
    i_src, i_ex = h.index("Source"), h.index("Instructions Executed")
    c, tot, nlines = collections.Counter(), 0, 0
    for r in rows[2:]:
        if len(r) <= i_ex or r[0].startswith("Kernel Name") or r[0] == "Address":
            continue
        try:
            n = int(r[i_ex])
        except Exception:
            continue
        s = re.sub(r"^@!?U?P\d+\s+", "", r[i_src].strip())
        op = s.split()[0].split(".")[0] if s else "?"
        c[op] += n
        tot += n
        nlines += 1
    print(f"opcode mix for {sys.argv[2]}: {tot/1e6:.1f} M warp-inst, {nlines} SASS lines")
    print("  " + "  ".join(f"{op} {100*n/tot:.1f}%" for op, n in c.most_common(24)))
