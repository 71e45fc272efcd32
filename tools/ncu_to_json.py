"""From an .ncu-rep of ONE steady-state frame ( or several: the LAST launch of every kernel name is taken ) to the two tables bench.py reads:
profiles/dram_traffic.json ( dram__bytes_read.sum + dram__bytes_write.sum per launch ) and profiles/inst_counts.json ( smsp__inst_executed.sum per launch ),
keyed by workload ( "2560x1440" for REBLUR, "relax 2560x1440", "sigma 2560x1440" ) and pass name as in the dispatch stream.
  python tools/ncu_to_json.py gpurun_out/ev/reblur_full.ncu-rep reblur 2560x1440
Kernels that belong to a pass without a dispatch of their own are added to that pass ( REBLUR's geometry plane -> Pre-pass; SIGMA's copy rides in Blur )."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PASS_OF = {
    "reblur": [("ClassifyTiles", "Classify tiles"), ("GeometryPlane", "Pre-pass"), ("PrePass", "Pre-pass"), ("TemporalAccumulation", "Temporal accumulation"),
               ("HistoryFix", "History fix"), ("PostBlur", "Post-blur"), ("Blur", "Blur"), ("TemporalStabilization", "Temporal stabilization"),
               ("HitDistReconstruction", "Hit distance reconstruction")],
    "relax": [("ClassifyTiles", "Classify tiles"), ("PrePass", "Pre-pass"), ("TemporalAccumulation", "Temporal accumulation"), ("HistoryFix", "History fix"),
              ("HistoryClamping", "History clamping"), ("AtrousSmem", "A-trous (SMEM)"), ("Atrous", "A-trous"), ("AntiFirefly", "Anti-firefly"), ("Copy", "Copy")],
    "sigma": [("ClassifyTiles", "Classify tiles"), ("SmoothTiles", "Smooth tiles"), ("Copy", "Copy"), ("Blur", "Blur"), ("TemporalStabilization", "Temporal stabilization")],
}


def main():
    rep, family, size = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}

    def val(r, m):
        v = float(r[idx[m]].replace(",", ""))
        u = units[idx[m]]
        return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)

    launches = []   # ( pass, dram bytes, warp instructions ) in launch order
    for r in data:
        name = r[idx["Kernel Name"]]
        p = next((pn for key, pn in PASS_OF[family] if re.search(family + key, name, re.I)), None)
        if p is None:
            continue
        launches.append((p, name, val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"), val(r, "smsp__inst_executed.sum")))
    # the last frame: walk back from the end until a kernel name repeats a pass that may only occur once per frame
    per_frame = {"sigma": 2}.get(family, 1)   # SIGMA runs its blur kernel twice ( Blur, Post-blur ), RELAX its a-trous kernel four times
    last, seen = [], {}
    for p, name, dram, inst in reversed(launches):
        limit = 4 if p == "A-trous" else (2 if (family == "sigma" and p == "Blur") else 1)
        if family == "reblur" and p == "Pre-pass":
            limit = 2   # plane + pre-pass
        if seen.get(p, 0) >= limit:
            break
        seen[p] = seen.get(p, 0) + 1
        last.append((p, name, dram, inst))
    last.reverse()
    dram_t, inst_t, count = {}, {}, {}
    blur_seen = 0
    for p, name, dram, inst in last:
        if family == "sigma" and p == "Blur":
            blur_seen += 1
            p = "Blur" if blur_seen == 1 else "Post-blur"
        dram_t[p] = dram_t.get(p, 0.0) + dram
        inst_t[p] = inst_t.get(p, 0.0) + inst
        count[p] = count.get(p, 0) + 1
    if "A-trous" in count:   # per launch, like the per-dispatch times of the bench line
        dram_t["A-trous"] /= count["A-trous"]
        inst_t["A-trous"] /= count["A-trous"]
    key = size if family == "reblur" else f"{family} {size}"
    for fname, table in (("dram_traffic.json", dram_t), ("inst_counts.json", inst_t)):
        path = os.path.join(ROOT, "profiles", fname)
        doc = json.load(open(path)) if os.path.exists(path) else {}
        doc[key] = {k: int(round(v)) for k, v in table.items()}
        doc["source"] = ("ncu --set full --clock-control none, per launch, last steady-state frame of tools/profile_frame.py 2560 1440 6 [relax|sigma] on the round-2 build "
                         "( tools/gpu_evidence.sh; summaries: profiles/r2_*_ncu_summary.txt ). dram_traffic: dram__bytes_read.sum + dram__bytes_write.sum; "
                         "inst_counts: smsp__inst_executed.sum ( warp-instructions ). REBLUR Pre-pass includes the geometry-plane kernel; SIGMA Blur includes the folded Copy.")
        json.dump(doc, open(path, "w"), indent=1)
    print(key, "dram MB", {k: round(v / 1e6, 1) for k, v in dram_t.items()})
    print(key, "Minst", {k: round(v / 1e6, 1) for k, v in inst_t.items()})


if __name__ == "__main__":
    main()
