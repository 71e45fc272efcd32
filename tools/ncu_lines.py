"""Attribute executed warp instructions of one kernel (from an .ncu-rep SASS page) to CUDA source lines (nvdisasm -g line info).
  python tools/ncu_lines.py gpurun_out/x.ncu-rep build/kernels_reblur_spatial.cu.o reblurBlurKernel [top] [mangled-substring]
The optional mangled substring picks one template instantiation in the object file, e.g. reblurBlurKernelILi3E for <SIGNAL_BOTH>."""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, obj, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
section = sys.argv[5] if len(sys.argv) > 5 else kern
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# walk the kernel's section: remember the last "//## File "...", line N" marker before each instruction offset
line_of = {}
inside, cur = False, None
for l in dis:
    if l.startswith("//---------------------") and ".text." in l:
        inside = section in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
i_addr, i_ex, i_src = h.index("Address"), h.index("Instructions Executed"), h.index("Source")
base = None
per_line, total = collections.Counter(), 0
ops = collections.defaultdict(collections.Counter)
for r in rows[2:]:
    if len(r) <= i_ex or r[0] in ("Kernel Name", "Address"):
        continue
    try:
        a, n = int(r[i_addr], 16), int(r[i_ex])
    except Exception:
        continue
    if base is None:
        base = a
    key = line_of.get(a - base)
    per_line[key] += n
    total += n
    op = re.sub(r"^@!?U?P\d+\s+", "", r[i_src].strip()).split()[0].split(".")[0]
    ops[key][op] += n
files = {}
print(f"{kern}: {total/1e6:.1f} M warp instructions")
for key, n in per_line.most_common(top):
    text = ""
    if key:
        for root in ("nrd_sample_b200/csrc/kernels", "nrd_sample_b200/csrc/host", "nrd_sample_b200/csrc"):
            p = os.path.join(root, key[0])
            if os.path.exists(p):
                files.setdefault(p, open(p).read().splitlines())
                text = files[p][key[1] - 1].strip()[:110]
                break
    mix = " ".join(f"{o}:{c*100//n}" for o, c in ops[key].most_common(4))
    print(f"{100*n/total:5.1f}%  {key[0] if key else '?':28s}:{key[1] if key else 0:<4d} {text}   [{mix}]")
