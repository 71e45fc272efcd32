"""Experiments only: build a variant of libnrd_b200.so HERE (no GPU needed) with extra -D flags on selected .cu files, reusing the other objects
of the main build. The variant lands in nrd_sample_b200/variants/libnrd_b200_<name>.so (git-ignored, travels to the GPU box) and is picked up
with NRD_B200_LIB=<path> (nrd_sample_b200/executor.py).
  python tools/build_variant.py hf3 "-DHF_MIN_BLOCKS=3" reblur_history_fix_stabilization.cu"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nrd_sample_b200 import build as b  # noqa: E402

name, flags, files = sys.argv[1], sys.argv[2].split(), sys.argv[3:]
b.build()
out_dir = os.path.join(b.PKG, "variants")
obj_dir = os.path.join(out_dir, name)
os.makedirs(obj_dir, exist_ok=True)
host, cuda = b.sources()
objs, procs = [], []
for src in host + cuda:
    main_obj = os.path.join(b.OBJ, os.path.relpath(src, b.CSRC).replace(os.sep, "_") + ".o")
    if os.path.basename(src) in files:
        obj = os.path.join(obj_dir, os.path.basename(main_obj))
        procs.append(subprocess.Popen([b._nvcc()] + b.NVCC_FLAGS + flags + ["-c", src, "-o", obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        objs.append(obj)
    else:
        objs.append(main_obj)
for p in procs:
    out, _ = p.communicate()
    for line in out.splitlines():
        if "error" in line or ("registers" in line):
            print(line.strip()[:160])
    assert p.returncode == 0, out[-2000:]
lib = os.path.join(out_dir, f"libnrd_b200_{name}.so")
subprocess.check_call([b._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-o", lib] + objs)
print(lib)
