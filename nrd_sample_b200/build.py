"""In-tree build of the product library `nrd_sample_b200/libnrd_b200.so`:
host pass graph (csrc/host/*.cpp) + CUDA executor and kernels (csrc/*.cu, csrc/kernels/*.cu), sm_100a only.
nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libnrd_b200.so")
OBJ = os.path.join(PKG, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo",
    # approximate division / sqrt / transcendentals (MUFU) and FTZ: the chain is ALU-bound, and strict-mode parity against the
    # fp32 oracle holds with a wide margin (profiles/README.md: precise 3.92 ms -> fast-math 2.00 ms per 1440p frame)
    "-use_fast_math",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off,-Wall",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
] + os.environ.get("NRD_B200_NVCC_EXTRA", "").split()  # experiments only, e.g. NRD_B200_NVCC_EXTRA=-use_fast_math
# HBM-bound data-movement kernels whose arithmetic should match the host build of the same header bit for bit: IEEE division / sqrt, no FMA contraction
PRECISE_FILES = {"frontend.cu"}
HOST_FLAGS = ["-std=c++17", "-O2", "-fPIC", "-fvisibility=hidden", "-ffp-contract=off", "-Wall", "-Wextra"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def sources():
    host = sorted(glob.glob(os.path.join(CSRC, "host", "*.cpp")))
    cuda = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "kernels", "*.cu")))
    return host, cuda


def headers():
    return sorted(glob.glob(os.path.join(CSRC, "**", "*.h"), recursive=True) + glob.glob(os.path.join(CSRC, "**", "*.cuh"), recursive=True) +
                  glob.glob(os.path.join(os.path.dirname(PKG), "include", "*.h")))


def host_library_path() -> str:
    """Path of the library exporting the nrd:: ABI (the same .so that holds the CUDA executor)."""
    build()
    return LIB


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    host, cuda = sources()
    return any(os.path.getmtime(s) > t for s in host + cuda + headers() + [os.path.abspath(__file__)])


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    host, cuda = sources()
    hdr_time = max([os.path.getmtime(h) for h in headers()] + [os.path.getmtime(os.path.abspath(__file__))])
    objs = []
    log = []
    procs = []
    for src in host + cuda:
        obj = os.path.join(OBJ, os.path.relpath(src, CSRC).replace(os.sep, "_") + ".o")
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_time):
            continue
        if src.endswith(".cu"):
            flags = [f for f in NVCC_FLAGS if f != "-use_fast_math"] + ["-fmad=false"] if os.path.basename(src) in PRECISE_FILES else NVCC_FLAGS
            cmd = [_nvcc()] + flags + ["-c", src, "-o", obj]
        else:
            cmd = ["g++"] + HOST_FLAGS + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"== {os.path.relpath(src, PKG)}\n{out}")
        if p.returncode != 0:
            failed = True
            sys.stderr.write(log[-1])
    with open(os.path.join(OBJ, "build.log"), "w") as f:
        f.write("\n".join(log))
    if failed:
        raise RuntimeError("nrd_b200 build failed (see above)")
    link = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-o", LIB] + objs
    subprocess.check_call(link)
    if verbose:
        sys.stdout.write("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
