#!/usr/bin/env python
"""Static SASS opcode histogram of one kernel in an object file (the hot REBLUR / RELAX kernels are straight-line, fully unrolled
code, so static counts track the executed warp-instructions that bound them).

  python tools/sass_hist.py nrd_sample_b200/build/kernels_reblur_spatial.cu.o 'reblurBlurKernelILi3ELb0E' [--dump]
"""
import collections
import re
import subprocess
import sys

XU = ("MUFU", "I2F", "F2I", "FRND", "F2F", "I2FP", "F2FP")  # quarter-rate pipe on sm_100 (I2FP/F2FP are ALU but listed to see them)


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    dump = "--dump" in sys.argv
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    name, hist, total = None, collections.Counter(), 0
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            continue
        if name is None or pat not in name:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if not m:
            continue
        ins = m.group(2)
        ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
        op = ins.split()[0]
        base = op.split(".")[0]
        key = base
        if base == "MUFU":
            key = op
        hist[key] += 1
        total += 1
        if dump:
            print(m.group(1), m.group(2))
    print(f"total {total}")
    for k, v in hist.most_common():
        print(f"{v:6d} {100.0 * v / total:5.1f}%  {k}")


if __name__ == "__main__":
    main()
