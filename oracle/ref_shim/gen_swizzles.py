"""TEST INFRASTRUCTURE. Writes hlsl_swizzles.inc: the union members that give hlsl_cpu.h's vectors their HLSL swizzles
(.xy, .zwxy, .rgb ...). Run once: python oracle/ref_shim/gen_swizzles.py"""
import itertools
import os

out = []
for n in (2, 3, 4):
    out.append(f"#define HLSL_SWIZZLES_{n}(T) \\")
    lines = []
    for names in ("xyzw", "rgba"):
        for k in (2, 3, 4):
            for combo in itertools.product(range(n), repeat=k):
                lines.append(f"    Swz<T, {n}, {', '.join(map(str, combo))}> {''.join(names[i] for i in combo)};")
    out.append(" \\\n".join(lines))
    out.append("")
open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "hlsl_swizzles.inc"), "w").write("\n".join(out) + "\n")
