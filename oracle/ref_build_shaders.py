#!/usr/bin/env python
"""TEST INFRASTRUCTURE. Compiles the reference's compute shaders — the .cs.hlsl files where they lie under /root/reference — as C++ into
oracle/_ref/libnrd_refshaders.so (see oracle/ref_shim/hlsl_cpu.h). Per permutation:

    gcc -E (C mode: the shaders' `#ifndef __cplusplus` branches stay active)  |  ref_shim/hlsl2cpp.py  |  g++ -x c++ -c -

Nothing of the shader text is written into the repository: the pipeline runs through pipes, objects go to a temporary directory and only
the linked .so lands in oracle/_ref/ (git-ignored; it travels to the GPU box as a prebuilt file).
Usage: python oracle/ref_build_shaders.py [-j N] [identifier-substring ...]"""
import concurrent.futures as cf
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "ref_shim"))
import hlsl2cpp  # noqa: E402

REF = os.environ.get("NRD_REFERENCE_ROOT", "/root/reference")
SHADERS = os.path.join(REF, "External/NRD/Shaders")
ML = os.path.join(REF, "External/NRIFramework/External/MathLib")
OUT = os.path.join(HERE, "_ref")
SHIM = os.path.join(HERE, "ref_shim")

# `shaderIdentifier`s ( NRDDescs.h:452 ): file name + the -D permutation of Shaders.cfg, exactly as the host library emits them
IDENTIFIERS = [
    "Clear.cs.hlsl|FLOAT=0", "Clear.cs.hlsl|FLOAT=1",
    "REBLUR_ClassifyTiles.cs.hlsl",
    "REBLUR_HitDistReconstruction.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=RADIANCE|MODE_5X5=0",
    "REBLUR_HitDistReconstruction.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=RADIANCE|MODE_5X5=1",
    "REBLUR_PrePass.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=RADIANCE",
    "REBLUR_TemporalAccumulation.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=RADIANCE",
    "REBLUR_HistoryFix.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=RADIANCE",
    "REBLUR_Blur.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=RADIANCE",
    "REBLUR_PostBlur.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=RADIANCE|TEMPORAL_STABILIZATION=0",
    "REBLUR_PostBlur.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=RADIANCE|TEMPORAL_STABILIZATION=1",
    "REBLUR_TemporalStabilization.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=RADIANCE",
    "REBLUR_SplitScreen.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=RADIANCE",
    "SIGMA_ClassifyTiles.cs.hlsl|TRANSLUCENCY=0", "SIGMA_SmoothTiles.cs.hlsl", "SIGMA_Copy.cs.hlsl",
    "SIGMA_Blur.cs.hlsl|TRANSLUCENCY=0|FIRST_PASS=1", "SIGMA_Blur.cs.hlsl|TRANSLUCENCY=0|FIRST_PASS=0",
    "SIGMA_TemporalStabilization.cs.hlsl|TRANSLUCENCY=0", "SIGMA_SplitScreen.cs.hlsl|TRANSLUCENCY=0",
    "RELAX_ClassifyTiles.cs.hlsl",
    "RELAX_HitDistReconstruction.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=RADIANCE|MODE_5X5=0",
    "RELAX_HitDistReconstruction.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=RADIANCE|MODE_5X5=1",
    "RELAX_PrePass.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=SH",
    "RELAX_TemporalAccumulation.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=SH",
    "RELAX_HistoryFix.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=SH",
    "RELAX_HistoryClamping.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=SH",
    "RELAX_Copy.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=SH",
    "RELAX_AntiFirefly.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=SH",
    "RELAX_AtrousSmem.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=SH",
    "RELAX_Atrous.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=SH",
    "RELAX_SplitScreen.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=SH",
    # RELAX_DIFFUSE_SPECULAR ( what NRDSample instantiates as shipped )
    "RELAX_PrePass.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=RADIANCE",
    "RELAX_TemporalAccumulation.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=RADIANCE",
    "RELAX_HistoryFix.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=RADIANCE",
    "RELAX_HistoryClamping.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=RADIANCE",
    "RELAX_Copy.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=RADIANCE",
    "RELAX_AntiFirefly.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=RADIANCE",
    "RELAX_AtrousSmem.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=RADIANCE",
    "RELAX_Atrous.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=RADIANCE",
    "RELAX_SplitScreen.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=RADIANCE",
    # SIGMA_SHADOW_TRANSLUCENCY
    "SIGMA_ClassifyTiles.cs.hlsl|TRANSLUCENCY=1", "SIGMA_Blur.cs.hlsl|TRANSLUCENCY=1|FIRST_PASS=1", "SIGMA_Blur.cs.hlsl|TRANSLUCENCY=1|FIRST_PASS=0",
    "SIGMA_TemporalStabilization.cs.hlsl|TRANSLUCENCY=1", "SIGMA_SplitScreen.cs.hlsl|TRANSLUCENCY=1",
    # REBLUR_DIFFUSE / REBLUR_SPECULAR
] + [f"{f}|NRD_SIGNAL={sig}|NRD_MODE=RADIANCE{suffix}" for sig in ("DIFF", "SPEC") for f, suffix in (
    ("REBLUR_HitDistReconstruction.cs.hlsl", "|MODE_5X5=0"), ("REBLUR_HitDistReconstruction.cs.hlsl", "|MODE_5X5=1"), ("REBLUR_PrePass.cs.hlsl", ""),
    ("REBLUR_TemporalAccumulation.cs.hlsl", ""), ("REBLUR_HistoryFix.cs.hlsl", ""), ("REBLUR_Blur.cs.hlsl", ""), ("REBLUR_PostBlur.cs.hlsl", "|TEMPORAL_STABILIZATION=0"),
    ("REBLUR_PostBlur.cs.hlsl", "|TEMPORAL_STABILIZATION=1"), ("REBLUR_TemporalStabilization.cs.hlsl", ""), ("REBLUR_SplitScreen.cs.hlsl", ""))] + [
    # RELAX_DIFFUSE / RELAX_DIFFUSE_SH / RELAX_SPECULAR / RELAX_SPECULAR_SH
] + [f"RELAX_HitDistReconstruction.cs.hlsl|NRD_SIGNAL={sig}|NRD_MODE=RADIANCE|MODE_5X5={m}" for sig in ("DIFF", "SPEC") for m in (0, 1)] + [
    f"RELAX_{f}.cs.hlsl|NRD_SIGNAL={sig}|NRD_MODE={mode}" for sig in ("DIFF", "SPEC") for mode in ("RADIANCE", "SH")
    for f in ("PrePass", "TemporalAccumulation", "HistoryFix", "HistoryClamping", "Copy", "AntiFirefly", "AtrousSmem", "Atrous", "SplitScreen")] + [
    # REBLUR_DIFFUSE_SH / REBLUR_SPECULAR_SH / REBLUR_DIFFUSE_SPECULAR_SH ( hit-distance reconstruction runs the RADIANCE permutation )
] + [f"{f}|NRD_SIGNAL={sig}|NRD_MODE=SH{suffix}" for sig in ("DIFF", "SPEC", "BOTH") for f, suffix in (
    ("REBLUR_PrePass.cs.hlsl", ""), ("REBLUR_TemporalAccumulation.cs.hlsl", ""), ("REBLUR_HistoryFix.cs.hlsl", ""), ("REBLUR_Blur.cs.hlsl", ""),
    ("REBLUR_PostBlur.cs.hlsl", "|TEMPORAL_STABILIZATION=0"), ("REBLUR_PostBlur.cs.hlsl", "|TEMPORAL_STABILIZATION=1"), ("REBLUR_TemporalStabilization.cs.hlsl", ""),
    ("REBLUR_SplitScreen.cs.hlsl", ""))] + [
    # REBLUR_*_OCCLUSION ( NRD_MODE = OCCLUSION: no pre-pass, no stabilization ) and REBLUR_DIFFUSE_DIRECTIONAL_OCCLUSION ( NRD_MODE = DO, diffuse only )
] + [f"{f}|NRD_SIGNAL={sig}|NRD_MODE=OCCLUSION{suffix}" for sig in ("DIFF", "SPEC", "BOTH") for f, suffix in (
    ("REBLUR_HitDistReconstruction.cs.hlsl", "|MODE_5X5=0"), ("REBLUR_HitDistReconstruction.cs.hlsl", "|MODE_5X5=1"), ("REBLUR_TemporalAccumulation.cs.hlsl", ""),
    ("REBLUR_HistoryFix.cs.hlsl", ""), ("REBLUR_Blur.cs.hlsl", ""), ("REBLUR_PostBlur.cs.hlsl", "|TEMPORAL_STABILIZATION=0"))] + [
    f"{f}|NRD_SIGNAL=DIFF|NRD_MODE=DO{suffix}" for f, suffix in (
        ("REBLUR_PrePass.cs.hlsl", ""), ("REBLUR_TemporalAccumulation.cs.hlsl", ""), ("REBLUR_HistoryFix.cs.hlsl", ""), ("REBLUR_Blur.cs.hlsl", ""),
        ("REBLUR_PostBlur.cs.hlsl", "|TEMPORAL_STABILIZATION=0"), ("REBLUR_PostBlur.cs.hlsl", "|TEMPORAL_STABILIZATION=1"), ("REBLUR_TemporalStabilization.cs.hlsl", ""))] + [
    # validation overlays ( CommonSettings::enableValidation )
    "REBLUR_Validation.cs.hlsl", "RELAX_Validation.cs.hlsl",
    # REFERENCE
    "REFERENCE_TemporalAccumulation.cs.hlsl", "REFERENCE_Copy.cs.hlsl",
    # ours: calls the application-side functions of the reference's NRD.hlsli ( oracle/ref_shim/Shaders/NRD_FrontEndProbe.cs.hlsl )
    "NRD_FrontEndProbe.cs.hlsl",
]
CXXFLAGS = ["-std=c++20", "-O2", "-fPIC", "-fopenmp", "-ffp-contract=off", "-fvisibility=hidden", "-w", "-fmax-errors=25", "-I", SHIM]


def cxx_for(identifier: str) -> str:
    parts = identifier.split("|")
    defines = [f"-D{p}" for p in parts[1:]]
    cpp = ["gcc", "-E", "-P", "-undef", "-nostdinc", "-x", "c", "-include", os.path.join(SHIM, "hlsl_engine_macros.h"), "-I", SHADERS, "-I", ML, "-I", SHIM] + defines + [
        os.path.join(SHADERS if os.path.exists(os.path.join(SHADERS, parts[0])) else os.path.join(SHIM, "Shaders"), parts[0])]
    pre = subprocess.run(cpp, capture_output=True, text=True)
    if pre.returncode:
        raise RuntimeError(f"{identifier}: preprocessing failed\n{pre.stderr}")
    ns = "s_" + "".join(c if c.isalnum() else "_" for c in identifier)
    return hlsl2cpp.transform(pre.stdout, identifier, ns)


def compile_one(identifier: str, obj: str):
    src = cxx_for(identifier)
    r = subprocess.run(["g++", "-x", "c++"] + CXXFLAGS + ["-c", "-o", obj, "-"], input=src, capture_output=True, text=True)
    return identifier, r.returncode, r.stderr, src


def main():
    if not os.path.isdir(SHADERS):
        print("reference not present, skipping the shader build")
        return 0
    jobs, picks, dump = os.cpu_count() or 4, [], False
    argv = sys.argv[1:]
    while argv:
        a = argv.pop(0)
        if a == "-j":
            jobs = int(argv.pop(0))
        elif a == "--show":    # debugging aid: print the numbered C++ around the first errors ( to the terminal only )
            dump = True
        else:
            picks.append(a)
    ids = [i for i in IDENTIFIERS if not picks or any(p in i for p in picks)]
    os.makedirs(OUT, exist_ok=True)
    failed = 0
    with tempfile.TemporaryDirectory() as tmp:
        objs = [os.path.join(tmp, f"s{k}.o") for k in range(len(ids))]
        with cf.ThreadPoolExecutor(max_workers=jobs) as pool:
            for identifier, rc, err, src in pool.map(compile_one, ids, objs):
                if rc:
                    failed += 1
                    print(f"FAILED {identifier}\n{err[:6000]}")
                    if dump:
                        lines = src.split("\n")
                        import re
                        for ln in sorted({int(m) for m in re.findall(r"<stdin>:(\d+):", err)})[:12]:
                            print(f"  {ln}: {lines[ln - 1].strip()[:200]}")
                else:
                    print(f"ok     {identifier}")
        if failed:
            print(f"{failed} of {len(ids)} shaders failed")
            return 1
        if picks:
            print("partial build: not linking")
            return 0
        so = os.path.join(OUT, "libnrd_refshaders.so")
        r = subprocess.run(["g++"] + CXXFLAGS + ["-shared", os.path.join(SHIM, "hlsl_runtime.cpp")] + objs + ["-o", so], capture_output=True, text=True)
        if r.returncode:
            print(r.stderr)
            return 1
        print(f"built {so} ({len(ids)} shaders)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
