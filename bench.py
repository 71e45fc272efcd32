#!/usr/bin/env python
"""Headline benchmark: denoised Mpixels/s of the REBLUR_DIFFUSE_SPECULAR pass chain at 2560x1440 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--width 2560 --height 1440] [--denoiser reblur|reblur_sh|relax|sigma]

A "step" is one nrdcuDenoise of one synthetic frame (7 passes: classify tiles, pre-pass, temporal accumulation, history fix, blur,
post-blur, temporal stabilisation) in steady state.

* `value`: K steps with the frame's inputs already in HBM (CUDA events on the launch stream, barrier + synchronize on both sides, max
  over ranks).
* `e2e`: the same K steps through the host-frame entry point nrdcuDenoiseHostFrames — the renderer-facing staging blocks in pinned,
  NUMA-local host memory -> ONE H2D copy -> 7 kernels -> ONE D2H copy of both denoised outputs, every step, uploads / downloads of
  neighbouring steps overlapping the kernels. `e2e.per_texture` is the older nrdcuDenoiseHostPipelined (one 2D copy per texture from
  caller-owned pinned buffers), `e2e.serial` nrdcuDenoiseHost (everything back to back on one stream).
* `vs_baseline`: the NRD README's 2.55 ms per 1440p frame (RTX 4080) was taken with HitDistanceReconstructionMode::AREA_3X3, so the ratio
  is computed from a second timed leg with that setting on inputs with one traced lobe per pixel (`baseline_leg`), not from `value`.
* N > 1 (torchrun): every rank denoises its own independent stream of frames (replicas, no data-path collective — BASELINE.json
  config 4, "64 independent 1440p frames"), `value` is the aggregate; in the same run the ranks then denoise ONE 3840x2160 frame as N strips
  with seam rows pushed over NVLink peer memory (BASELINE.json config 3) and the line carries it as `tiled_4k`.

`--impl reference`: the reference's denoisers are HLSL compute shaders with no CPU or CUDA path; this arm executes the reference's OWN
shaders compiled as C++ (oracle/_ref/libnrd_refshaders.so, built in the build container from /root/reference) on all host threads, driven
by the reference's OWN host library (oracle/_ref/libnrd_ref.so), at the SAME resolution and settings, rank 0 only. Falls back to the
oracle port (and says so) when those libraries were not shipped.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "denoised Mpixels/s (REBLUR diff+spec, 1440p)"
UNIT = "Mpixels/s"
# Compulsory bytes per pixel per pass (every bound texel read once, every output written once; SURVEY.md App. B / DESIGN.md)
PASS_BYTES_PER_PIXEL = {
    "Classify tiles": 4, "Pre-pass": 42, "Temporal accumulation": 94, "History fix": 52, "Blur": 46, "Post-blur": 46, "Temporal stabilization": 66,
}
RING = 4  # distinct frames cycled through (4 x 118 MB of inputs at 1440p >> 126 MB L2)

# Secondary workloads (`--denoiser relax|sigma`): BASELINE.json configs 2 and 0, same harness, own metric name. Compulsory bytes per pixel per
# pass from SURVEY.md App. B; published RTX 4080 times from External/NRD/README.md:30-34.
WORKLOADS = {
    "reblur": dict(denoiser="REBLUR_DIFFUSE_SPECULAR", frame="reblur_frame", metric=METRIC, published_ms=2.55,
                   outputs=(("OUT_DIFF_RADIANCE_HITDIST", "RGBA16_SFLOAT"), ("OUT_SPEC_RADIANCE_HITDIST", "RGBA16_SFLOAT")),
                   pass_bytes=PASS_BYTES_PER_PIXEL, what="REBLUR_DIFFUSE_SPECULAR full pass chain, {w}x{h}, synthetic 1spp noisy radiance + G-buffer"),
    # NRD_MODE = SH: a second RGBA16F per lobe through every pass ( +8 B/px per lobe per texture read or written; no published number )
    "reblur_sh": dict(denoiser="REBLUR_DIFFUSE_SPECULAR_SH", frame="reblur_frame_sh", metric="denoised Mpixels/s (REBLUR diff+spec SH, 1440p)", published_ms=None,
                      outputs=(("OUT_DIFF_SH0", "RGBA16_SFLOAT"), ("OUT_DIFF_SH1", "RGBA16_SFLOAT"), ("OUT_SPEC_SH0", "RGBA16_SFLOAT"), ("OUT_SPEC_SH1", "RGBA16_SFLOAT")),
                      pass_bytes={"Classify tiles": 4, "Pre-pass": 74, "Temporal accumulation": 142, "History fix": 84, "Blur": 78, "Post-blur": 78, "Temporal stabilization": 98},
                      what="REBLUR_DIFFUSE_SPECULAR_SH full pass chain, {w}x{h}, synthetic 1spp noisy SH radiance + G-buffer"),
    "relax": dict(denoiser="RELAX_DIFFUSE_SPECULAR_SH", frame="relax_frame", metric="denoised Mpixels/s (RELAX diff+spec SH, 1440p)", published_ms=4.80,
                  outputs=(("OUT_DIFF_SH0", "RGBA16_SFLOAT"), ("OUT_DIFF_SH1", "RGBA16_SFLOAT"), ("OUT_SPEC_SH0", "RGBA16_SFLOAT"), ("OUT_SPEC_SH1", "RGBA16_SFLOAT")),
                  pass_bytes={"Classify tiles": 4, "Pre-pass": 72, "Temporal accumulation": 200, "History fix": 5, "History clamping": 150, "A-trous (SMEM)": 83, "A-trous": 74},
                  what="RELAX_DIFFUSE_SPECULAR_SH full pass chain (5 a-trous iterations), {w}x{h}, synthetic SH radiance + G-buffer"),
    "sigma": dict(denoiser="SIGMA_SHADOW", frame="sigma_frame", metric="denoised Mpixels/s (SIGMA shadow, 1440p)", published_ms=0.40,
                  outputs=(("OUT_SHADOW_TRANSLUCENCY", "R8_UNORM"),),
                  pass_bytes={"Classify tiles": 6, "Smooth tiles": 0, "Copy": 10, "Blur": 13, "Post-blur": 14, "Temporal stabilization": 25},
                  what="SIGMA_SHADOW full pass chain, {w}x{h}, synthetic 1spp penumbra + G-buffer"),
}
INPUT_FORMATS = {"IN_VIEWZ": "R32_SFLOAT", "IN_NORMAL_ROUGHNESS": "R10_G10_B10_A2_UNORM", "IN_PENUMBRA": "R16_SFLOAT"}   # everything else RGBA16_SFLOAT


def usable_cores() -> int:
    """Host threads this process may really use: the scheduler affinity capped by the cgroup CPU quota (the GPU boxes expose
    128 logical CPUs with a 16-CPU quota — 128 OpenMP threads would only oversubscribe it)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return max(1, n)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", float(d.get("sm_max_mhz", 1965.0))
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def config_of(args, world):
    """The `config` object both arms print: same workload, resolution, settings and keys, so the driver's same-config check compares like with like."""
    wl = WORKLOADS[args.denoiser]
    recon = args.hitdist_reconstruction
    return {"workload": wl["what"].format(w=args.width, h=args.height), "denoiser": wl["denoiser"], "resolution": [args.width, args.height],
            "settings": "library defaults" if recon == "off" else f"library defaults + hitDistanceReconstructionMode = AREA_{recon.upper()} (the NRD README's setting), one lobe traced per pixel",
            "streams_per_gpu": 1, "parallelism": f"replicas x{world} (independent frame streams, no collective)",
            "l2_policy": f"ring of {RING} distinct frames of inputs + the pools: larger than the 126 MB L2 at 1440p"}


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons of the GPU, sampled while the timed region runs: through NVML every 5 ms ( the timed region of a short chain is a few tens of
    milliseconds: spawning nvidia-smi every 200 ms caught one sample of it ), falling back to nvidia-smi when the NVML binding is not importable."""

    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40), ("sw_power_cap", 0x4))   # nvmlClocksEventReason* bits

    def __init__(self, index: int, uuid: str = None):
        super().__init__(daemon=True)
        self.index, self.uuid, self.samples, self._halt, self.source = index, uuid, [], threading.Event(), "nvml"
        self._nv = None
        try:   # NVML is initialised HERE, before the timed region: nvmlInit alone takes longer than a short chain's 40 steps
            import pynvml
            pynvml.nvmlInit()
            h = None
            if self.uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(self.uuid if self.uuid.startswith("GPU-") else "GPU-" + self.uuid)
                except Exception:
                    h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            self._nv = (pynvml, h, pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM), get_reasons)
        except Exception:
            self._nv = None

    def _sample_nvml(self):
        pynvml, h, mx, get_reasons = self._nv
        bits = int(get_reasons(h))
        self.samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), mx, [n for n, b in self.REASONS if bits & b]))

    def _nvml(self):
        if self._nv is None:
            raise RuntimeError("no NVML")
        while not self._halt.is_set():
            self._sample_nvml()
            self._halt.wait(0.002)

    def _smi(self):
        self.source = "nvidia-smi"
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append((int(float(f[0])), int(float(f[1])), [n for n, v in zip(names, f[3:7]) if v.lower().startswith("active")]))
            except Exception:
                pass
            self._halt.wait(0.05)

    def run(self):
        try:
            self._nvml()
        except Exception:
            self._smi()

    def stop(self):
        if self._nv is not None and not self.samples:   # the region was shorter than the thread's start-up: one sample while the GPU is still under load
            try:
                self._sample_nvml()
            except Exception:
                pass
        self._halt.set()
        self.join(timeout=3)
        sm = sorted(s[0] for s in self.samples)
        mx = [s[1] for s in self.samples]
        reasons = sorted({n for s in self.samples for n in s[2]})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.samples), "source": self.source}


def cpu_denoiser(which: str, W: int, H: int, recon: int, prefer_reference: bool):
    """A CPU engine for the workload: the reference's shaders driven by the reference's host library when both were shipped
    (oracle/_ref, `prefer_reference`), else the oracle port driven by the product's host library. Returns (denoiser, step, kind, what, threads)."""
    import torch  # noqa: F401
    from nrd_sample_b200 import nrd_api as api, synth
    from oracle import runner

    threads = usable_cores()
    runner.lib().nrd_oracle_set_threads(threads)
    wl = WORKLOADS[which]
    use_ref = prefer_reference and runner.ref_shaders() is not None
    if use_ref:
        runner.ref_shaders().nrd_refshader_set_threads(threads)
    ref_host = use_ref and os.path.exists(runner.REF_LIB_PATH)
    host = api.NrdLibrary(runner.REF_LIB_PATH) if ref_host else runner.default_host_library()
    den = runner.OracleDenoiser(host, getattr(api.Denoiser, wl["denoiser"]), W, H, engine="reference" if use_ref else "oracle")
    for name, fmt in wl["outputs"]:
        den.set_user_texture(getattr(api.ResourceType, name), runner.alloc_texture(getattr(api.Format, fmt), W, H))
    extra = {"holes": True} if recon else {}
    frames = [getattr(synth, wl["frame"])(i, W, H, period=RING, **extra) for i in range(RING)]
    settings = api.ReblurSettings(hitDistanceReconstructionMode=recon) if recon else None

    def step(i):
        for k, v in frames[i % RING].items():
            den.set_user_texture(getattr(api.ResourceType, k), v)
        den.denoise(synth.common_settings(i, W, H, period=RING), settings=settings if i == 0 else None)

    if use_ref:
        what = ("the reference's compute shaders compiled as C++ (oracle/_ref/libnrd_refshaders.so), thread groups spread over the host threads, dispatch stream from "
                + ("the reference's host library (oracle/_ref/libnrd_ref.so)" if ref_host else "the product's host library (libnrd_ref.so was not shipped)"))
    else:
        what = "oracle/ CPU restatement, OpenMP over rows" + (" (libnrd_refshaders.so was not shipped)" if prefer_reference else "")
    return den, step, ("reference" if use_ref else "port"), what, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    W, H = args.width, args.height
    recon = {"off": 0, "3x3": 1, "5x5": 2}[args.hitdist_reconstruction]
    den, step, kind, what, threads = cpu_denoiser(args.denoiser, W, H, recon, prefer_reference=True)
    wl = WORKLOADS[args.denoiser]
    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.warmup, args.warmup + args.steps):
        step(i)
    dt = time.perf_counter() - t0
    value = W * H * args.steps / dt / 1e6
    line = {
        "impl": "reference", "metric": wl["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{args.steps} steady-state frames of the {W}x{H} stream itself (every step is one whole frame): {what}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference's denoisers are HLSL compute shaders with no CPU or CUDA path; this arm executes those shaders on the CPU through oracle/ref_shim/hlsl_cpu.h",
    }
    print(json.dumps(line))


def cpu_baseline(which="reblur"):
    """The oracle port on this box's host cores, bounded sample (N = 1 only)."""
    W, H, warm, steps = 1280, 720, 4, 8
    den, step, kind, what, threads = cpu_denoiser(which, W, H, 0, prefer_reference=False)
    t0 = 0.0
    for i in range(warm + steps):
        if i == warm:
            t0 = time.perf_counter()
        step(i)
    dt = time.perf_counter() - t0
    return {"value": W * H * steps / dt / 1e6, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"{steps} steady-state frames of a {W}x{H} stream of the same synthetic scene ({what})"}


def profile_numbers(denoiser_key: str, W: int, H: int):
    """DRAM traffic and executed warp-instructions per launch from the committed ncu captures (profiles/dram_traffic.json, profiles/inst_counts.json)."""
    out = {}
    for name in ("dram_traffic", "inst_counts"):
        p = os.path.join(ROOT, "profiles", f"{name}.json")
        try:
            out[name] = json.load(open(p)).get(f"{W}x{H}" if denoiser_key == "reblur" else f"{denoiser_key} {W}x{H}", {})
        except Exception:
            out[name] = {}
    return out


def run_product(args):
    import torch
    import torch.distributed as dist
    from nrd_sample_b200 import executor as ex, nrd_api as api, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the denoiser has no CPU fallback (use --impl reference for the CPU reference arm)")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))

    W, H = args.width, args.height
    px = W * H
    RT = api.ResourceType
    wl = WORKLOADS[args.denoiser]
    pass_bytes = wl["pass_bytes"]

    def fmt_of(name):
        return getattr(api.Format, INPUT_FORMATS.get(name, "RGBA16_SFLOAT"))

    recon = {"off": 0, "3x3": 1, "5x5": 2}[args.hitdist_reconstruction]
    if recon:
        assert args.denoiser == "reblur", "--hitdist-reconstruction applies to the REBLUR workload"
        pass_bytes = dict(pass_bytes, **{"Hit distance reconstruction": 40})
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    class Leg:
        """One denoiser instance + a ring of frames; every rank gets its own stream of frames (seeds offset by the rank)."""

        def __init__(self, recon_mode, flags=None):
            extra = {"holes": True} if recon_mode else {}
            self.frames = [getattr(synth, wl["frame"])(i + 1000 * rank, W, H, device=dev, period=RING, **extra) for i in range(RING)]
            self.outs = [(getattr(RT, name), getattr(api.Format, fmt), ex.alloc_texture(getattr(api.Format, fmt), W, H, dev)) for name, fmt in wl["outputs"]]
            self.den = ex.CudaDenoiser(getattr(api.Denoiser, wl["denoiser"]), W, H, device=local, **({"flags": flags} if flags is not None else {}))
            if recon_mode:
                self.den.set_denoiser_settings(api.ReblurSettings(hitDistanceReconstructionMode=recon_mode))
            self.in_bytes = sum(v.numel() * v.element_size() for v in self.frames[0].values())
            self.out_bytes = sum(t.numel() * t.element_size() for _, _, t in self.outs)

        def settings(self, i):
            return synth.common_settings(i + 1000 * rank, W, H, period=RING)

        def step_device(self, i):
            for k, v in self.frames[i % RING].items():
                self.den.set_user_texture(getattr(RT, k), v, fmt_of(k))
            for rt, fmt, t in self.outs:
                self.den.set_user_texture(rt, t, fmt)
            self.den.set_common_settings(self.settings(i))
            self.den.denoise(stream)

    def timed(den, step_fn, first, profile=False, finish=None):
        for i in range(first, first + args.warmup):
            step_fn(i)
        barrier()
        if profile:
            den.reset_profile()
            den.set_profiling(True)
        launches0 = ex.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(first + args.warmup, first + args.warmup + args.steps):
            step_fn(i)
        if finish:
            finish()   # the timed region ends after the last download has landed
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        launches = ex.launch_count() - launches0
        prof = den.profile() if profile else {}
        den.set_profiling(False)
        return max_over_ranks(ms), launches, prof

    # ---- leg 1: device-resident ( `value`, per-pass roofline ) --------------------------------------------------------------------
    leg = Leg(recon)
    den = leg.den
    try:
        gpu_uuid = str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        gpu_uuid = None
    sampler = ClockSampler(local, gpu_uuid) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_dev, launches, _ = timed(den, leg.step_device, 0)
    clocks = sampler.stop() if sampler else None
    # ... the same steps once more with a CUDA event pair around every dispatch: the per-pass times of `roofline.passes` ( kept out of `value`: the event records
    # sit between the kernels and cost the short chains several percent )
    ms_profiled, _, prof = timed(den, leg.step_device, 0, profile=True)
    # ... and through NRDCU_FLAG_CUDA_GRAPH ( frames replayed from cached CUDA graphs, kernel-node parameters patched per frame )
    gleg = Leg(recon, flags=ex.FLAG_QUAD_INTRINSICS | ex.FLAG_CUDA_GRAPH)
    ms_graph, _, _ = timed(gleg.den, gleg.step_device, 0)
    graph_stats = gleg.den.graph_stats()
    gleg.den.close()
    del gleg

    # ---- leg 2: host buffers ( `e2e` ) ------------------------------------------------------------------------------------------
    host_frames = [{k: v.cpu().pin_memory() for k, v in f.items()} for f in leg.frames]
    host_outs = [torch.zeros_like(t, device="cpu").pin_memory() for _, _, t in leg.outs]

    def step_host(pipelined):
        def fn(i):
            for k, v in host_frames[i % RING].items():
                den.set_host_texture(getattr(RT, k), v, fmt_of(k), is_output=False)
            for (rt, fmt, _), h in zip(leg.outs, host_outs):
                den.set_host_texture(rt, h, fmt, is_output=True)
            den.set_common_settings(leg.settings(i))
            (den.denoise_host_pipelined if pipelined else den.denoise_host)(stream)
        return fn

    ms_host, _, _ = timed(den, step_host(False), args.warmup + args.steps)
    ms_pipe, _, _ = timed(den, step_host(True), 2 * (args.warmup + args.steps), finish=lambda: den.host_flush(stream))
    den.host_flush(stream)
    torch.cuda.synchronize()

    # host frames: the renderer-facing staging blocks ( own instance: its user slots point into the executor's device blocks )
    fden = ex.CudaDenoiser(getattr(api.Denoiser, wl["denoiser"]), W, H, device=local)
    if recon:
        fden.set_denoiser_settings(api.ReblurSettings(hitDistanceReconstructionMode=recon))
    for k in leg.frames[0]:
        fden.declare_host_texture(getattr(RT, k), fmt_of(k), is_output=False)
    for rt, fmt, _ in leg.outs:
        fden.declare_host_texture(rt, fmt, is_output=True)
    in_frames = [fden.host_frame(False) for _ in range(RING)]
    out_frame = fden.host_frame(True)
    for hf, f in zip(in_frames, leg.frames):
        for k, v in f.items():
            hf.tensor(getattr(RT, k)).copy_(v.cpu())   # the "renderer" writes its frame straight into the staging block

    def step_frames(i):
        fden.set_common_settings(leg.settings(i))
        fden.denoise_host_frames(in_frames[i % RING], out_frame, stream)

    ms_frames, _, _ = timed(fden, step_frames, 3 * (args.warmup + args.steps), finish=lambda: fden.host_flush(stream))
    fden.host_flush(stream)
    torch.cuda.synchronize()
    numa_node = in_frames[0].numa_node
    for hf in in_frames + [out_frame]:
        hf.close()
    fden.close()
    del host_frames, host_outs

    # ---- leg 3 ( REBLUR at 1440p ): the README's setting, for vs_baseline ---------------------------------------------------------
    baseline_leg = None
    if args.denoiser == "reblur" and (W, H) == (2560, 1440) and wl["published_ms"]:
        if recon == 1:
            ms_b = ms_dev
        else:
            den.close()
            den = None
            legb = Leg(1)
            ms_b, _, _ = timed(legb.den, legb.step_device, 0)
            legb.den.close()
            del legb
        baseline_leg = {"ms_per_step": ms_b / args.steps, "value": px * args.steps * world / (ms_b * 1e-3) / 1e6, "settings": "hitDistanceReconstructionMode = AREA_3X3, one lobe traced per pixel",
                        "published": f"{wl['published_ms']} ms per 1440p frame on an RTX 4080 (External/NRD/README.md:30)"}
    if den is not None:
        den.close()
    del leg
    torch.cuda.empty_cache()

    # ---- leg 4 ( N > 1, REBLUR ): ONE 3840x2160 frame as N strips ( BASELINE.json config 3 ) -----------------------------------------
    tiled = None
    if world > 1 and args.denoiser == "reblur" and not args.no_tiled:
        try:
            tiled = tiled_4k(args, ex, api, synth, dist, torch, rank, world, local, barrier, max_over_ranks)
        except Exception as e:   # a failed strip run must not take the replica numbers with it
            tiled = {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank == 0:
        peak, peak_src, sm_mhz = measured_peaks()
        total_px = px * args.steps * world
        value = total_px / (ms_dev * 1e-3) / 1e6
        numbers = profile_numbers(args.denoiser, W, H)
        issue_peak = 148 * 4 * (clocks["sm_mhz"] if clocks and clocks.get("sm_mhz") else sm_mhz) * 1e6   # warp-instructions per second: 148 SMs x 4 schedulers x clock
        passes = {}
        # SIGMA's Copy pass rides inside the first blur pass ( kernels/sigma.cu ): its dispatch launches nothing, its bytes are the blur kernel's
        shorts = {name.split(" - ")[-1]: (tot, cnt) for name, (tot, cnt) in prof.items()}
        fused_copy = args.denoiser == "sigma" and "Copy" in shorts and shorts["Copy"][1] and shorts["Copy"][0] / shorts["Copy"][1] < 0.004
        if fused_copy:
            pass_bytes = dict(pass_bytes, **{"Blur": pass_bytes["Blur"] + pass_bytes["Copy"]})
        for name, (tot, cnt) in prof.items():
            short = name.split(" - ")[-1]
            if fused_copy and short == "Copy":
                continue
            if cnt and short in pass_bytes:
                avg_ms = tot / cnt
                gbs = pass_bytes[short] * px / (avg_ms * 1e-3) / 1e9
                passes[short] = {"avg_us": round(avg_ms * 1e3, 2), "launches_per_step": cnt // args.steps, "alg_bytes_per_px": pass_bytes[short], "achieved_gbs": round(gbs, 1),
                                 "frac": round(gbs / peak, 4)}
                inst = numbers["inst_counts"].get(short)
                if inst:   # executed warp-instructions per launch ( ncu smsp__inst_executed.sum ) over what the schedulers could issue in the measured time
                    passes[short]["issue_frac"] = round(inst / (avg_ms * 1e-3 * issue_peak), 4)
        dom = max(passes, key=lambda k: passes[k]["avg_us"] * passes[k]["launches_per_step"]) if passes else None
        chain_bytes = sum(pass_bytes[k] * v["launches_per_step"] for k, v in passes.items())
        chain_gbs = chain_bytes * px / (ms_dev / args.steps * 1e-3) / 1e9
        chain_inst = sum(numbers["inst_counts"].get(k, 0) * v["launches_per_step"] for k, v in passes.items())
        roofline = None
        if dom:
            roofline = {"bound": "hbm", "kernel": dom, "achieved": passes[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": passes[dom]["frac"],
                        "traffic": numbers["dram_traffic"].get(dom), "peak_source": peak_src, "alg_bytes_per_launch": pass_bytes[dom] * px,
                        "issue_frac": passes[dom].get("issue_frac"),
                        "chain": {"alg_bytes_per_px": chain_bytes, "achieved": round(chain_gbs, 1), "frac": round(chain_gbs / peak, 4),
                                  "issue_frac": round(chain_inst / (ms_dev / args.steps * 1e-3 * issue_peak), 4) if chain_inst else None},
                        "passes": passes,
                        "note": "the chains are FP32-issue bound on B200 (30-40 FLOP per algorithmic byte vs a ~10 FLOP/B fp32 ridge): issue_frac = executed warp-instructions (ncu, profiles/inst_counts.json) / (time x 148 SMs x 4 schedulers x SM clock); see DESIGN.md"}
        cfg = config_of(args, world)
        cfg["baseline_note"] = ("vs_baseline = per-GPU Mpixels/s of baseline_leg (AREA_3X3, like the published run) / (3.6864 Mpixels / 2.55 ms on an RTX 4080, NRD README)" if baseline_leg
                                else (f"vs_baseline = per-GPU value / ({wl['published_ms']} ms per 1440p frame on an RTX 4080, NRD README)" if wl["published_ms"] and (W, H) == (2560, 1440)
                                      else "no published number for this denoiser / resolution"))
        if baseline_leg:
            vs_baseline = baseline_leg["value"] / world / (3.6864 / (wl["published_ms"] * 1e-3))
        elif wl["published_ms"] and (W, H) == (2560, 1440):
            vs_baseline = value / world / (3.6864 / (wl["published_ms"] * 1e-3))
        else:
            vs_baseline = None

        def mpx(ms):
            return total_px / (ms * 1e-3) / 1e6

        e2e_ms = ms_frames / args.steps
        line = {
            "metric": wl["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": vs_baseline, "dtype": "f32", "data": "synthetic", "config": cfg,
            "e2e": {"value": mpx(ms_frames), "unit": UNIT, "h2d_bytes_per_step": leg_in_bytes(W, H, wl, recon, api), "d2h_bytes_per_step": leg_out_bytes(W, H, wl, api), "ms_per_step": e2e_ms,
                    "call": "nrdcuDenoiseHostFrames: renderer-facing staging blocks in pinned NUMA-local host memory -> ONE H2D copy -> chain -> ONE D2H copy every step, copies of neighbouring steps overlap the kernels",
                    "pcie_gbs_per_gpu": {"h2d": round(leg_in_bytes(W, H, wl, recon, api) / (e2e_ms * 1e-3) / 1e9, 1), "d2h": round(leg_out_bytes(W, H, wl, api) / (e2e_ms * 1e-3) / 1e9, 1)},
                    "host_numa_node_rank0": numa_node,
                    "per_texture": {"value": mpx(ms_pipe), "ms_per_step": ms_pipe / args.steps, "call": "nrdcuDenoiseHostPipelined: one 2D copy per texture from caller-owned pinned buffers, pipelined"},
                    "serial": {"value": mpx(ms_host), "ms_per_step": ms_host / args.steps, "call": "nrdcuDenoiseHost: the same copies and kernels back to back on one stream"}},
            "gpu_launches": launches, "roofline": roofline, "clocks": clocks,
            "cuda_graph": {"ms_per_step": ms_graph / args.steps, "value": mpx(ms_graph), "flag": "NRDCU_FLAG_CUDA_GRAPH", **graph_stats},
            "ms_per_step_with_per_pass_events": ms_profiled / args.steps,
        }
        if baseline_leg:
            line["baseline_leg"] = baseline_leg
        if tiled is not None:
            line["tiled_4k"] = tiled
        # CPU baseline: the oracle port on this box's host cores, bounded sample (N=1 only)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.denoiser)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def leg_in_bytes(W, H, wl, recon, api):
    from nrd_sample_b200 import synth
    names = list(getattr(synth, wl["frame"])(0, 32, 32).keys())
    return sum(W * H * api.FORMAT_BYTES[getattr(api.Format, INPUT_FORMATS.get(k, "RGBA16_SFLOAT"))] for k in names)


def leg_out_bytes(W, H, wl, api):
    return sum(W * H * api.FORMAT_BYTES[getattr(api.Format, fmt)] for _, fmt in wl["outputs"])


def tiled_4k(args, ex, api, synth, dist, torch, rank, world, local, barrier, max_over_ranks):
    """BASELINE.json config 3: ONE 3840x2160 REBLUR_DIFFUSE_SPECULAR frame per step, rank r computing strip r of every pass (cuts balanced by
    denoising-range pixels), seam rows pushed into the neighbours' HBM over NVLink peer memory after each pass (nrd_sample_b200/tiling.py,
    csrc/kernels/peer_halo.cu). Strong scaling: `speedup_vs_1` = whole-frame time on one GPU (rank 0 measures it in the same run) / strip time."""
    from nrd_sample_b200 import tiling
    W, H = 3840, 2160
    dev = f"cuda:{local}"
    RT, F16 = api.ResourceType, api.Format.RGBA16_SFLOAT
    fmt = {"IN_VIEWZ": api.Format.R32_SFLOAT, "IN_NORMAL_ROUGHNESS": api.Format.R10_G10_B10_A2_UNORM}
    frames = [synth.reblur_frame(i, W, H, device=dev, period=RING) for i in range(RING)]
    stream = torch.cuda.current_stream()
    steps, warm = args.steps, max(args.warmup, 6)

    def run(den):
        def step(i):
            for k, v in frames[i % RING].items():
                den.set_user_texture(getattr(RT, k), v, fmt.get(k, F16))   # IN_MV is bound read-write by temporal stabilisation but only written on clear frames
            den.set_common_settings(synth.common_settings(i, W, H, period=RING))
            den.denoise()
        for i in range(warm):
            step(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(warm, warm + steps):
            step(i)
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1) / steps

    # one GPU, whole frame ( every rank runs it so that the barriers inside run() stay matched; rank 0's time is the reference )
    whole = ex.CudaDenoiser(api.Denoiser.REBLUR_DIFFUSE_SPECULAR, W, H, device=local)
    keep = [ex.alloc_texture(F16, W, H, dev), ex.alloc_texture(F16, W, H, dev)]
    whole.set_user_texture(RT.OUT_DIFF_RADIANCE_HITDIST, keep[0], F16)
    whole.set_user_texture(RT.OUT_SPEC_RADIANCE_HITDIST, keep[1], F16)
    ms_one = run(whole)
    whole.close()
    t = torch.tensor([ms_one], device=dev)
    dist.broadcast(t, 0)
    ms_one = float(t.item())

    weights = tiling.tile_row_weights(frames[0]["IN_VIEWZ"])   # the same on every rank: all hold the full input frame

    def strips(flags):
        den = tiling.TiledDenoiser(api.Denoiser.REBLUR_DIFFUSE_SPECULAR, W, H, rank, world, local, mode="peer", row_weights=weights, flags=flags)
        outs = [den.shared_texture(RT.OUT_DIFF_RADIANCE_HITDIST, F16), den.shared_texture(RT.OUT_SPEC_RADIANCE_HITDIST, F16)]   # in the context, so the neighbours can map them
        den.attach_peers()
        sent0 = den.status()[0]
        ms = max_over_ranks(run(den))
        sent, wait_error = den.status()
        info = (den.strips, den.halo, (sent - sent0) // (steps + warm), wait_error)
        den.close()
        del outs
        return ms, info

    # A/B: every pass as one launch followed by the push of its seam rows, then ( the default ) seam rows first with their push overlapping the interior
    ms_serial, _ = strips(ex.FLAG_QUAD_INTRINSICS | ex.FLAG_NO_SEAM_OVERLAP)
    ms, (strip_rows, halo, per_frame, wait_error) = strips(ex.FLAG_QUAD_INTRINSICS)
    tb = torch.tensor([float(per_frame)], device=dev)
    dist.all_reduce(tb)
    out = {"ms_per_frame": ms, "ms_per_frame_1gpu": ms_one, "speedup_vs_1": ms_one / ms, "mpixels_per_s": W * H / ms / 1e3, "halo_bytes": int(tb.item()),
           "halo_bytes_note": "seam bytes pushed per frame, summed over all strips",
           "mode": "peer stores over NVLink (CUDA IPC), work-balanced cuts, seam rows of every pass computed first so that their push overlaps the interior",
           "ms_per_frame_without_seam_overlap": ms_serial, "strips": strip_rows, "halo_rows": halo, "max_vertical_motion_rows": halo - 2, "wait_error": wait_error,
           "resolution": [W, H], "steps": steps}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--width", type=int, default=2560)
    ap.add_argument("--height", type=int, default=1440)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tiled", action="store_true", help="N > 1: skip the 3840x2160 strip run (tiled_4k)")
    ap.add_argument("--hitdist-reconstruction", default="off", choices=["off", "3x3", "5x5"],
                    help="REBLUR only: run with ReblurSettings::hitDistanceReconstructionMode (the README's 2.55 ms was taken with 3x3) on inputs with one lobe per pixel")
    ap.add_argument("--denoiser", default="reblur", choices=sorted(WORKLOADS), help="reblur = the headline workload (default); relax / sigma = BASELINE.json configs 2 and 0")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
