#!/usr/bin/env python
"""Headline benchmark: denoised Mpixels/s of the REBLUR_DIFFUSE_SPECULAR pass chain at 2560x1440 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--width 2560 --height 1440]

A "step" is one nrdcuDenoise of one synthetic frame (7 passes: classify tiles, pre-pass, temporal accumulation,
history fix, blur, post-blur, temporal stabilisation) in steady state. `value` times K steps with the frame's inputs
already in HBM (CUDA events on the launch stream, barrier + synchronize on both sides, max over ranks); `e2e` times the
same K steps through the host-buffer entry point nrdcuDenoiseHostPipelined: pinned host inputs -> H2D -> 7 kernels -> D2H of both
denoised outputs, every step, with the uploads / downloads of neighbouring steps overlapping the kernels (`e2e.serial` = the same
through nrdcuDenoiseHost, everything back to back on one stream). N > 1 (torchrun): every rank denoises its own independent stream of frames (replicas, no
data-path collective — REBLUR frames shard as independent streams, BASELINE.json config 5), `value` is the aggregate.

`--impl reference`: the reference has no CPU (or CUDA) implementation of this path — its HLSL cannot be built or run
here — so this arm times the oracle's CPU restatement of the same chain (oracle/, "port") on all host cores, on a
bounded sample of the workload (a 1280x720 stream of the same synthetic scene), rank 0 only.
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
CHAIN_BYTES_PER_PIXEL = sum(PASS_BYTES_PER_PIXEL.values())  # 350
PUBLISHED_MPX_S = 3.6864 / 2.55e-3  # NRD/README.md:30: REBLUR_DIFFUSE_SPECULAR 2.55 ms @1440p on RTX 4080
RING = 4  # distinct frames cycled through (4 x 118 MB of inputs at 1440p >> 126 MB L2)

# Secondary workloads (`--denoiser relax|sigma`): BASELINE.json configs 2 and 0/1', same harness, own metric name. Compulsory bytes per pixel per
# pass from SURVEY.md App. B; published RTX 4080 times from External/NRD/README.md:32,34.
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


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons, sampled every 200 ms while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self._halt = index, [], threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        sm = sorted(int(float(s[0])) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = [int(float(s[1])) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.samples)}


def run_reference(args):
    """Reference arm: the reference's OWN compute shaders (External/NRD/Shaders/*.cs.hlsl compiled as C++ into oracle/_ref/libnrd_refshaders.so
    in the build container) executing the chain on the host cores; falls back to the oracle port when that library was not
    shipped. The reference has no CPU path of its own: this is the closest thing to "the reference on this box's CPU"."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch  # noqa: F401
    from nrd_sample_b200 import nrd_api as api, synth
    from oracle import runner

    W, H = 960, 540  # bounded sample: a smaller stream of the same scene (cost per pixel is resolution independent)
    threads = usable_cores()
    runner.lib().nrd_oracle_set_threads(threads)
    engine = "reference" if runner.ref_shaders() is not None else "oracle"
    if engine == "reference":
        runner.ref_shaders().nrd_refshader_set_threads(threads)
    wl = WORKLOADS[args.denoiser]
    den = runner.OracleDenoiser(runner.default_host_library(), getattr(api.Denoiser, wl["denoiser"]), W, H, engine=engine)
    for name, fmt in wl["outputs"]:
        den.set_user_texture(getattr(api.ResourceType, name), runner.alloc_texture(getattr(api.Format, fmt), W, H))
    frames = [getattr(synth, wl["frame"])(i, W, H, period=RING) for i in range(RING)]

    def step(i):
        for k, v in frames[i % RING].items():
            den.set_user_texture(getattr(api.ResourceType, k), v)
        den.denoise(synth.common_settings(i, W, H, period=RING))

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.warmup, args.warmup + args.steps):
        step(i)
    dt = time.perf_counter() - t0
    value = W * H * args.steps / dt / 1e6
    kind = "reference" if engine == "reference" else "port"
    what = ("the reference's compute shaders compiled as C++ (oracle/_ref/libnrd_refshaders.so), thread groups spread over the host threads" if engine == "reference"
            else "oracle/ CPU restatement (libnrd_refshaders.so was not shipped)")
    line = {
        "impl": "reference", "metric": wl["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["what"].format(w=args.width, h=args.height), "denoiser": wl["denoiser"], "resolution": [args.width, args.height], "settings": "library defaults"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{args.steps} steady-state frames of a {W}x{H} stream of the same synthetic scene: {what}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference's denoisers are HLSL compute shaders with no CPU or CUDA path; this arm executes those shaders on the CPU through oracle/ref_shim/hlsl_cpu.h",
    }
    print(json.dumps(line))


def run_product(args):
    import torch
    import torch.distributed as dist
    from nrd_sample_b200 import executor as ex, nrd_api as api, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the denoiser has no CPU fallback (use --impl reference for the CPU oracle arm)")
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

    # every rank gets its own stream of frames (different seeds per rank via the frame index offset)
    recon = {"off": 0, "3x3": 1, "5x5": 2}[args.hitdist_reconstruction]
    extra = {"holes": True} if recon else {}
    if recon:
        assert args.denoiser == "reblur", "--hitdist-reconstruction applies to the REBLUR workload"
        pass_bytes = dict(pass_bytes, **{"Hit distance reconstruction": 40})
    frames = [getattr(synth, wl["frame"])(i + 1000 * rank, W, H, device=dev, period=RING, **extra) for i in range(RING)]
    host_frames = [{k: v.cpu().pin_memory() for k, v in f.items()} for f in frames]
    outs = [(getattr(RT, name), getattr(api.Format, fmt), ex.alloc_texture(getattr(api.Format, fmt), W, H, dev)) for name, fmt in wl["outputs"]]
    host_outs = [torch.zeros_like(t, device="cpu").pin_memory() for _, _, t in outs]
    in_bytes = sum(v.numel() * v.element_size() for v in frames[0].values())
    out_bytes = sum(t.numel() * t.element_size() for _, _, t in outs)

    den = ex.CudaDenoiser(getattr(api.Denoiser, wl["denoiser"]), W, H, device=local)
    if recon:
        den.set_denoiser_settings(api.ReblurSettings(hitDistanceReconstructionMode=recon))
    stream = torch.cuda.current_stream()

    def settings(i):
        return synth.common_settings(i + 1000 * rank, W, H, period=RING)

    def step_device(i):
        for k, v in frames[i % RING].items():
            den.set_user_texture(getattr(RT, k), v, fmt_of(k))
        for rt, fmt, t in outs:
            den.set_user_texture(rt, t, fmt)
        den.set_common_settings(settings(i))
        den.denoise(stream)

    def step_host(i):
        for k, v in host_frames[i % RING].items():
            den.set_host_texture(getattr(RT, k), v, fmt_of(k), is_output=False)
        for (rt, fmt, _), h in zip(outs, host_outs):
            den.set_host_texture(rt, h, fmt, is_output=True)
        den.set_common_settings(settings(i))
        den.denoise_host(stream)

    def step_host_pipelined(i):
        for k, v in host_frames[i % RING].items():
            den.set_host_texture(getattr(RT, k), v, fmt_of(k), is_output=False)
        for (rt, fmt, _), h in zip(outs, host_outs):
            den.set_host_texture(rt, h, fmt, is_output=True)
        den.set_common_settings(settings(i))
        den.denoise_host_pipelined(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, first, profile=False, finish=None):
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
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, prof

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_dev, launches, prof = timed(step_device, 0, profile=True)
    clocks = sampler.stop() if sampler else None
    ms_host, _, _ = timed(step_host, args.warmup + args.steps)
    ms_pipe, _, _ = timed(step_host_pipelined, 2 * (args.warmup + args.steps), finish=lambda: den.host_flush(stream))

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        total_px = px * args.steps * world
        value = total_px / (ms_dev * 1e-3) / 1e6
        e2e_value = total_px / (ms_pipe * 1e-3) / 1e6
        e2e_serial = total_px / (ms_host * 1e-3) / 1e6
        passes = {}
        for name, (tot, cnt) in prof.items():
            short = name.split(" - ")[-1]
            if cnt and short in pass_bytes:
                avg_ms = tot / cnt
                gbs = pass_bytes[short] * px / (avg_ms * 1e-3) / 1e9
                passes[short] = {"avg_us": round(avg_ms * 1e3, 2), "launches_per_step": cnt // args.steps, "alg_bytes_per_px": pass_bytes[short], "achieved_gbs": round(gbs, 1),
                                 "frac": round(gbs / peak, 4)}
        dom = max(passes, key=lambda k: passes[k]["avg_us"]) if passes else None
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
        if dom and os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(f"{W}x{H}" if args.denoiser == "reblur" else f"{args.denoiser} {W}x{H}", {}).get(dom)
            except Exception:
                traffic = None
        chain_bytes = sum(pass_bytes[k] * v["launches_per_step"] for k, v in passes.items())
        chain_gbs = chain_bytes * px / (ms_dev / args.steps * 1e-3) / 1e9
        roofline = None
        if dom:
            roofline = {"bound": "hbm", "kernel": dom, "achieved": passes[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": passes[dom]["frac"], "traffic": traffic,
                        "peak_source": peak_src, "alg_bytes_per_launch": pass_bytes[dom] * px,
                        "chain": {"alg_bytes_per_px": chain_bytes, "achieved": round(chain_gbs, 1), "frac": round(chain_gbs / peak, 4)}, "passes": passes,
                        "note": "the chains are FP32-issue bound on B200 (30-40 FLOP per algorithmic byte vs a ~10 FLOP/B fp32 ridge); see DESIGN.md"}
        line = {
            "metric": wl["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": value / world / (3.6864 / (wl["published_ms"] * 1e-3)) if (W, H) == (2560, 1440) and wl["published_ms"] else None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": wl["what"].format(w=W, h=H), "denoiser": wl["denoiser"],
                       "resolution": [W, H],
                       "settings": "library defaults" if not recon else f"library defaults + hitDistanceReconstructionMode = AREA_{args.hitdist_reconstruction.upper()} (the NRD README's setting), one lobe traced per pixel",
                       "streams_per_gpu": 1, "parallelism": f"replicas x{world} (independent frame streams, no collective)",
                       "l2_policy": f"ring of {RING} distinct frames: {RING * in_bytes // 2**20} MiB of inputs + pools > 126 MB L2",
                       "baseline_note": (f"vs_baseline = per-GPU value / ({wl['published_ms']} ms per 1440p frame on an RTX 4080, NRD README)" if wl["published_ms"]
                                         else "no published number for this denoiser")},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes, "ms_per_step": ms_pipe / args.steps,
                    "call": "nrdcuDenoiseHostPipelined: pinned host inputs -> H2D -> chain -> D2H of the outputs every step, uploads / downloads of neighbouring steps overlap the kernels",
                    "serial": {"value": e2e_serial, "ms_per_step": ms_host / args.steps, "call": "nrdcuDenoiseHost: the same copies and kernels back to back on one stream"}},
            "gpu_launches": launches, "roofline": roofline, "clocks": clocks,
        }
        # CPU baseline: the oracle port on this box's host cores, bounded sample (N=1 only)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.denoiser)
        print(json.dumps(line))
    den.close()
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(which="reblur"):
    import torch  # noqa: F401
    from nrd_sample_b200 import nrd_api as api, synth
    from oracle import runner

    W, H, warm, steps = 1280, 720, 4, 8
    threads = usable_cores()
    runner.lib().nrd_oracle_set_threads(threads)
    wl = WORKLOADS[which]
    den = runner.OracleDenoiser(runner.default_host_library(), getattr(api.Denoiser, wl["denoiser"]), W, H)
    for name, fmt in wl["outputs"]:
        den.set_user_texture(getattr(api.ResourceType, name), runner.alloc_texture(getattr(api.Format, fmt), W, H))
    frames = [getattr(synth, wl["frame"])(i, W, H, period=RING) for i in range(RING)]
    t0 = 0.0
    for i in range(warm + steps):
        if i == warm:
            t0 = time.perf_counter()
        for k, v in frames[i % RING].items():
            den.set_user_texture(getattr(api.ResourceType, k), v)
        den.denoise(synth.common_settings(i, W, H, period=RING))
    dt = time.perf_counter() - t0
    return {"value": W * H * steps / dt / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{steps} steady-state frames of a {W}x{H} stream of the same synthetic scene (oracle/ CPU restatement, OpenMP over rows)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--width", type=int, default=2560)
    ap.add_argument("--height", type=int, default=1440)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
