"""Multi-GPU strip tiling on real GPUs (BASELINE.json config 3). Run under torchrun, one rank per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/tiled_check.py [W H [FRAMES [STEPS]]]

1. parity: FRAMES frames of REBLUR_DIFFUSE_SPECULAR at WxH denoised as N strips with NCCL halo exchange must equal, row for row
   and bit for bit, the same frames denoised whole on one GPU (every rank computes the whole-frame run itself);
2. timing: STEPS steady-state frames, CUDA events on the launch stream, barrier on both sides, max over ranks."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nrd_sample_b200 import executor as ex, nrd_api as api, synth, tiling  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 3840
H = int(sys.argv[2]) if len(sys.argv) > 2 else 2160
FRAMES = int(sys.argv[3]) if len(sys.argv) > 3 else 6
STEPS = int(sys.argv[4]) if len(sys.argv) > 4 else 20
MODE = sys.argv[5] if len(sys.argv) > 5 else "peer"   # "peer": seam rows pushed over NVLink peer memory by the executor; "nccl": torch.distributed send/recv
SPLIT = sys.argv[6] if len(sys.argv) > 6 else "balanced"   # "balanced": cuts by denoising-range pixels per tile row; "even": equal row counts
WHAT = sys.argv[7] if len(sys.argv) > 7 else "reblur"      # reblur | relax ( RELAX_DIFFUSE_SPECULAR_SH ) | sigma ( SIGMA_SHADOW ): every family takes row ranges
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = f"cuda:{local}"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))
RT, F16 = api.ResourceType, api.Format.RGBA16_SFLOAT
FMT = {"IN_VIEWZ": api.Format.R32_SFLOAT, "IN_NORMAL_ROUGHNESS": api.Format.R10_G10_B10_A2_UNORM, "IN_PENUMBRA": api.Format.R16_SFLOAT}
RING = 4
DENOISER, FRAME_FN, OUTPUTS = {
    "reblur": (api.Denoiser.REBLUR_DIFFUSE_SPECULAR, synth.reblur_frame, [(RT.OUT_DIFF_RADIANCE_HITDIST, F16), (RT.OUT_SPEC_RADIANCE_HITDIST, F16)]),
    "relax": (api.Denoiser.RELAX_DIFFUSE_SPECULAR_SH, synth.relax_frame, [(getattr(RT, n), F16) for n in ("OUT_DIFF_SH0", "OUT_DIFF_SH1", "OUT_SPEC_SH0", "OUT_SPEC_SH1")]),
    "sigma": (api.Denoiser.SIGMA_SHADOW, synth.sigma_frame, [(RT.OUT_SHADOW_TRANSLUCENCY, api.Format.R8_UNORM)]),
}[WHAT]
frames = [FRAME_FN(i, W, H, device=dev, period=RING) for i in range(RING)]


def make(cls, *a, **kw):
    den = cls(*a, **kw)
    if WHAT == "sigma":
        import ctypes as C
        den.set_denoiser_settings(api.SigmaSettings(lightDirection=(C.c_float * 3)(0.0, 0.0, 1.0)))
    if kw.get("mode") == "peer" and world > 1:   # outputs live in the context so that the neighbouring strips can map them
        outs = [den.shared_texture(rt, fmt) for rt, fmt in OUTPUTS]
        den.attach_peers()
        return den, outs
    outs = [ex.alloc_texture(fmt, W, H, dev) for _, fmt in OUTPUTS]
    for (rt, fmt), t in zip(OUTPUTS, outs):
        den.set_user_texture(rt, t, fmt)
    return den, outs


def step(den, i):
    for k, v in frames[i % RING].items():
        den.set_user_texture(getattr(RT, k), v.clone() if k == "IN_MV" else v, FMT.get(k, F16))   # IN_MV is bound read-write by TS
    den.set_common_settings(synth.common_settings(i, W, H, period=RING))
    den.denoise()


weights = tiling.tile_row_weights(frames[0]["IN_VIEWZ"]) if SPLIT == "balanced" else None   # same on every rank: all hold the full input frame
tiled, t_out = make(tiling.TiledDenoiser, DENOISER, W, H, rank, world, local, mode=MODE, row_weights=weights)
whole, w_out = make(ex.CudaDenoiser, DENOISER, W, H, 0, local)
y0, y1 = tiled.rows
ok = True
for i in range(FRAMES):
    step(tiled, i)
    step(whole, i)
    torch.cuda.synchronize()
    for a, b in zip(t_out, w_out):
        same = torch.equal(a[y0:y1].contiguous().view(torch.uint8), b[y0:y1].contiguous().view(torch.uint8))
        if not same:
            d = (a[y0:y1].float() - b[y0:y1].float()).abs()
            print(f"rank {rank} frame {i}: strip rows [{y0},{y1}) differ from the whole-frame run: {int((d > 0).any(-1).sum())} px, max {d.max().item():.4g}", flush=True)
        ok &= same
flag = torch.tensor([1 if ok else 0], device=dev)
if world > 1:
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
whole.close()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


stream = torch.cuda.current_stream()
for i in range(FRAMES, FRAMES + 6):
    step(tiled, i)
barrier()
sent0 = tiled.bytes_sent + tiled.status()[0]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for i in range(FRAMES + 6, FRAMES + 6 + STEPS):
    step(tiled, i)
e1.record(stream)
barrier()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    per = float(ms.item()) / STEPS
    print(json.dumps({"check": "tiled strips == whole frame (bit exact)", "denoiser": WHAT, "passed": bool(flag.item()), "n_gpus": world, "resolution": [W, H], "strips": tiled.strips,
                      "motion_bound_rows_and_worst_overshoot": list(tiled.motion_bound()),
                      "halo_rows": tiled.halo, "ms_per_frame": per, "mpixels_per_s": W * H / per / 1e3,
                      "mode": tiled.mode, "split": SPLIT, "halo_bytes_sent_per_frame_rank0": (tiled.bytes_sent + tiled.status()[0] - sent0) // STEPS, "wait_error": tiled.status()[1]}))
tiled.close()
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
