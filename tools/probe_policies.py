"""A/B of the traversal-loop variants on one GPU (CUDA events, medians):
  * K6 frame kernel, vrt_context_set_option "trav_policy": 0 = Trav (round 1 loop), 1 = Trav2, 2 = Trav2 with the cone test
    compiled out of the coef-0 casts (default) — cfg-4 frame without GI / 1 bounce / 2 bounces, and the 1/8 slice;
  * K1 batched cast, "cast_variant" 0 = Trav, 2 = Trav2 (K1b), 1 = K1p — 1080p primary rays and random rays at 2048^3.
Every variant must give the same bytes (checked by hash).  VRT_LIBRARY=tools/libvrt_r01.so runs the round-1 build of the
library through the same script (it has no variants: one line per workload).  One JSON line per measurement."""
import hashlib
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpuvoxelraycaster_b200 as vrt  # noqa: E402
from cpuvoxelraycaster_b200.frame import FrameRenderer  # noqa: E402

TAG = os.environ.get("PROBE_TAG", "current")
BEAMS = [int(x) for x in os.environ.get("PROBE_BEAMS", "0,4,8,16,32").split(",")]


def time_frame(fr, cs, p, stream, reps=3):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    with torch.cuda.stream(stream):
        fr.accum.zero_()
        fr.accumulate(cs, p)
        for i in range(reps):
            fr.accum.zero_()
            ev[i].record(stream)
            fr.accumulate(cs, p)
        ev[reps].record(stream)
    stream.synchronize()
    return float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]))


def set_opt(ctx, key, value):
    try:
        ctx.set_option(key, value)
        return True
    except Exception:
        return False


def main():
    D = int(os.environ.get("PROBE_DEPTH", "11"))
    S = float(1 << D)
    stream = torch.cuda.Stream()
    ctx = vrt.Context(0, stream.cuda_stream)
    scene = vrt.LSVO.from_terrain(ctx, D)
    t = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "textures.npz"))
    scene.set_textures(t["top"], t["side"])
    cam = vrt.Camera(position=(S / 2, S / 2 - 56, S / 2), view_angle=(0, 0), aperture=0.5)
    cam.autofocus(scene)
    cs = cam.as_struct()
    policies = [0, 1, 2] if set_opt(ctx, "trav_policy", 0) else [None]
    for world in (1, 8):
        fr = FrameRenderer(scene, 1920, 1080, 0, world, None, None, stream, exchange="nccl")
        fr.light = np.float32([-200, -1000, -300]) * np.float32(1.0 / S) + np.float32(1.0)
        for (gi, bounces) in ((False, 1), (True, 1), (True, 2)):
            if world == 8 and not (gi and bounces == 2):
                continue
            fr.use_gi, fr.gi_bounces = gi, bounces
            p = fr.params(64)
            for pol in policies:
                if pol is not None:
                    ctx.set_option("trav_policy", pol)
                for beam in (BEAMS if pol in (2, None) else [0]):
                    have_beam = set_opt(ctx, "beam_tile", 0 if beam < 0 else beam)
                    set_opt(ctx, "bounds_exit", 0 if beam <= 0 and beam != -1 else 1)      # -1 = bounds only; > 0 = beam + bounds; 0 = neither
                    if beam and not have_beam:
                        continue
                    ms = time_frame(fr, cs, p, stream)
                    st = fr.stats()
                    h = hashlib.sha256(fr.accum.cpu().numpy().tobytes()).hexdigest()[:16]
                    print(json.dumps(dict(tag=TAG, what="K6 frame", world=world, gi=gi, bounces=bounces, policy=pol, beam_tile=beam, ms=round(ms, 3),
                                          rays=sum(st["rays"]), trips=sum(st["complexity"]), primary_trips=st["complexity"][0], hash=h)), flush=True)
            set_opt(ctx, "beam_tile", 8)
            set_opt(ctx, "bounds_exit", 1)
    if policies[0] is not None:
        ctx.set_option("trav_policy", 2)
    # batched casts
    W, H = 1920, 1080
    x, y = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    d = np.stack([x / np.float32(H) - np.float32(W / H * 0.5), y / np.float32(H) - np.float32(0.5), np.ones_like(x)], -1).reshape(-1, 3)
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    o = np.broadcast_to(np.float32([S / 2, S / 2 - 56, S / 2]) / np.float32(S) + np.float32(1), d.shape).astype(np.float32).copy()
    rng = np.random.default_rng(7)
    n = 1 << 22
    ro = rng.uniform([1, 1, 1], [2, 1.45, 2], (n, 3)).astype(np.float32)
    rd = rng.normal(size=(n, 3)).astype(np.float32)
    rd /= np.linalg.norm(rd, axis=1, keepdims=True)
    for name, (oo, dd) in (("primary 1080p", (o, d)), ("random 4M", (ro, rd))):
        do, ddv = torch.from_numpy(oo).cuda(), torch.from_numpy(dd).cuda()
        for coef in (0.0, 0.5):
            out = torch.empty(len(oo) * 16, dtype=torch.int32, device="cuda")
            for variant in (0, 2, 1, 3):
                if not set_opt(ctx, "cast_variant", variant):
                    continue
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
                with torch.cuda.stream(stream):
                    scene.cast_rays_device(do, ddv, len(oo), out, coef, 0.0)
                    for i in range(5):
                        ev[i].record(stream)
                        scene.cast_rays_device(do, ddv, len(oo), out, coef, 0.0)
                    ev[5].record(stream)
                stream.synchronize()
                ms = float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(5)]))
                h = hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:16]
                print(json.dumps(dict(tag=TAG, what="cast " + name, coef=coef, cast_variant=variant, ms=round(ms, 4),
                                      grays_s=round(len(oo) / ms / 1e6, 2), trips=scene.last_complexity(), hash=h)), flush=True)
    set_opt(ctx, "cast_variant", 3)


if __name__ == "__main__":
    main()
