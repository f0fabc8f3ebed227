"""What one rank of an N-GPU tile split costs, measured on ONE GPU: the cfg-4 frame's slice of rank r of `world` (4-row tiles dealt
round-robin) is rendered alone and timed with CUDA events; slice_ms * world / full_ms - 1 is the scaling loss that comes from the
frame kernels themselves (fixed costs, tails), as opposed to the exchange.  PROBE_SLICE_ONE=world,rank renders only that slice, once
warmed up — for an ncu launch list; PROBE_CASES="1,0;8,3" picks the (world, rank) cases; PROBE_OPTS="key=value,..." sets context
options first.  One JSON line per measurement."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpuvoxelraycaster_b200 as vrt  # noqa: E402
from cpuvoxelraycaster_b200.frame import FrameRenderer  # noqa: E402


def time_frame(fr, cs, p, stream, reps=5):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    with torch.cuda.stream(stream):
        for _ in range(2):
            fr.accum.zero_()
            fr.accumulate(cs, p)
        for i in range(reps):
            fr.accum.zero_()
            ev[i].record(stream)
            fr.accumulate(cs, p)
        ev[reps].record(stream)
    stream.synchronize()
    return float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]))


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
    for kv in filter(None, os.environ.get("PROBE_OPTS", "").split(",")):          # context options, e.g. PROBE_OPTS=beam_overlap=0
        k, v = kv.split("=")
        try:
            ctx.set_option(k, int(v))
        except Exception as e:                                                       # an older library build: say so, carry on
            print(json.dumps(dict(option=k, error=str(e)[:80])), flush=True)
    one = os.environ.get("PROBE_SLICE_ONE")
    cases = [tuple(int(x) for x in one.split(","))] if one else [tuple(int(x) for x in c.split(",")) for c in os.environ["PROBE_CASES"].split(";")] if os.environ.get("PROBE_CASES") else [(1, 0), (2, 0), (2, 1), (4, 0), (4, 3), (8, 0), (8, 3), (8, 7)]
    full = None
    for world, rank in cases:
        fr = FrameRenderer(scene, 1920, 1080, rank, world, None, None, stream, exchange="nccl")
        fr.light = np.float32([-200, -1000, -300]) * np.float32(1.0 / S) + np.float32(1.0)
        fr.use_gi, fr.gi_bounces = True, 2
        p = fr.params(int(os.environ.get("PROBE_SPP", "64")))
        ctx.set_option("time_frame_kernels", 1)
        ms = time_frame(fr, cs, p, stream, reps=1 if one else 5)
        kt = ctx.take_kernel_timings()
        ctx.set_option("time_frame_kernels", 0)
        if world == 1:
            full = ms
        st = fr.stats()
        print(json.dumps(dict(world=world, rank=rank, ms=round(ms, 3), sort_and_trace_ms=round(float(np.median(kt)), 3) if len(kt) else None,
                              rays=sum(st["rays"]), trips=sum(st["complexity"]), loss_pct=round(100 * (ms * world / full - 1), 2) if full else None)), flush=True)


if __name__ == "__main__":
    main()
