"""A/B of the traversal-loop policies of K6 (vrt_context_set_option "trav_policy") on the cfg-4 frame:
-1 = plain per-lane loop, 0 = warp-synchronous loop executing both paths, 1 = descend priority, 2 = majority path,
3 / 4 = both paths when the minority has >= 8 / 12 lanes.  Every variant must give the same accumulators and ray
statistics (checked here by hash).  Prints one JSON line per (GI mode, policy)."""
import hashlib
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpuvoxelraycaster_b200 as vrt  # noqa: E402
from cpuvoxelraycaster_b200.frame import FrameRenderer  # noqa: E402


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


def main():
    D = int(os.environ.get("PROBE_DEPTH", "11"))
    S = float(1 << D)
    policies = [int(x) for x in os.environ.get("PROBE_POLICIES", "-1,0,1,2,3,4").split(",")]
    stream = torch.cuda.Stream()
    ctx = vrt.Context(0, stream.cuda_stream)
    scene = vrt.LSVO.from_terrain(ctx, D)
    t = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "textures.npz"))
    scene.set_textures(t["top"], t["side"])
    cam = vrt.Camera(position=(S / 2, S / 2 - 56, S / 2), view_angle=(0, 0), aperture=0.5)
    cam.autofocus(scene)
    cs = cam.as_struct()
    for world in (1, 8):
        fr = FrameRenderer(scene, 1920, 1080, 0, world, None, None, stream)
        fr.light = np.float32([-200, -1000, -300]) * np.float32(1.0 / S) + np.float32(1.0)
        for (gi, bounces) in ((False, 1), (True, 1), (True, 2)):
            if world == 8 and not (gi and bounces == 2):
                continue
            fr.use_gi, fr.gi_bounces = gi, bounces
            p = fr.params(64)
            ref = None
            for pol in policies:
                ctx.set_option("trav_policy", pol)
                ms = time_frame(fr, cs, p, stream)
                st = fr.stats()
                h = hashlib.sha256(fr.accum.cpu().numpy().tobytes()).hexdigest()[:16]
                sig = (h, tuple(st["rays"]), tuple(st["complexity"]))
                if ref is None:
                    ref = sig
                print(json.dumps(dict(world=world, gi=gi, bounces=bounces, policy=pol, ms=round(ms, 3), same=sig == ref,
                                      rays=sum(st["rays"]), trips=sum(st["complexity"]), hash=h)), flush=True)
    ctx.set_option("trav_policy", -1)


if __name__ == "__main__":
    main()
