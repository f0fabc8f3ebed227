"""Times the direction-sorted frame kernel K5 on the cfg-4 frame (and on the 1/8 slice an 8-GPU rank owns) for
combinations of sample runs and angle bins; K4 (render_variant 2) as the baseline; plus per-class breakdowns
(GI off / 1 bounce / 2 bounces)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpuvoxelraycaster_b200 as vrt  # noqa: E402
from cpuvoxelraycaster_b200.frame import FrameRenderer  # noqa: E402


def time_frame(ctx, fr, cs, p, stream):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    with torch.cuda.stream(stream):
        fr.accum.zero_()
        fr.accumulate(cs, p)
        for i in range(3):
            fr.accum.zero_()
            ev[i].record(stream)
            fr.accumulate(cs, p)
        ev[3].record(stream)
    stream.synchronize()
    return float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(2)]))


def main():
    D, S = 11, 2048.0
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
            fr.use_gi, fr.gi_bounces = gi, bounces
            p = fr.params(64)
            combos = [(2, 0, 0, 0), (3, 0, 0, 0)]
            if gi and bounces == 2:
                combos += [(3, 1, 8, 1), (3, 1, 16, 1), (3, 1, 32, 1), (3, 1, 16, 4), (3, 2, 4, 1), (3, 2, 8, 1), (3, 2, 16, 1), (3, 2, 32, 1),
                           (3, 2, 128, 1), (3, 2, 8, 4), (3, 2, 16, 4), (3, 2, 1, 1), (3, 4, 8, 1), (3, 4, 16, 1), (3, 8, 8, 1), (3, 8, 16, 1)]
            elif gi:
                combos += [(3, 2, 8, 1), (3, 2, 16, 1), (3, 2, 32, 1), (3, 2, 1, 1)]
            else:
                combos += [(3, 2, 1, 1)]
            for variant, chunks, b1, b2 in combos:
                ctx.set_option("render_variant", variant)
                ctx.set_option("spp_chunks", chunks)
                ctx.set_option("sort_bins1", b1)
                ctx.set_option("sort_bins2", b2)
                ms = time_frame(ctx, fr, cs, p, stream)
                print(json.dumps(dict(world=world, gi=gi, bounces=bounces, variant=variant, spp_chunks=chunks, bins=[b1, b2], ms=round(ms, 3))), flush=True)


if __name__ == "__main__":
    main()
