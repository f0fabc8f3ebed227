"""Times K4 on the slice one of N GPUs would own (4-row tiles t % N == 0 of the cfg-4 frame) for combinations of
spp_chunks / samples_per_warp — tuning aid for the launcher's automatic choice."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpuvoxelraycaster_b200 as vrt  # noqa: E402
from cpuvoxelraycaster_b200.frame import FrameRenderer  # noqa: E402


def main():
    D, S = 11, 2048.0
    stream = torch.cuda.Stream()
    ctx = vrt.Context(0, stream.cuda_stream)
    scene = vrt.LSVO.from_terrain(ctx, D)
    t = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "textures.npz"))
    scene.set_textures(t["top"], t["side"])
    cam = vrt.Camera(position=(S / 2, S / 2 - 56, S / 2), view_angle=(0, 0), aperture=0.5)
    cam.autofocus(scene)
    for world in (1, 8):
        fr = FrameRenderer(scene, 1920, 1080, 0, world, None, None, stream)
        fr.use_gi, fr.gi_bounces = True, 2
        fr.light = np.float32([-200, -1000, -300]) * np.float32(1.0 / S) + np.float32(1.0)
        p, cs = fr.params(64), cam.as_struct()
        for chunks, q in ((0, 0), (1, 32), (2, 32), (4, 16), (8, 8), (16, 4), (32, 2), (64, 1), (29, 1), (4, 1), (16, 1), (32, 1)):
            ctx.set_option("spp_chunks", chunks)
            ctx.set_option("samples_per_warp", q)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            with torch.cuda.stream(stream):
                fr.accum.zero_()
                fr.accumulate(cs, p)
                for i in range(4):
                    fr.accum.zero_()
                    ev[i].record(stream)
                    fr.accumulate(cs, p)
                ev[4].record(stream)
            stream.synchronize()
            ms = float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(3)]))
            print(json.dumps(dict(world=world, spp_chunks=chunks, samples_per_warp=q, ms=round(ms, 3))), flush=True)


if __name__ == "__main__":
    main()
