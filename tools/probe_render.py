"""Exploratory timing of the frame kernel per ray class mix (not the bench): primary+shadow only, +GI 1 bounce,
+GI 2 bounces, with/without DOF, on T(D) at 1080p."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpuvoxelraycaster_b200 as vrt  # noqa: E402
from cpuvoxelraycaster_b200.frame import FrameRenderer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depth", type=int, default=11)
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    D, S = a.depth, 1 << a.depth
    stream = torch.cuda.Stream()
    ctx = vrt.Context(0, stream.cuda_stream)
    scene = vrt.LSVO(ctx, vrt.host_build_terrain_lsvo(D), D)
    t = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "textures.npz"))
    scene.set_textures(t["top"], t["side"])
    fr = FrameRenderer(scene, 1920, 1080, 0, 1, None, None, stream)
    fr.light = np.float32([-200, -1000, -300]) * np.float32(1.0 / S) + np.float32(1.0)
    for name, use_gi, bounces, ap_ in (("primary+shadow", False, 1, 0.0), ("primary+shadow DOF", False, 1, 0.5),
                                       ("GI 1 bounce DOF", True, 1, 0.5), ("GI 2 bounces DOF", True, 2, 0.5)):
        cam = vrt.Camera(position=(S / 2, S / 2 - 56, S / 2), view_angle=(0, 0), aperture=ap_, focal_length=209.0)
        fr.use_gi, fr.gi_bounces = use_gi, bounces
        for _ in range(2):
            fr.render_device(cam, a.spp)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters + 1)]
        with torch.cuda.stream(stream):
            ev[0].record(stream)
            for i in range(a.iters):
                fr.render_device(cam, a.spp)
                ev[i + 1].record(stream)
        stream.synchronize()
        ms = float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(a.iters)]))
        st = fr.stats()
        rays = sum(st["rays"])
        print(json.dumps(dict(mode=name, spp=a.spp, ms=round(ms, 3), rays=st["rays"], grays_s=round(rays / ms / 1e6, 2),
                              mean_complexity=[round(c / max(1, r), 1) for c, r in zip(st["complexity"], st["rays"])])), flush=True)


if __name__ == "__main__":
    main()
