"""Is the first frame a fresh context renders the same as its later ones?  (One GPU; rank r of a 4-way tile split, K6.)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpuvoxelraycaster_b200 as vrt  # noqa: E402
from cpuvoxelraycaster_b200.frame import FrameRenderer  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
nodes = vrt.host_build_terrain_lsvo(9)
t = np.load(os.path.join(ROOT, "tests", "golden", "textures.npz"))
W, H, spp = 330, 187, 8
cam = vrt.Camera(position=(256, 200, 256), view_angle=(0.3, -0.35), aperture=0.5, focal_length=60.0)
light = np.float32([-200, -1000, -300]) * np.float32(1 / 512.0) + np.float32(1)
for trial in range(3):
    for rank in (1, 2, 3):
        stream = torch.cuda.Stream()
        ctx = vrt.Context(0, stream.cuda_stream)
        scene = vrt.LSVO(ctx, nodes, 9)
        scene.set_textures(t["top"], t["side"])
        fr = FrameRenderer(scene, W, H, rank, 4, None, None, stream, exchange="nccl")
        fr.use_gi, fr.gi_bounces, fr.light = True, 2, light
        imgs, accs = [], []
        for k in range(3):
            p = fr.params(spp, 0)
            with torch.cuda.stream(stream):
                fr.accum.zero_()
                fr.accumulate(cam.as_struct(), p)
                fr.resolve(p)
            stream.synchronize()
            imgs.append(fr.rgba.cpu().numpy().reshape(-1, W, 4).copy())
            accs.append(fr.accum.cpu().numpy().copy())
        print("trial %d rank %d: image 0==2 %s 1==2 %s | accum 0==2 %s 1==2 %s | nonzero px %d" % (
            trial, rank, np.array_equal(imgs[0], imgs[2]), np.array_equal(imgs[1], imgs[2]), np.array_equal(accs[0], accs[2]),
            np.array_equal(accs[1], accs[2]), int((imgs[2][..., :3].max(axis=2) > 0).sum())), flush=True)
        scene.close()
        ctx.close()
