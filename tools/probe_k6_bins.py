import sys, os, json, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
import cpuvoxelraycaster_b200 as vrt
from cpuvoxelraycaster_b200.frame import FrameRenderer
from probe_sorted import time_frame
D, S = 11, 2048.0
stream = torch.cuda.Stream()
ctx = vrt.Context(0, stream.cuda_stream)
scene = vrt.LSVO.from_terrain(ctx, D)
t = np.load('/root/repo/tests/golden/textures.npz'); scene.set_textures(t["top"], t["side"])
cam = vrt.Camera(position=(S / 2, S / 2 - 56, S / 2), view_angle=(0, 0), aperture=0.5); cam.autofocus(scene); cs = cam.as_struct()
for world in (1, 8):
    fr = FrameRenderer(scene, 1920, 1080, 0, world, None, None, stream)
    fr.light = np.float32([-200, -1000, -300]) * np.float32(1.0 / S) + np.float32(1.0)
    fr.use_gi, fr.gi_bounces = True, 2
    p = fr.params(64)
    ctx.set_option("render_variant", 4)
    for b1, b2 in ((0, 0), (8, 1), (12, 1), (16, 1), (24, 1), (32, 1), (64, 1), (16, 2), (8, 2), (16, 4), (32, 2)):
        ctx.set_option("sort_bins1", b1); ctx.set_option("sort_bins2", b2)
        print(json.dumps(dict(world=world, bins=[b1, b2], ms=round(time_frame(ctx, fr, cs, p, stream), 3))), flush=True)
