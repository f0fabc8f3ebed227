"""How tight are the beam floors (beam_kernels.cu)?  cfg-4 camera at 2048^3: per tile size, the floors against the hit distance
of each pixel's centre ray (K1 cast), and the time of the beam kernel itself.  One JSON line per tile size."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpuvoxelraycaster_b200 as vrt  # noqa: E402


def main():
    D, W, H = 11, 1920, 1080
    S = float(1 << D)
    ctx = vrt.Context(0)
    scene = vrt.LSVO.from_terrain(ctx, D)
    t = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "textures.npz"))
    scene.set_textures(t["top"], t["side"])
    cam = vrt.Camera(position=(S / 2, S / 2 - 56, S / 2), view_angle=(0, 0), aperture=0.5)
    cam.autofocus(scene)
    x, y = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    d = np.stack([x / np.float32(H) - np.float32(W / H * 0.5), y / np.float32(H) - np.float32(0.5), np.ones_like(x)], -1).reshape(-1, 3)
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    o = np.broadcast_to(np.float32([S / 2, S / 2 - 56, S / 2]) / np.float32(S) + np.float32(1), d.shape).astype(np.float32).copy()
    hits = scene.cast_rays(o, d)
    hit = (hits["flags"] & 1) != 0
    t_hit = np.where(hit, hits["distance"], np.float32(10.0)).reshape(H, W)
    rc = vrt.RayCaster(scene, (W, H))
    for tile in (2, 4, 8, 16, 32, 64):
        import time
        rc.beam_floors(cam, tile)
        t0 = time.time()
        fl = rc.beam_floors(cam, tile)
        dt = time.time() - t0
        pp = np.repeat(np.repeat(fl, tile, 0), tile, 1)[:H, :W]
        m = t_hit < 10
        ratio = pp[m] / t_hit[m]
        print(json.dumps(dict(tile=tile, tiles=int(fl.size), wall_ms_incl_copy=round(dt * 1e3, 3), violations=int((pp > t_hit).sum()),
                              mean_floor_over_hit=round(float(ratio.mean()), 4), p10=round(float(np.percentile(ratio, 10)), 4),
                              p50=round(float(np.percentile(ratio, 50)), 4), zero_floors=int((pp[m] == 0).sum()),
                              sky_pixels_skipped=int(((pp >= 3.0) & ~m).sum()), sky_pixels=int((~m).sum()))), flush=True)


if __name__ == "__main__":
    main()
