"""Renders a few small frames on the GPU and writes PNGs (visual sanity for humans; parity is in tests/)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpuvoxelraycaster_b200 as vrt  # noqa: E402


def save(path, img):
    from PIL import Image
    Image.fromarray(img[..., :3]).save(path)


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    t = np.load(os.path.join(ROOT, "tests", "golden", "textures.npz"))
    ctx = vrt.Context(0)
    scene = vrt.LSVO(ctx, vrt.host_build_terrain_lsvo(9), 9)
    scene.set_textures(t["top"], t["side"])
    light = np.float32([-200, -1000, -300]) * np.float32(1 / 512.0) + np.float32(1)
    # cfg 1: the reference's default view, primary + sun shadow
    rc = vrt.RayCaster(scene, (640, 360))
    rc.setLightPosition(light)
    rc.use_samples = True
    save(os.path.join(out_dir, "cfg1_primary_shadow.png"), rc.render(vrt.Camera(position=(256, 200, 256), focal_length=100.0), 1))
    # GI + DOF, 64 spp
    rc = vrt.RayCaster(scene, (640, 360))
    rc.setLightPosition(light)
    rc.use_samples, rc.use_gi, rc.gi_bounces = True, True, 2
    cam = vrt.Camera(position=(256, 200, 256), view_angle=(0.3, -0.35), aperture=0.5)
    cam.autofocus(scene)
    save(os.path.join(out_dir, "gi2_dof_64spp.png"), rc.render(cam, 64))
    # grid with a mirror lake, blurry reflections; world "up" is -y like the reference (event_manager.hpp:125)
    h = vrt.host_terrain_heights(256)
    surface = 200 - np.clip(h, 0, 100)                      # first solid y of each column
    y = np.arange(256)[None, :, None]
    cells = (y >= surface[:, None, :]).astype(np.uint8)
    water = 176                                             # flood the valleys: a flat Cell::Mirror lake at y = water
    xs, zs = np.nonzero(surface > water)
    cells[xs, water + 1:, zs] = 1
    cells[xs, water, zs] = 2
    grid = vrt.MipmapGrid3D(ctx, cells, 3)
    grid.set_textures(t["top"], t["side"])
    rc = vrt.RayCaster(grid, (640, 360))
    rc.setLightPosition(np.float32([-300.0, -1500.0, -400.0]))
    rc.use_samples, rc.roughness = True, 0.06
    save(os.path.join(out_dir, "grid_mirror_reflections_32spp.png"),
         rc.render(vrt.Camera(position=(40.0, 160.0, 6.0), view_angle=(0.35, -0.12), focal_length=80.0), 32))
    print("gallery written to", out_dir)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/gallery")
