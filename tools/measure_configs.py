"""Measures the BASELINE.json configurations that are not the bench headline (cfg 1, 2, 3, 5) on one GPU:
device-resident inputs, CUDA events, median of N.  Prints one JSON line per configuration.
(cfg 4 is bench.py.)  Used for profiles/rNN_summary.md; parity for the same configurations is in tests/."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpuvoxelraycaster_b200 as vrt  # noqa: E402
from cpuvoxelraycaster_b200.frame import FrameRenderer  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def timed(stream, fn, iters):
    for _ in range(2):
        fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for i in range(iters):
            fn()
            ev[i + 1].record(stream)
    stream.synchronize()
    return float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]))


def camera_rays(D, W, H, voxel_units=False):
    S = float(1 << D)
    cam = np.float32([S / 2, S / 2 - 56, S / 2])
    x, y = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    d = np.stack([x / np.float32(H) - np.float32(W / H * 0.5), y / np.float32(H) - np.float32(0.5), np.ones_like(x)], -1).reshape(-1, 3)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = cam if voxel_units else cam / np.float32(S) + np.float32(1)
    return np.broadcast_to(o, d.shape).astype(np.float32).copy(), d.astype(np.float32)


def terrain_grid(size, mirrored):
    h = vrt.host_terrain_heights(size)
    hm = np.maximum(16, np.minimum(size, h))
    y = np.arange(size)[None, :, None]
    g = ((y >= size // 2 + 1) & (y <= (size // 2 + hm - 1)[:, None, :])).astype(np.uint8)
    return np.ascontiguousarray(g[::-1, ::-1, ::-1]) if mirrored else g


def measure(ctx, stream, want, iters=5, cfg5_rays=100_000_000, emit=None):
    """Runs the configurations in `want` on `ctx` / `stream`; returns the result dicts (and passes each to `emit`)."""
    class _A:
        pass
    a = _A()
    a.iters, a.cfg5_rays = iters, cfg5_rays
    tex = np.load(os.path.join(ROOT, "tests", "golden", "textures.npz"))
    results = []

    def emit_line(d):
        results.append(d)
        if emit:
            emit(d)

    if 1 in want:   # default terrain in LSVO<9>, primary + sun shadow, 1280x720, 1 spp
        scene = vrt.LSVO(ctx, vrt.host_build_terrain_lsvo(9), 9)
        scene.set_textures(tex["top"], tex["side"])
        fr = FrameRenderer(scene, 1280, 720, 0, 1, None, None, stream)
        fr.light = np.float32([-200, -1000, -300]) * np.float32(1 / 512.0) + np.float32(1)
        cam = vrt.Camera(position=(256, 200, 256), view_angle=(0, 0), focal_length=100.0)
        ms = timed(stream, lambda: fr.render_device(cam, 1), a.iters)
        st = fr.stats()
        rays = sum(st["rays"])
        emit_line(dict(cfg=1, what="T(9) LSVO, 1280x720, 1 spp, primary + sun shadow (K4 + resolve)", ms_per_frame=round(ms, 4),
                              rays=st["rays"][:2], mrays_s=round(rays / ms / 1e3, 1),
                              algo_GBs=round((8 * sum(st["complexity"]) + 64 * rays + 16 * 1280 * 720) / ms / 1e6, 1)))
        scene.close()

    if 2 in want or 3 in want:
        for cfg, size, W, H, mip in ((2, 512, 1920, 1080, 0), (3, 1024, 3840, 2160, 4)):
            if cfg not in want:
                continue
            D = size.bit_length() - 1
            cells = terrain_grid(size, mirrored=True)       # the same world the LSVO shows
            scene = vrt.MipmapGrid3D(ctx, cells, mip) if mip else vrt.Grid3D(ctx, cells)
            o, d = camera_rays(D, W, H, voxel_units=True)
            n = len(o)
            do, dd = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
            out = torch.empty(n * 16, dtype=torch.int32, device="cuda")
            ms = timed(stream, lambda: scene.cast_rays_device(do, dd, n, out), a.iters)
            steps = scene.last_complexity()
            hits = int((out.view(n, 16)[:, 10] & 1).sum())
            line = dict(cfg=cfg, what="T(%d) dense grid %d^3 (%s), %dx%d primary rays" % (D, size, "mip pyramid, %d levels" % mip if mip else "flat", W, H),
                        ms=round(ms, 4), rays=n, hits=hits, mrays_s=round(n / ms / 1e3, 1), mean_steps=round(steps / n, 1),
                        algo_GBs=round((steps + 64 * n) / ms / 1e6, 1), device_MB=round(scene.info()["device_bytes"] / 1e6, 1))
            if mip:   # same rays through the flat grid: identical records, fetch counts differ
                flat = vrt.Grid3D(ctx, cells)
                out2 = torch.empty_like(out)
                ms_flat = timed(stream, lambda: flat.cast_rays_device(do, dd, n, out2), a.iters)
                line.update(ms_flat_grid=round(ms_flat, 4), identical_to_flat=bool(torch.equal(out, out2)))
                flat.close()
            emit_line(line)
            scene.close()
            if cfg == 3:
                # the configuration's frame: primary + sun shadow + blurry reflections off a Cell::Mirror lake that floods
                # the valleys (columns whose top lies below the water level), 3840x2160, device resident
                h = vrt.host_terrain_heights(size)[::-1, ::-1]
                top = size // 2 - np.maximum(16, np.minimum(size, h))          # first solid y of each column (up = -y)
                water = size // 2 - 30
                xs, zs = np.nonzero(top > water)
                for x, z in zip(xs.tolist(), zs.tolist()):
                    cells[x, water + 1:top[x, z], z] = 1
                cells[xs, water, zs] = 2
                lake = vrt.MipmapGrid3D(ctx, cells, mip)
                lake.set_textures(tex["top"], tex["side"])
                fr = FrameRenderer(lake, W, H, 0, 1, None, None, stream)
                fr.use_gi, fr.roughness, fr.max_bounds = False, 0.06, 4
                fr.light = np.float32([-200, -1000, -300]) * np.float32(size / 512.0)
                cam = vrt.Camera(position=(size / 2, size / 2 - 56, size / 2), view_angle=(0.0, 0.0), focal_length=100.0)
                for spp in (1, 4):
                    ms = timed(stream, lambda: fr.render_device(cam, spp), a.iters)
                    st = fr.stats()
                    rays, steps = sum(st["rays"][:3]), sum(st["complexity"][:3])
                    emit_line(dict(cfg=3, what="T(%d) mip grid %d^3 with a mirror lake, %dx%d frame: primary + sun shadow + blurry reflections "
                                          "(roughness 0.06, max_bounds 4), %d spp" % (D, size, W, H, spp), ms_per_frame=round(ms, 4),
                                          rays=dict(primary=st["rays"][0], shadow=st["rays"][1], reflection=st["rays"][2]),
                                          mirror_cells=int(len(xs)), mrays_s=round(rays / ms / 1e3, 1), mean_steps=round(steps / max(rays, 1), 1),
                                          algo_GBs=round((steps + 64 * rays + 16 * W * H) / ms / 1e6, 1)))
                lake.close()
            del cells

    if 5 in want:   # LSVO 4096^3, incoherent random rays
        t0 = time.time()
        scene = vrt.LSVO.from_terrain(ctx, 12, guard=-1)   # built on the GPU; lsvo.hpp:72 guard lifted, see DESIGN.md §2
        ctx.synchronize()
        tb = time.time() - t0
        n_slots = len(scene)
        n = a.cfg5_rays
        g = torch.Generator(device="cuda").manual_seed(0xD1CE)
        o = torch.rand(n, 3, device="cuda", generator=g)
        o[:, 0] += 1.0
        o[:, 2] += 1.0
        o[:, 1] = 1.0 + o[:, 1] * (0.5 - 96.0 / 4096.0)
        d = torch.randn(n, 3, device="cuda", generator=g)
        d /= d.norm(dim=1, keepdim=True)
        out = torch.empty(n * 16, dtype=torch.int32, device="cuda")
        res = {}
        for variant in (3, 1, 0):                                    # 3 = automatic (the default): classifier + gated K1b / K1p
            ctx.set_option("cast_variant", variant)
            ms = timed(stream, lambda: scene.cast_rays_device(o, d, n, out), max(2, a.iters // 2))
            res[variant] = ms
        cx = scene.last_complexity()
        hits = int((out.view(n, 16)[:, 10] & 1).sum())
        # end to end through vrt_cast_rays with HOST buffers: 24 B per ray in, 64 B per hit record out over PCIe (pageable numpy)
        ne = min(n, 10_000_000)
        ho, hd = o[:ne].cpu().numpy(), d[:ne].cpu().numpy()
        scene.cast_rays(ho[:1000], hd[:1000])
        t0 = time.time()
        scene.cast_rays(ho, hd)
        e2e_s = time.time() - t0
        emit_line(dict(cfg=5, what="T(12) LSVO 4096^3, %d random rays" % n, device_build_s=round(tb, 3), slots=n_slots,
                              ms_automatic=round(res[3], 3), ms_persistent_adaptive=round(res[1], 3), ms_one_thread_per_ray=round(res[0], 3),
                              mrays_s=round(n / res[3] / 1e3, 1), hit_fraction=round(hits / n, 4), mean_complexity=round(cx / n, 2),
                              e2e_host_buffers=dict(rays=ne, ms=round(e2e_s * 1e3, 1), mrays_s=round(ne / e2e_s / 1e6, 1), bytes_per_ray=88),
                              algo_GBs=round((8 * cx + 64 * n) / res[3] / 1e6, 1)))
        scene.close()
        ctx.set_option("cast_variant", 3)
    return results


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,5")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--cfg5-rays", type=int, default=100_000_000)
    a = ap.parse_args()
    stream = torch.cuda.Stream()
    ctx = vrt.Context(0, stream.cuda_stream)
    measure(ctx, stream, [int(c) for c in a.configs.split(",")], a.iters, a.cfg5_rays, emit=lambda d: print(json.dumps(d), flush=True))
    ctx.close()


if __name__ == "__main__":
    main()
