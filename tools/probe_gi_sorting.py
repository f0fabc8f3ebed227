"""Would regrouping the GI rays of a pixel tile by direction pay?  For a 480x272 window of the 1080p frame, 64 GI rays per
primary hit (the estimator's distribution, cone coefficient 0.5) are cast with K1 (one thread per ray, a warp = 32
consecutive rays) in three orders:
  pixel     a warp = 32 samples of one pixel                       (what K4's 32-lanes-per-pixel mapping gives the GI stage)
  tile-dir  inside each 8x4-pixel tile (2048 rays) sorted by face normal, then by the angle of the tangent-plane noise
  dir-only  inside each tile sorted by (normal, c1, c2) lattice cell — the same direction from neighbouring pixels
Prints the time and loop-trip rate of each; exploratory (random numbers are numpy's, not the frame's Philox stream)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpuvoxelraycaster_b200 as vrt  # noqa: E402


def time_cast(ctx, scene, stream, o, d, iters=5):
    n = len(o)
    do, dd = torch.from_numpy(np.ascontiguousarray(o)).cuda(), torch.from_numpy(np.ascontiguousarray(d)).cuda()
    out = torch.empty(n * 16, dtype=torch.int32, device="cuda")
    with torch.cuda.stream(stream):
        for _ in range(2):
            scene.cast_rays_device(do, dd, n, out, 0.5, 0.0)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
        ev[0].record(stream)
        for i in range(iters):
            scene.cast_rays_device(do, dd, n, out, 0.5, 0.0)
            ev[i + 1].record(stream)
    stream.synchronize()
    ms = float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]))
    return ms, scene.last_complexity(), out.view(n, 16).cpu().numpy()


def main(D=11, W=1920, H=1080, X0=720, Y0=600, WW=480, HH=272, SPP=64):
    S = float(1 << D)
    stream = torch.cuda.Stream()
    ctx = vrt.Context(0, stream.cuda_stream)
    ctx.set_option("cast_variant", 0)
    scene = vrt.LSVO.from_terrain(ctx, D)
    cam = np.float32([S / 2, S / 2 - 56, S / 2])
    x, y = np.meshgrid(np.arange(X0, X0 + WW, dtype=np.float32), np.arange(Y0, Y0 + HH, dtype=np.float32))
    d = np.stack([x / np.float32(H) - np.float32(W / H * 0.5), y / np.float32(H) - np.float32(0.5), np.ones_like(x)], -1).reshape(-1, 3)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = np.broadcast_to(cam / np.float32(S) + np.float32(1), d.shape).astype(np.float32).copy()
    hits = scene.cast_rays(o, d.astype(np.float32))
    hit = (hits["flags"] & 1) != 0
    print(json.dumps(dict(window_pixels=int(hit.size), primary_hits=int(hit.sum()))), flush=True)
    # tile id of every pixel (8x4 tiles), then 64 samples per hit pixel
    px, py = (x.reshape(-1) - X0).astype(np.int64), (y.reshape(-1) - Y0).astype(np.int64)
    tile = (py // 4) * (WW // 8) + (px // 8)
    idx = np.flatnonzero(hit)
    pos, nrm, tl = np.repeat(hits["position"][idx], SPP, 0), np.repeat(hits["normal"][idx], SPP, 0), np.repeat(tile[idx], SPP)
    pix = np.repeat(idx, SPP)
    n = len(pos)
    rng = np.random.default_rng(1)
    ci = rng.integers(0, 100, (n, 2))
    c = (-1000 + 2000 * (ci / 100.0)).astype(np.float32)
    n_norm = np.float32(1.0 / S * 0.0078125 * 2.0)
    ax = np.argmax(nrm != 0, axis=1)
    others = np.array([[1, 2], [0, 2], [0, 1]])[ax]
    noise = np.zeros_like(nrm)
    noise[np.arange(n), others[:, 0]] = c[:, 0]
    noise[np.arange(n), others[:, 1]] = c[:, 1]
    gd = (nrm + noise) * n_norm
    gd /= np.linalg.norm(gd, axis=1, keepdims=True)
    go = (pos + nrm * n_norm).astype(np.float32)
    gd = gd.astype(np.float32)
    normal_id = ax * 2 + (nrm[np.arange(n), ax] > 0)
    angle = np.arctan2(c[:, 1], c[:, 0])
    orders = {
        "pixel": np.lexsort((np.arange(n), pix)),
        "tile-dir": np.lexsort((angle, normal_id, tl)),
        "dir-only": np.lexsort((pix, ci[:, 1], ci[:, 0], normal_id, tl)),
        "tile-dir-16bins-then-pixel": np.lexsort((pix, np.floor((angle + np.pi) / (2 * np.pi) * 16), normal_id, tl)),
        # what a kernel can know before tracing the primary ray: the noise angle only (256 bins), not the face normal
        "tile-angle256 (no normal)": np.lexsort((pix, np.floor((angle + np.pi) / (2 * np.pi) * 256), tl)),
        # the same, per (tile, half of the samples): 1024 rays sorted at a time
        "tile-half-angle256 (no normal)": np.lexsort((pix, np.floor((angle + np.pi) / (2 * np.pi) * 256), np.tile(np.arange(SPP) // 32, n // SPP), tl)),
    }
    light = np.float32([-200, -1000, -300]) / np.float32(512 * (1 << (D - 9))) * np.float32(1.0) + np.float32(1)
    for name, order in orders.items():
        ms, cx, rec = time_cast(ctx, scene, stream, go[order], gd[order])
        print(json.dumps(dict(order=name, rays=n, ms=round(ms, 3), mean_complexity=round(cx / n, 2), grays_s=round(n / ms / 1e6, 2),
                              giga_trips_s=round(cx / ms / 1e6, 1))), flush=True)
        # the GI-shadow stage as the same lanes would run it: from the GI hit toward the light; lanes whose GI ray missed idle
        # (a ray that starts outside the cube pointing away: one trip)
        h = rec.view(vrt.HIT).reshape(-1)
        ok = (h["flags"] & 1) != 0
        so = np.where(ok[:, None], h["position"] + h["normal"] * n_norm, np.float32(3.0)).astype(np.float32)
        sd = light[None, :] - so
        sd /= np.linalg.norm(sd, axis=1, keepdims=True)
        sd = np.where(ok[:, None], sd, np.float32(0.57735)).astype(np.float32)
        ms2, cx2, _ = time_cast(ctx, scene, stream, so, sd)
        print(json.dumps(dict(order=name, stage="gi-shadow in place", rays=int(ok.sum()), ms=round(ms2, 3),
                              giga_trips_s=round(cx2 / ms2 / 1e6, 1))), flush=True)


if __name__ == "__main__":
    main()
