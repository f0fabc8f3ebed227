"""Why does the GI part of the frame run at ~11 of 32 lanes?  Casts, for a subset of the 1080p primary hits, 32 GI rays
each (the estimator's distribution, cone coefficient 0.5) and prints the distribution of loop trips per ray and the
ratio E[max over the 32 lanes of a warp] / mean — the lane idleness a one-ray-per-lane mapping must pay."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpuvoxelraycaster_b200 as vrt  # noqa: E402


def main(D=11, W=1920, H=1080):
    S = float(1 << D)
    ctx = vrt.Context(0)
    scene = vrt.LSVO.from_terrain(ctx, D)
    cam = np.float32([S / 2, S / 2 - 56, S / 2])
    x, y = np.meshgrid(np.arange(0, W, 4, dtype=np.float32), np.arange(0, H, 4, dtype=np.float32))
    d = np.stack([x / np.float32(H) - np.float32(W / H * 0.5), y / np.float32(H) - np.float32(0.5), np.ones_like(x)], -1).reshape(-1, 3)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = np.broadcast_to(cam / np.float32(S) + np.float32(1), d.shape).astype(np.float32).copy()
    hits = scene.cast_rays(o, d.astype(np.float32))
    m = (hits["flags"] & 1) != 0
    pos, nrm = np.repeat(hits["position"][m], 32, axis=0), np.repeat(hits["normal"][m], 32, axis=0)
    n_norm = np.float32(1.0 / S * 0.0078125 * 2.0)
    rng = np.random.default_rng(1)
    c = (-1000 + 2000 * (rng.integers(0, 100, (len(pos), 2)) / 100.0)).astype(np.float32)
    noise = np.zeros_like(nrm)
    ax = np.argmax(nrm != 0, axis=1)
    others = np.array([[1, 2], [0, 2], [0, 1]])[ax]
    noise[np.arange(len(pos)), others[:, 0]] = c[:, 0]
    noise[np.arange(len(pos)), others[:, 1]] = c[:, 1]
    gd = (nrm + noise) * n_norm
    gd /= np.linalg.norm(gd, axis=1, keepdims=True)
    go = pos + nrm * n_norm
    light = np.float32([-200, -1000, -300]) / np.float32(512) + np.float32(1)
    for name, oo, dd in (("gi", go, gd),):
        h = scene.cast_rays(oo.astype(np.float32), dd.astype(np.float32), 0.5, 0.0)
        cx = h["complexity"].astype(np.int64)
        g = cx.reshape(-1, 32)
        print(json.dumps(dict(rays=name, n=int(len(cx)), mean=float(cx.mean()), p50=float(np.percentile(cx, 50)),
                              p90=float(np.percentile(cx, 90)), p99=float(np.percentile(cx, 99)), max=int(cx.max()),
                              mean_of_warp_max=float(g.max(1).mean()), lane_utilisation=float(cx.mean() / g.max(1).mean()),
                              hit_fraction=float(((h["flags"] & 1) != 0).mean()),
                              hist=np.bincount(np.minimum(cx, 99) // 5).tolist())), flush=True)
        hm = (h["flags"] & 1) != 0
        so = h["position"][hm] + h["normal"][hm] * n_norm
        sd = light - so
        sd /= np.linalg.norm(sd, axis=1, keepdims=True)
        hs = scene.cast_rays(so.astype(np.float32), sd.astype(np.float32), 0.5, 0.0)
        cs = hs["complexity"].astype(np.int64)
        k = len(cs) // 32 * 32
        gs = cs[:k].reshape(-1, 32)
        print(json.dumps(dict(rays="gi_shadow (compacted)", n=int(len(cs)), mean=float(cs.mean()), p50=float(np.percentile(cs, 50)),
                              p90=float(np.percentile(cs, 90)), p99=float(np.percentile(cs, 99)), max=int(cs.max()),
                              mean_of_warp_max=float(gs.max(1).mean()), lane_utilisation=float(cs.mean() / gs.max(1).mean()),
                              hist=np.bincount(np.minimum(cs, 99) // 5).tolist())), flush=True)


if __name__ == "__main__":
    main()
