"""Would a wavefront GI pass pay?  Generates the first-bounce GI rays of a 1080p frame (1 spp) on the host from the
primary hits, then times them through the batched kernels: K1 (one thread per ray) and K1p (persistent, regenerating),
cone coefficient 0.5 like raycaster.hpp:194.  Compare the loop-trip rate with what K4 achieves on the same ray class in
place (tools/probe_render.py).  Exploratory, not a parity test: the rays follow the estimator's distribution."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpuvoxelraycaster_b200 as vrt  # noqa: E402


def main(D=11, W=1920, H=1080, iters=7):
    S = float(1 << D)
    stream = torch.cuda.Stream()
    ctx = vrt.Context(0, stream.cuda_stream)
    scene = vrt.LSVO.from_terrain(ctx, D)
    cam = np.float32([S / 2, S / 2 - 56, S / 2])
    x, y = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    d = np.stack([x / np.float32(H) - np.float32(W / H * 0.5), y / np.float32(H) - np.float32(0.5), np.ones_like(x)], -1).reshape(-1, 3)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = np.broadcast_to(cam / np.float32(S) + np.float32(1), d.shape).astype(np.float32).copy()
    hits = scene.cast_rays(o, d.astype(np.float32))
    m = (hits["flags"] & 1) != 0
    pos, nrm = hits["position"][m], hits["normal"][m]
    n_norm = np.float32(1.0 / S * 0.0078125 * 2.0)
    rng = np.random.default_rng(1)
    c = (-1000 + 2000 * (rng.integers(0, 100, (len(pos), 2)) / 100.0)).astype(np.float32)
    noise = np.zeros_like(nrm)
    ax = np.argmax(nrm != 0, axis=1)                     # the face axis; noise lives in the other two
    others = np.array([[1, 2], [0, 2], [0, 1]])[ax]
    noise[np.arange(len(pos)), others[:, 0]] = c[:, 0]
    noise[np.arange(len(pos)), others[:, 1]] = c[:, 1]
    gd = (nrm + noise) * n_norm
    gd /= np.linalg.norm(gd, axis=1, keepdims=True)
    go = pos + nrm * n_norm
    n = len(go)
    do, dd = torch.from_numpy(go.astype(np.float32)).cuda(), torch.from_numpy(gd.astype(np.float32)).cuda()
    out = torch.empty(n * 16, dtype=torch.int32, device="cuda")
    for name, variant in (("K1 one thread per ray", 0), ("K1p persistent regenerating", 1)):
        ctx.set_option("cast_variant", variant)
        with torch.cuda.stream(stream):
            for _ in range(3):
                scene.cast_rays_device(do, dd, n, out, 0.5, 0.0)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
            ev[0].record(stream)
            for i in range(iters):
                scene.cast_rays_device(do, dd, n, out, 0.5, 0.0)
                ev[i + 1].record(stream)
        stream.synchronize()
        ms = float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]))
        cx = scene.last_complexity()
        print(json.dumps(dict(kernel=name, gi_rays=n, ms=round(ms, 4), mean_complexity=round(cx / n, 2),
                              grays_s=round(n / ms / 1e6, 2), giga_trips_s=round(cx / ms / 1e6, 1))), flush=True)


if __name__ == "__main__":
    main()
