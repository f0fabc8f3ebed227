"""Exploratory timing of the batched LSVO cast kernel (not the bench): coherent primary rays and
incoherent random rays on T(D), device-resident buffers, CUDA events on the launching stream."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpuvoxelraycaster_b200 as vrt  # noqa: E402


def primary_rays(D, W, H):
    S = float(1 << D)
    cam = np.float32([S / 2, S / 2 - 56, S / 2])
    x, y = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    lx = x / np.float32(H) - np.float32(W / H * 0.5)
    ly = y / np.float32(H) - np.float32(0.5)
    d = np.stack([lx, ly, np.ones_like(lx)], -1).reshape(-1, 3)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = np.broadcast_to(cam / np.float32(S) + np.float32(1), d.shape).copy()
    return o.astype(np.float32), d.astype(np.float32)


def random_rays(D, n, seed=0xD1CE):
    S = float(1 << D)
    rng = np.random.default_rng(seed)
    o = rng.uniform([1, 1, 1], [2, 1.5 - 96 / S, 2], (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o, d.astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depths", default="9,11")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--random", type=int, default=1 << 22)
    a = ap.parse_args()
    torch.cuda.init()
    stream = torch.cuda.Stream()
    ctx = vrt.Context(0, stream.cuda_stream)
    for D in [int(x) for x in a.depths.split(",")]:
        t = time.time()
        nodes = vrt.host_build_terrain_lsvo(D)
        tb = time.time() - t
        scene = vrt.LSVO(ctx, nodes, D, guard=-1 if D >= 12 else 0)
        for name, (o, d) in (("primary_1080p", primary_rays(D, 1920, 1080)), ("random", random_rays(D, a.random))):
            n = len(o)
            do, dd = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
            out = torch.empty(n * 16, dtype=torch.int32, device="cuda")
            with torch.cuda.stream(stream):
                for _ in range(3):
                    scene.cast_rays_device(do, dd, n, out)
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters + 1)]
                ev[0].record(stream)
                for i in range(a.iters):
                    scene.cast_rays_device(do, dd, n, out)
                    ev[i + 1].record(stream)
            stream.synchronize()
            ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.iters)]
            cx = scene.last_complexity()
            best, med = min(ms), float(np.median(ms))
            print(json.dumps(dict(depth=D, rays=name, n=n, build_s=round(tb, 2), slots=len(nodes), ms_best=round(best, 4),
                                  ms_median=round(med, 4), mrays_s=round(n / med / 1e3, 1), mean_complexity=round(cx / n, 2),
                                  algo_GBs=round((8 * cx + 64 * n) / med / 1e6, 1))), flush=True)
        scene.close()
    ctx.close()


if __name__ == "__main__":
    main()
