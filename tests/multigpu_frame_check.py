"""Run under torchrun (one rank per GPU): every rank renders its share of a frame — 4-row tiles dealt round-robin, or
spp / world of the samples — the frame is assembled on rank 0 by libvrt's communicator (peer stores over NVLink) or
all-gathered over NCCL, and rank 0 compares it byte for byte with the same frame rendered alone on one GPU.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multigpu_frame_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpuvoxelraycaster_b200 as vrt  # noqa: E402
from cpuvoxelraycaster_b200.frame import FrameRenderer  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    stream = torch.cuda.Stream(device)
    ctx = vrt.Context(local, stream.cuda_stream)
    for kv in filter(None, os.environ.get("FRAME_CHECK_OPTS", "").split(",")):      # diagnostics: context options, e.g. beam_tile=0
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    scene = vrt.LSVO(ctx, vrt.host_build_terrain_lsvo(9), 9)
    t = np.load(os.path.join(ROOT, "tests", "golden", "textures.npz"))
    scene.set_textures(t["top"], t["side"])
    W, H = 330, 187                                          # ragged on purpose
    spp = 2 * world                                          # the sample split needs spp % world == 0
    cam = vrt.Camera(position=(256, 200, 256), view_angle=(0.3, -0.35), aperture=0.5, focal_length=60.0)
    light = np.float32([-200, -1000, -300]) * np.float32(1 / 512.0) + np.float32(1)

    def frames(r, w, exchange, split, n_frames=3):
        """n_frames consecutive frames (different sample offsets: exercises the double buffering); returns rank 0's copies."""
        fr = FrameRenderer(scene, W, H, r, w, None, device, stream, exchange=exchange, split=split)
        fr.use_gi, fr.gi_bounces, fr.light = True, 2, light
        out, rays = [], []
        for k in range(n_frames):
            f = fr.render(cam, spp, sample_offset=k * spp)
            out.append(None if f is None else f.copy())
            rays.append(fr.stats()["rays"])
        if fr.peer is not None:
            fr.peer.close()
        return out, rays

    ok = True
    if os.environ.get("FRAME_CHECK_LOCAL"):                  # diagnostics: is a rank's OWN first frame right before any exchange?
        ref, _ = frames(0, 1, "nccl", "tiles", 1)
        fr = FrameRenderer(scene, W, H, rank, world, None, device, stream, exchange="nccl", split="tiles")
        fr.use_gi, fr.gi_bounces, fr.light = True, 2, light
        for k in range(2):
            p = fr.params(spp, 0)
            with torch.cuda.stream(stream):
                fr.accum.zero_()
                fr.accumulate(cam.as_struct(), p)
                fr.resolve(p)
            stream.synchronize()
            mine = fr.rgba.cpu().numpy().reshape(-1, W, 4)[:H]
            rows = [y for y in range(H) if (y >> 2) % world == rank]
            bad = [y for y in rows if not np.array_equal(mine[y], ref[0][y])]
            print("rank %d local frame %d: %d of %d own rows differ %s; sample mine %s ref %s" % (
                rank, k, len(bad), len(rows), bad[:6], mine[bad[0], 100].tolist() if bad else "", ref[0][bad[0], 100].tolist() if bad else ""), flush=True)
    single, rays1 = frames(0, 1, "nccl", "tiles") if rank == 0 else (None, None)
    for exchange, split in (("peer", "tiles"), ("peer", "samples"), ("nccl", "tiles")):
        multi, rays = frames(rank, world, exchange, split)
        tot = torch.tensor(rays, dtype=torch.int64, device=device)
        dist.all_reduce(tot)
        if rank == 0:
            same = all(np.array_equal(a, b) for a, b in zip(multi, single)) and tot.tolist() == rays1 and int(multi[0][..., :3].max()) > 0
            print("multigpu frame check world=%d exchange=%s split=%s: %s" % (world, exchange, split, "OK" if same else "MISMATCH"), flush=True)
            if not same:                                     # say what differs: which frames, which rows, the ray totals
                for k, (a, b) in enumerate(zip(multi, single)):
                    rows = np.flatnonzero((a != b).reshape(H, -1).any(axis=1))
                    print("  frame %d: %d differing rows %s; rays %s vs %s" % (k, rows.size, rows[:12].tolist(), tot.tolist()[k], rays1[k]), flush=True)
                    if rows.size:
                        y = int(rows[0])
                        print("    row %d, pixels 100..103: multi %s single %s" % (y, a[y, 100:104].tolist(), b[y, 100:104].tolist()), flush=True)
            ok = ok and same
    flag = torch.tensor([1 if ok else 0], device=device)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
