"""Run under torchrun (one rank per GPU): every rank renders its round-robin tiles, the frame is all-gathered over
NCCL, and rank 0 compares it byte for byte with the same frame rendered alone on one GPU.
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
    scene = vrt.LSVO(ctx, vrt.host_build_terrain_lsvo(9), 9)
    t = np.load(os.path.join(ROOT, "tests", "golden", "textures.npz"))
    scene.set_textures(t["top"], t["side"])
    W, H, spp = 330, 187, 3                                  # ragged on purpose
    cam = vrt.Camera(position=(256, 200, 256), view_angle=(0.3, -0.35), aperture=0.5, focal_length=60.0)
    light = np.float32([-200, -1000, -300]) * np.float32(1 / 512.0) + np.float32(1)

    def frame(r, w):
        fr = FrameRenderer(scene, W, H, r, w, None, device, stream)
        fr.use_gi, fr.gi_bounces, fr.light = True, 2, light
        return fr.render(cam, spp).copy(), fr.stats()

    multi, st = frame(rank, world)
    rays = torch.tensor(st["rays"], dtype=torch.int64, device=device)
    dist.all_reduce(rays)
    ok = True
    if rank == 0:
        single, st1 = frame(0, 1)
        ok = np.array_equal(multi, single) and rays.tolist() == st1["rays"] and int(multi[..., :3].max()) > 0
        print("multigpu frame check world=%d: %s (rays %s)" % (world, "OK" if ok else "MISMATCH", rays.tolist()), flush=True)
    flag = torch.tensor([1 if ok else 0], device=device)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
