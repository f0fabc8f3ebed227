"""Dynamic scenes: time of one terrain edit = upload of the changed column heights + re-flattening of
the whole world on the device (vrt_scene_edit_heights), wall clock around the synchronous call."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpuvoxelraycaster_b200 as vrt  # noqa: E402


def main():
    ctx = vrt.Context(0)
    for depth in (9, 10, 11):
        S = 1 << depth
        scene = vrt.LSVO.from_heightfield(ctx, depth)
        h = scene.heights()
        for size in (16, 256):
            times = []
            for k in range(5):
                x0, z0 = (37 * k) % (S - size), (91 * k) % (S - size)
                patch = (h[x0:x0 + size, z0:z0 + size] + 12 * ((k % 2) * 2 - 1)).astype(np.int32)   # raise / dig by 12 voxels
                t0 = time.perf_counter()
                scene.edit_heights(x0, z0, patch)
                times.append((time.perf_counter() - t0) * 1e3)
            print(json.dumps(dict(depth=depth, world="%d^3" % S, edited_columns=size * size, slots=scene.n_nodes,
                                  ms_per_edit=round(float(np.median(times)), 3))), flush=True)
        scene.close()


def voxel_sets():
    """Arbitrary voxel sets: the T(9) terrain as a plain voxel list (8.6 M voxels), flattened on the device; then edits."""
    ctx = vrt.Context(0)
    S, depth = 512, 9
    h = vrt.host_terrain_heights(S)
    hmax = np.maximum(16, np.minimum(S, h)).astype(np.int64)
    xs, zs = np.meshgrid(np.arange(S), np.arange(S), indexing="ij")
    cols = np.repeat(np.stack([xs.reshape(-1), zs.reshape(-1)], 1), hmax.reshape(-1) - 1, axis=0)
    ys = np.concatenate([np.arange(1, m) for m in hmax.reshape(-1)]) + S // 2
    vox = np.stack([cols[:, 0], ys, cols[:, 1]], 1).astype(np.uint32)
    t0 = time.perf_counter()
    scene = vrt.LSVO.from_voxels(ctx, depth, vox, on_device=True)
    t_build = (time.perf_counter() - t0) * 1e3
    same = np.array_equal(scene.download_nodes().view(np.uint64), vrt.host_build_terrain_lsvo(depth).view(np.uint64))
    t0 = time.perf_counter()
    vrt.host_build_lsvo_from_voxels(depth, vox)
    t_host = (time.perf_counter() - t0) * 1e3
    rng = np.random.default_rng(0)
    times = []
    for k in range(14):                                     # the first edits size the scene's memory pool; steady state after that
        edit = rng.integers(0, S, (1000, 3)).astype(np.uint32)
        t0 = time.perf_counter()
        scene.set_cells(edit, k % 2 == 0)
        if k >= 4:
            times.append((time.perf_counter() - t0) * 1e3)
    print(json.dumps(dict(world="512^3 terrain as a voxel list", voxels=len(vox), slots=scene.n_nodes, device_build_ms=round(t_build, 2),
                          identical_to_terrain_builder=bool(same), host_flattener_ms=round(t_host, 1),
                          ms_per_1000_voxel_edit=round(float(np.median(times)), 3))), flush=True)
    scene.close()


if __name__ == "__main__":
    voxel_sets()
    main()
