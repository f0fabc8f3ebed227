"""Wall clock of consecutive 1000-voxel edits (vrt_scene_set_cells) of the T(9) terrain held as a voxel set; run under
`ncu --metrics gpu__time_duration.sum` for the kernel list of an edit."""
import sys, time; sys.path.insert(0, ".")
import numpy as np, cpuvoxelraycaster_b200 as vrt
ctx = vrt.Context(0)
S, depth = 512, 9
h = vrt.host_terrain_heights(S)
hmax = np.maximum(16, np.minimum(S, h)).astype(np.int64)
xs, zs = np.meshgrid(np.arange(S), np.arange(S), indexing="ij")
cols = np.repeat(np.stack([xs.reshape(-1), zs.reshape(-1)], 1), hmax.reshape(-1) - 1, axis=0)
ys = np.concatenate([np.arange(1, m) for m in hmax.reshape(-1)]) + S // 2
vox = np.stack([cols[:, 0], ys, cols[:, 1]], 1).astype(np.uint32)
scene = vrt.LSVO.from_voxels(ctx, depth, vox, on_device=True)
rng = np.random.default_rng(0)
for k in range(6):
    edit = rng.integers(0, S, (1000, 3)).astype(np.uint32)
    t0 = time.perf_counter(); scene.set_cells(edit, k % 2 == 0); print("edit ms", (time.perf_counter() - t0) * 1e3, flush=True)
