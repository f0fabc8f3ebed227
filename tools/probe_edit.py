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


if __name__ == "__main__":
    main()
