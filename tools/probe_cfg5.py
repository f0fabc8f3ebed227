"""cfg 5 (LSVO 4096^3, incoherent random rays — the one memory-latency-visible workload): A/B of the node-memory design
points north_star (1) names, all with identical hit records (checked by hash):
  layout 0 = the reference's LNode array (8 slots of 8 B per node), layout 1 = compact breadth-first array of live nodes
  (8 B per node), + an L2 access-policy window over its front; kernels K1p (persistent, regenerating), K1 and K1b (one
  thread per ray).  The variant with the top octree levels staged in shared memory (option "smem_top_nodes", removed again)
  was measured with this script at commit 71c5298: profiles/r02_probe_cfg5.txt, profiles/r02_summary.md.
One JSON line per variant.  PROBE_RAYS overrides the ray count (default 1e8)."""
import hashlib
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpuvoxelraycaster_b200 as vrt  # noqa: E402


def main():
    n = int(float(os.environ.get("PROBE_RAYS", "1e8")))
    only = os.environ.get("PROBE_ONLY")                  # "layout,l2,top,variant": run one variant (for ncu)
    stream = torch.cuda.Stream()
    ctx = vrt.Context(0, stream.cuda_stream)
    scene = vrt.LSVO.from_terrain(ctx, 12, guard=-1)
    g = torch.Generator(device="cuda").manual_seed(0xD1CE)
    o = torch.rand(n, 3, device="cuda", generator=g)
    o[:, 0] += 1.0
    o[:, 2] += 1.0
    o[:, 1] = 1.0 + o[:, 1] * (0.5 - 96.0 / 4096.0)
    d = torch.randn(n, 3, device="cuda", generator=g)
    d /= d.norm(dim=1, keepdim=True)
    out = torch.empty(n * 16, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    variants = [(0, 0, 0, 1), (1, 0, 0, 1), (1, 1, 0, 1), (0, 0, 0, 0), (0, 0, 0, 2), (1, 0, 0, 0)]
    if only:
        variants = [tuple(int(x) for x in only.split(","))]
    ref = None
    for layout, l2, top, variant in variants:
        scene.set_layout(layout, l2)
        ctx.set_option("cast_variant", variant)
        reps = 1 if only else 3
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        with torch.cuda.stream(stream):
            if not only:
                scene.cast_rays_device(o, d, n, out)
            for i in range(reps):
                ev[i].record(stream)
                scene.cast_rays_device(o, d, n, out)
            ev[reps].record(stream)
        stream.synchronize()
        ms = float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]))
        h = hashlib.sha256(out[: 16 * min(n, 4_000_000)].cpu().numpy().tobytes()).hexdigest()[:16]
        ref = ref or h
        print(json.dumps(dict(layout=layout, l2_window=l2, smem_top_nodes=top, cast_variant=variant, ms=round(ms, 3), grays_s=round(n / ms / 1e6, 3),
                              same_records=h == ref, trips=scene.last_complexity())), flush=True)
    ctx.set_option("cast_variant", 1)


if __name__ == "__main__":
    main()
