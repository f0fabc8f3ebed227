"""Replay-driven fly-through of the demo's interactive loop, device resident:

  per tick of the replay file (include/replay.hpp format):  camera ← tick;  focal length ← centre-ray autofocus
  (main.cpp:114-121);  flip the checkerboard offset (:137);  render that half of the pixels, 1 sample, sun shadow + GI
  (:139-152);  0.4/0.6 temporal blend (raycaster.hpp:79-85);  persistence blend into the display surface (:159-172).

Prints one JSON line per configuration: ms per frame (CUDA events around the whole loop, autofocus read-back included),
frames/s and Mrays/s.  The reference runs this loop at 960x540 on 16 threads (main.cpp:30-32,89)."""
import argparse
import json
import math
import os
import sys
import tempfile

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpuvoxelraycaster_b200 as vrt  # noqa: E402
from cpuvoxelraycaster_b200.frame import FrameRenderer  # noqa: E402


def write_replay(path, depth, ticks):
    """A synthetic recording: one lap around the map centre, looking along the path and down at the hills."""
    S = float(1 << depth)
    with open(path, "w") as f:
        for i in range(ticks):
            a = 2.0 * math.pi * i / ticks
            x, z = S / 2 + 0.3 * S * math.cos(a), S / 2 + 0.3 * S * math.sin(a)
            y = S / 2 - 110.0 - 20.0 * math.sin(3 * a)        # hill tops reach S/2 - 81
            f.write("%.6f %.4f %.4f %.4f %.6f %.6f\n" % (i / 60.0, x, y, z, a + math.pi / 2, -0.45 + 0.15 * math.sin(2 * a)))


def fly(scene, W, H, ticks, use_gi, checker, median, save_png=None, host_focus=False):
    fr = FrameRenderer(scene, W, H)
    fr.autofocus = not host_focus                                          # centre-ray focus on the device
    fr.use_samples, fr.use_gi = False, use_gi
    fr.light = np.float32([-200, -1000, -300]) * np.float32(1.0 / (1 << scene.depth)) + np.float32(1.0)
    fr.checker_area_height = H // 4 if checker else 0                      # main.cpp:91,132: 4x4 thread areas
    cam = vrt.Camera(aperture=0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rays = 0
    for timed in (False, True):
        offset = 0
        if timed:
            ev0.record(fr.stream)
        for k, tick in enumerate(ticks if timed else ticks[:8]):
            tick.apply(cam)
            if host_focus:
                cam.autofocus(scene)                                       # host round trip per frame
            offset = 1 - offset
            fr.checker_board_offset = offset if checker else None
            frame = fr.render_device(cam, 1)
            display = fr.present_device(frame, median, 0.1)
            if timed and k % 16 == 0:
                rays += sum(fr.stats()["rays"]) * 16
        if timed:
            ev1.record(fr.stream)
    fr.stream.synchronize()
    ms = ev0.elapsed_time(ev1) / len(ticks)
    if save_png:
        from render_gallery import save
        save(save_png, display.cpu().numpy())
    return dict(width=W, height=H, ticks=len(ticks), gi=bool(use_gi), checkerboard=bool(checker), median=median,
                autofocus="host" if host_focus else "device",
                ms_per_frame=round(ms, 4), frames_per_s=round(1000.0 / ms, 1), mrays_s=round(rays / len(ticks) / ms / 1e3, 1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depth", type=int, default=9)
    ap.add_argument("--ticks", type=int, default=240)
    ap.add_argument("--png", default=None)
    a = ap.parse_args()
    ctx = vrt.Context(0)
    scene = vrt.LSVO.from_terrain(ctx, a.depth)
    here = os.path.dirname(os.path.abspath(__file__))
    tex = np.load(os.path.join(here, "..", "tests", "golden", "textures.npz"))
    scene.set_textures(tex["top"], tex["side"])
    path = os.path.join(tempfile.mkdtemp(), "replay.txt")
    write_replay(path, a.depth, a.ticks)
    ticks = vrt.ReplayElements.loadFromFile(path)
    assert len(ticks) == a.ticks
    for (W, H, gi, checker, median, host_focus) in ((960, 540, True, True, 0, True), (960, 540, True, True, 0, False),
                                                    (960, 540, True, False, 0, False), (1920, 1080, True, True, 0, False),
                                                    (1920, 1080, True, True, 3, False), (3840, 2160, True, True, 0, False)):
        png = a.png if (a.png and (W, checker, median) == (1920, True, 0)) else None
        print(json.dumps(dict(depth=a.depth, **fly(scene, W, H, ticks, gi, checker, median, png, host_focus))), flush=True)


if __name__ == "__main__":
    main()
