"""The C++ drop-in headers (include/vrt/vrt.hpp): compile a miniature of the reference's main.cpp against them
(CPU box: compile + link + fail loudly without a device), and on the GPU check it against the Python API."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden

EXE = os.path.join(ROOT, "tests", "cpp", "dropin_test")
TEX = os.path.join(ROOT, "tests", "cpp", "textures.bin")


def build():
    lib_dir = os.path.join(ROOT, "cpuvoxelraycaster_b200")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "dropin_test.cpp"), "-o", EXE, "-L" + lib_dir, "-lvrt",
                    "-Wl,-rpath," + lib_dir], check=True)
    t = golden("textures.npz")
    open(TEX, "wb").write(t["top"].tobytes() + t["side"].tobytes())


def test_dropin_compiles_and_fails_loudly_without_gpu(vrt):
    build()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([EXE, TEX], capture_output=True, text=True)
    assert r.returncode == 3 and "vrt::Error -2" in r.stdout      # VRT_ERR_CUDA, no silent CPU path


@pytest.mark.gpu
def test_dropin_matches_python_api(vrt, ctx, terrain9_nodes, textures):
    build()
    r = subprocess.run([EXE, TEX], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    out = {}
    for line in r.stdout.strip().splitlines():
        key, _, rest = line.partition(" ")
        if "=" in key:                                   # "autofocus=100.000000"
            key, _, rest = line.partition("=")
        out[key] = rest
    assert out["kat"].startswith("nodes=73 hit=1 complexity=14 distance=0.41210938 normal=(-0,-0,-4)")
    assert out["batch"] == "hits=101"
    assert out["terrain"] == "nodes=10528393"
    assert "same=1" in out["grid"] and "hit=1" in out["grid"]
    # the same frame through the Python mirror
    s = vrt.LSVO(ctx, terrain9_nodes, 9)
    s.set_textures(*textures)
    cam = vrt.Camera(position=(256, 200, 256), view_angle=(0.3, -0.35), aperture=0.5, focal_length=60.0)
    rc = vrt.RayCaster(s, (256, 144))
    rc.setLightPosition(np.float32([-200, -1000, -300]) * np.float32(1.0 / 512) + np.float32(1.0))
    rc.use_samples, rc.use_gi = True, True
    rc.render(cam, spp=3)
    h = 1469598103934665603
    for b in rc.render_image.tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    assert out["frame"].startswith("hash=%016x samples=3" % h)
    # the interactive loop (checkerboard + temporal blend + median/persistence presentation)
    live = vrt.RayCaster(s, (256, 144))
    live.setLightPosition(rc.light_position)
    live.checker_area_height = 36
    cam0 = vrt.Camera(position=(256, 200, 256), view_angle=(0.3, -0.35), aperture=0.0, focal_length=60.0)
    for frame in range(3):
        live.checker_board_offset = 1 - (frame & 1)
        live.render(cam0)
        live.present(median=3)

    def fnv(a):
        h = 1469598103934665603
        for b in a.tobytes():
            h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return h
    assert out["live"] == "hash=%016x display=%016x" % (fnv(live.render_image), fnv(live.display))
    af = vrt.Camera(position=(256, 200, 256), view_angle=(0.3, -0.35)).autofocus(s)
    assert out["autofocus"] == "%.6f" % af
