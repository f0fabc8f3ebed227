"""Host logic of the product (scene construction, camera basis) against the golden vectors and the oracle."""
import hashlib

import numpy as np

from conftest import golden


def test_terrain_heights(vrt):
    g = golden("terrain_heights.npz")
    assert np.array_equal(vrt.host_terrain_heights(256), g["heights8"])
    assert hashlib.sha256(vrt.host_terrain_heights(512).tobytes()).hexdigest() == str(g["sha_heights9"])
    assert hashlib.sha256(vrt.host_terrain_heights(1024).tobytes()).hexdigest() == str(g["sha_heights10"])


def test_terrain_flattening(vrt, terrain9_nodes):
    g = golden("terrain_heights.npz")
    assert len(terrain9_nodes) == 10528393
    assert hashlib.sha256(terrain9_nodes.tobytes()).hexdigest() == str(g["sha_nodes9"])


def test_terrain_flattening_depth10_matches_oracle(vrt, port):
    a = vrt.host_build_terrain_lsvo(10)
    b = port.build_terrain(10)
    assert len(a) == 42436609
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))


def test_voxel_list_flattening(vrt):
    g = golden("lsvo_random6.npz")
    nodes = vrt.host_build_lsvo_from_voxels(6, g["voxels"])
    assert np.array_equal(nodes.view(np.uint64), g["nodes"].view(np.uint64))
    k = golden("lsvo_kat.npz")
    nodes = vrt.host_build_lsvo_from_voxels(9, k["voxels"])
    ref_nodes = k["nodes"].copy()
    ref_nodes["pad"] = 0
    assert len(nodes) == 73 and np.array_equal(nodes.view(np.uint64), ref_nodes.view(np.uint64))


def test_voxel_list_edge_cases(vrt, port):
    empty = vrt.host_build_lsvo_from_voxels(4, np.zeros((0, 3), np.uint32))
    assert len(empty) == 1 and empty["child_offset"][0] == 1 and empty["child_mask"][0] == 0
    full = np.stack(np.meshgrid(*[np.arange(8)] * 3, indexing="ij"), -1).reshape(-1, 3)
    a = vrt.host_build_lsvo_from_voxels(3, np.concatenate([full, full]))          # duplicates collapse
    b = port.build_dense(3, np.ones((8, 8, 8), np.uint8))
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))


def test_camera_basis(vrt):
    g = golden("frame_cfg1_small.npz")
    cam = vrt.Camera(view_angle=g["view_angle"])
    assert np.array_equal(cam.rot_mat.view(np.uint32), g["rot_mat"].view(np.uint32))
    assert np.array_equal(cam.camera_vec.view(np.uint32), g["camera_vec"].view(np.uint32))


def test_replay_file_and_controllers(vrt, tmp_path):
    """include/replay.hpp:18-33 (six floats per tick, reading stops at the first bad token or incomplete tick) and the
    camera controllers (camera_controller.hpp:64-78, fly_controller.hpp)."""
    f = tmp_path / "replay.txt"
    f.write_text("0.0 256 200 256 0.0 0.0\n0.016 257.5 199 256 0.1 -0.2\n0.033 259 198 256.5 0.2\t-0.4\n0.05 1 2 oops 4 5\n")
    ticks = vrt.ReplayElements.loadFromFile(str(f))
    assert len(ticks) == 3
    assert (ticks[1].timestamp, ticks[1].x, ticks[1].view_y) == (float(np.float32(0.016)), 257.5, float(np.float32(-0.2)))
    assert vrt.ReplayElements.loadFromFile(str(tmp_path / "missing.txt")) == []
    cam = vrt.Camera()
    ticks[2].apply(cam)
    assert np.array_equal(cam.position, np.float32([259, 198, 256.5])) and np.array_equal(cam.view_angle, np.float32([0.2, -0.4]))
    ctl = vrt.FlyController()
    ctl.updateCameraView((0.5, -3.0), cam)                 # pitch clamps at -PI/2 with the reference's PI literal
    assert cam.view_angle[0] == np.float32(0.2) + np.float32(0.5)
    assert cam.view_angle[1] == -(np.float32(3.141592653) * np.float32(0.5))
    ctl.move((1.0, -2.0, 0.5), cam)
    assert np.array_equal(cam.position, np.float32([260, 196, 257]))
