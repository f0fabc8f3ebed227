"""The CPU restatement (oracle/port.c) against the committed golden vectors, which were produced by the
reference's own compiled sources (tests/golden/make_golden.py).  Bit-exact."""
import hashlib

import numpy as np

from conftest import assert_hits_equal, golden


def test_philox_known_answers(port):
    # Random123 kat_vectors, philox4x32 10 rounds
    assert list(port.philox([0, 0, 0, 0], [0, 0])) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert list(port.philox([0xffffffff] * 4, [0xffffffff] * 2)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert list(port.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])) == [
        0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_terrain_heights_match_fastnoise(port):
    g = golden("terrain_heights.npz")
    assert np.array_equal(port.terrain_heights(256), g["heights8"])
    assert hashlib.sha256(port.terrain_heights(512).tobytes()).hexdigest() == str(g["sha_heights9"])
    assert hashlib.sha256(port.terrain_heights(1024).tobytes()).hexdigest() == str(g["sha_heights10"])


def test_terrain_flattening_matches_compileSVO(port):
    g = golden("terrain_heights.npz")
    nodes = port.build_terrain(9)
    assert len(nodes) == int(g["n_nodes9"]) == 10528393          # SURVEY.md §8c known answer
    assert hashlib.sha256(nodes.tobytes()).hexdigest() == str(g["sha_nodes9"])
    assert int(g["solid_voxels9"]) == 8583552


def test_single_voxel_known_answers(port):
    g = golden("lsvo_kat.npz")
    occ = np.zeros((512, 512, 512), np.uint8)
    occ[100, 200, 300] = 1
    nodes = port.build_dense(9, occ)
    assert len(nodes) == 73
    ref_nodes = g["nodes"].copy()
    ref_nodes["pad"] = 0
    assert np.array_equal(nodes.view(np.uint64), ref_nodes.view(np.uint64))
    hits = port.lsvo_cast(nodes, 9, g["origin"], g["dir"])
    assert_hits_equal(hits, g["hits"], hits["hit"] != 0, "kat")
    # point mirroring: setCell(100,200,300) is seen at (411,311,211)
    assert list(hits["voxel"][0]) == [411, 311, 211] and hits["distance"][1] == 0.0
    assert hits["hit"][3] == 0 and hits["hit"][4] == 0           # t > 1 is a miss; un-mirrored column misses


def test_lsvo_terrain_rays(port):
    g = golden("lsvo_terrain9.npz")
    nodes = port.build_terrain(9)
    for key, coef, bias in (("hits_coef0", 0.0, 0.0), ("hits_coef05", 0.5, 0.0), ("hits_bias", 0.25, 0.001)):
        hits = port.lsvo_cast(nodes, 9, g["origin"], g["dir"], coef, bias, threads=4)
        assert_hits_equal(hits, g[key], hits["hit"] != 0, key)
    assert 0.2 < (g["hits_coef0"]["hit"] != 0).mean() < 0.9


def test_lsvo_random_scene(port):
    g = golden("lsvo_random6.npz")
    occ = np.zeros((64, 64, 64), np.uint8)
    v = g["voxels"]
    occ[v[:, 0], v[:, 1], v[:, 2]] = 1
    nodes = port.build_dense(6, occ)
    assert np.array_equal(nodes.view(np.uint64), g["nodes"].view(np.uint64))
    for key, coef in (("hits_coef0", 0.0), ("hits_coef05", 0.5)):
        hits = port.lsvo_cast(nodes, 6, g["origin"], g["dir"], coef, 0.0)
        assert_hits_equal(hits, g[key], hits["hit"] != 0, key)


def test_grid_dda(port):
    g = golden("grid_random5.npz")
    hits, steps = port.grid_cast(g["occ"], g["origin"], g["dir"])
    assert_hits_equal(hits, g["hits"], hits["hit"] != 0, "grid")
    assert np.all(steps[hits["hit"] != 0] == hits["complexity"][hits["hit"] != 0])


def test_svo_intended(port):
    g = golden("svo_random5.npz")
    hits = port.svo_cast(g["occ"], 5, g["origin"], g["dir"], 1 << 20)
    assert_hits_equal(hits, g["hits"], hits["hit"] != 0, "svo")
    hits = port.svo_cast(g["occ"], 5, g["origin"], g["dir"], 16)
    assert_hits_equal(hits, g["hits_iter16"], hits["hit"] != 0, "svo max_iter=16")


def test_frame_primary_shadow(port, textures):
    """Deterministic RayCaster frame (primary + sun shadow): u8-exact against the reference's RayCaster."""
    from oracle import loader
    g = golden("frame_cfg1_small.npz")
    p = loader.PortRenderParams()
    p.width, p.height, p.depth, p.guard = int(g["width"]), int(g["height"]), 9, 9
    p.cam_position[:] = [float(x) for x in g["cam_position"]]
    p.rot_mat[:] = [float(x) for x in g["rot_mat"]]
    p.fov, p.aperture, p.focal_length = 1.0, 0.0, 100.0
    p.light_position[:] = [float(x) for x in g["light"]]
    p.use_gi, p.gi_bounces, p.use_samples, p.spp = 0, 1, 1, 1
    p.seed_lo, p.seed_hi, p.threads = 0x5EED, 0, 4
    nodes = port.build_terrain(9)
    accum, rgba, stats = port.render(nodes, p, *textures)
    assert np.array_equal(accum, g["samples"])
    assert np.array_equal(rgba, g["image"])
    assert stats.rays[0] == p.width * p.height and 0 < stats.rays[1] < stats.rays[0]


def test_checkerboard_frames(port, vrt, textures):
    """The interactive loop's checkerboard halves (main.cpp:137-143) with the temporal blend, 4 frames, and two half frames in
    sample mode: the oracle against frames produced by the reference's own RayCaster (golden)."""
    from oracle import loader
    g = golden("frame_checker_small.npz")
    nodes = port.build_terrain(9)
    cam = vrt.Camera(position=g["cam_position"], view_angle=g["view_angle"], focal_length=100.0)   # host-side basis only
    for tag in ("a", "b"):
        W, H, area = (int(v) for v in g["size_" + tag])
        p = loader.PortRenderParams()
        p.width, p.height, p.depth, p.guard = W, H, 9, 9
        p.cam_position[:] = [float(x) for x in g["cam_position"]]
        p.rot_mat[:] = [float(x) for x in cam.rot_mat]
        p.fov, p.aperture, p.focal_length = 1.0, 0.0, 100.0
        p.light_position[:] = [float(x) for x in g["light"]]
        p.use_gi, p.gi_bounces, p.use_samples, p.spp = 0, 1, 0, 1
        p.seed_lo, p.threads, p.checker_area_height = 0x5EED, 4, area
        img = None
        for frame in range(4):
            p.checker = 1 + ((1 + frame) & 1)
            _, img, _ = port.render(nodes, p, *textures, prev_rgba=img)
        assert np.array_equal(img, g["blend4_" + tag]), tag
        p.use_samples = 1
        acc = np.zeros((H, W, 4), np.uint32)
        for frame in range(2):
            p.checker = 1 + (frame & 1)
            acc += port.render(nodes, p, *textures)[0]
        assert np.array_equal(acc, g["samples2_" + tag]), tag
