"""K0+K4 parity: frames rendered on the B200 through the C ABI against (a) the golden frame produced by the
reference's own RayCaster and (b) the oracle restatement with the same Philox lattice RNG.  u8-exact."""
import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu


def port_params(W, H, depth, cam, light, use_gi, bounces, use_samples, spp, seed=(0x5EED, 0), offset=0, rows=(0, 0)):
    from oracle import loader
    p = loader.PortRenderParams()
    p.width, p.height, p.depth, p.guard = W, H, depth, depth
    p.cam_position[:] = [float(x) for x in cam.position]
    p.rot_mat[:] = [float(x) for x in cam.rot_mat]
    p.fov, p.aperture, p.focal_length = cam.fov, cam.aperture, cam.focal_length
    p.light_position[:] = [float(x) for x in light]
    p.use_gi, p.gi_bounces, p.use_samples, p.spp = int(use_gi), bounces, int(use_samples), spp
    p.seed_lo, p.seed_hi, p.sample_offset = seed[0], seed[1], offset
    p.row_begin, p.row_end, p.threads = rows[0], rows[1], 8
    return p


@pytest.fixture(scope="module")
def scene9(vrt, ctx, terrain9_nodes, textures):
    s = vrt.LSVO(ctx, terrain9_nodes, 9)
    s.set_textures(*textures)
    return s


def default_light(depth=9):
    return np.float32([-200, -1000, -300]) * np.float32(1.0 / (1 << depth)) + np.float32(1.0)


def test_golden_frame_primary_shadow(vrt, scene9):
    g = golden("frame_cfg1_small.npz")
    cam = vrt.Camera(position=g["cam_position"], view_angle=g["view_angle"], focal_length=100.0)
    rc = vrt.RayCaster(scene9, (int(g["width"]), int(g["height"])))
    rc.setLightPosition(g["light"])
    rc.use_samples = True
    img = rc.render(cam, spp=1)
    assert np.array_equal(rc.colors, g["samples"])
    assert np.array_equal(img, g["image"])
    assert rc.last_stats["rays"][0] == img.shape[0] * img.shape[1]


@pytest.mark.parametrize("use_gi,bounces,aperture,spp", [(0, 1, 0.0, 1), (1, 1, 0.0, 2), (1, 2, 0.5, 3), (0, 1, 0.5, 2)])
def test_frame_matches_oracle(vrt, scene9, port, terrain9_nodes, textures, use_gi, bounces, aperture, spp):
    W, H = 256, 144
    cam = vrt.Camera(position=(256, 200, 256), view_angle=(0.3, -0.35), aperture=aperture, focal_length=60.0)
    rc = vrt.RayCaster(scene9, (W, H))
    rc.setLightPosition(default_light())
    rc.use_samples, rc.use_gi, rc.gi_bounces = True, bool(use_gi), bounces
    img = rc.render(cam, spp=spp)
    p = port_params(W, H, 9, cam, default_light(), use_gi, bounces, True, spp)
    accum, rgba, stats = port.render(terrain9_nodes, p, *textures)
    assert np.array_equal(rc.colors, accum)
    assert np.array_equal(img, rgba)
    assert rc.last_stats["rays"] == list(stats.rays)
    assert rc.last_stats["complexity"] == list(stats.complexity)
    assert img[..., :3].max() > 0


def test_progressive_and_row_split_are_exact(vrt, scene9):
    """spp batches and row slabs (the multi-GPU partitions) reproduce the single-call frame bit for bit."""
    W, H = 192, 108
    cam = vrt.Camera(position=(256, 200, 256), view_angle=(0.3, -0.35), aperture=0.5, focal_length=60.0)

    def make():
        rc = vrt.RayCaster(scene9, (W, H))
        rc.setLightPosition(default_light())
        rc.use_samples, rc.use_gi, rc.gi_bounces = True, True, 2
        return rc
    one = make()
    one.render(cam, spp=4)
    prog = make()
    prog.render(cam, spp=1)
    prog.render(cam, spp=3)
    assert np.array_equal(one.colors, prog.colors) and np.array_equal(one.render_image, prog.render_image)
    slab = make()
    for r0, r1 in ((0, 27), (27, 54), (54, 81), (81, 108)):
        slab.sample_count = 0
        slab.render(cam, spp=4, row_begin=r0, row_end=r1)
    assert np.array_equal(one.colors, slab.colors) and np.array_equal(one.render_image, slab.render_image)


def test_temporal_blend_mode(vrt, scene9, port, terrain9_nodes, textures):
    W, H = 160, 90
    cam = vrt.Camera(position=(256, 200, 256), view_angle=(0.3, -0.35), focal_length=100.0)
    rc = vrt.RayCaster(scene9, (W, H))
    rc.setLightPosition(default_light())
    rc.use_samples = False
    prev = None
    for frame in range(3):
        img = rc.render(cam, spp=1).copy()
        p = port_params(W, H, 9, cam, default_light(), 0, 1, False, 1)
        _, want, _ = port.render(terrain9_nodes, p, *textures, prev_rgba=prev)
        assert np.array_equal(img, want), "frame %d" % frame
        prev = want


def test_golden_checkerboard_frames(vrt, scene9):
    """K4's checkerboard mapping + resolve against frames the reference's own RayCaster produced when driven like
    main.cpp:137-143 (4 alternating half frames with the 0.4/0.6 blend; two half frames in sample mode)."""
    g = golden("frame_checker_small.npz")
    cam = vrt.Camera(position=g["cam_position"], view_angle=g["view_angle"], focal_length=100.0)
    for tag in ("a", "b"):
        W, H, area = (int(v) for v in g["size_" + tag])
        rc = vrt.RayCaster(scene9, (W, H))
        rc.setLightPosition(g["light"])
        rc.checker_area_height = area
        for frame in range(4):
            rc.checker_board_offset = (1 + frame) & 1
            img = rc.render(cam)
        assert np.array_equal(img, g["blend4_" + tag]), tag
        rs = vrt.RayCaster(scene9, (W, H))
        rs.setLightPosition(g["light"])
        rs.checker_area_height, rs.use_samples = area, True
        for frame in range(2):
            rs.checker_board_offset = frame & 1
            rs.render(cam)
        assert np.array_equal(rs.colors, g["samples2_" + tag]), tag


@pytest.mark.parametrize("W,H,area_height,use_samples", [(160, 90, 0, False), (161, 75, 15, False), (96, 60, 15, True), (70, 44, 11, True)])
def test_checkerboard_frames(vrt, scene9, port, terrain9_nodes, textures, W, H, area_height, use_samples):
    """main.cpp:137-143: alternating checkerboard halves; the unrendered pixels keep their value."""
    cam = vrt.Camera(position=(256, 200, 256), view_angle=(0.3, -0.35), focal_length=80.0, aperture=0.4)
    rc = vrt.RayCaster(scene9, (W, H))
    rc.setLightPosition(default_light())
    rc.use_samples, rc.use_gi = use_samples, True
    rc.checker_area_height = area_height
    prev, acc = None, np.zeros((H, W, 4), np.uint32)
    for frame in range(4):
        rc.checker_board_offset = 1 - (frame & 1)
        img = rc.render(cam, spp=2 if use_samples else 1).copy()
        p = port_params(W, H, 9, cam, default_light(), 1, 1, use_samples, 2 if use_samples else 1, offset=2 * frame if use_samples else frame)   # blend mode: frame k draws sample stream k
        p.checker, p.checker_area_height = 1 + (1 - (frame & 1)), area_height
        a, want, _ = port.render(terrain9_nodes, p, *textures, prev_rgba=prev)
        if use_samples:
            acc += a
            assert np.array_equal(rc.colors, acc), "frame %d" % frame
            rendered = acc[..., 3] > 0
            assert np.array_equal(img[rendered][:, :3], (acc[rendered][:, :3] // acc[rendered][:, 3:4]).astype(np.uint8))
        else:
            assert np.array_equal(img, want), "frame %d" % frame
            prev = want
    n = rc.last_stats["rays"][0]
    assert abs(n - (2 if use_samples else 1) * W * H / 2) <= (2 if use_samples else 1) * (H + W)


@pytest.mark.parametrize("W,H", [(160, 90), (333, 77), (31, 9)])
def test_present_matches_oracle(vrt, ctx, port, W, H):
    """vrt_present (median + persistence blend, main.cpp:159-177) byte-exact against the oracle."""
    import ctypes as C
    from cpuvoxelraycaster_b200 import capi
    rng = np.random.default_rng(W)
    frame = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    for median in (0, 3, 5):
        for ovc in (0.1, 0.0, 0.7):
            display = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
            want = port.present(frame, display, median, ovc)
            got = display.copy()
            p = capi.PresentParams(W, H, median, ovc)
            capi.check(capi.lib().vrt_present(ctx.handle, capi.ptr(frame), capi.ptr(got), C.byref(p)))
            assert np.array_equal(got, want), (median, ovc)
    bad = capi.PresentParams(W, H, 4, 0.1)
    with pytest.raises(capi.VrtError):
        capi.check(capi.lib().vrt_present(ctx.handle, capi.ptr(frame), capi.ptr(got), C.byref(bad)))


def test_camera_inside_solid_terminates(vrt, scene9, port, terrain9_nodes, textures):
    """Autofocus from inside a hill gives focal_length 0 (main.cpp:116-118); with an open aperture Camera::getRay then
    normalises a zero vector whenever both lens numbers are 0 — a NaN ray the reference would never finish.  Such
    samples are black here and in the oracle, and the frame still matches."""
    W, H = 128, 72
    cam = vrt.Camera(position=(307.2727, 210.45, 400.7897), view_angle=(2.801253, -0.355602), aperture=0.3)
    assert cam.autofocus(scene9) == 0.0
    rc = vrt.RayCaster(scene9, (W, H))
    rc.setLightPosition(default_light())
    rc.use_samples, rc.use_gi = True, True
    rc.render(cam, spp=4)
    p = port_params(W, H, 9, cam, default_light(), 1, 1, True, 4)
    acc, _, st = port.render(terrain9_nodes, p, *textures)
    assert np.array_equal(rc.colors, acc) and rc.last_stats["rays"] == list(st.rays)
    assert rc.last_stats["complexity"][0] < 40 * W * H * 4


@pytest.mark.parametrize("view", [(0.0, -0.6), (0.3, -0.35), (0.0, 0.9)])
def test_device_side_autofocus(vrt, scene9, view):
    """vrt_render_params::autofocus: the frame equals the one rendered with Camera::autofocus's focal length."""
    W, H = 128, 72
    cam = vrt.Camera(position=(256, 200, 256), view_angle=view, aperture=0.5, focal_length=3.0)
    rc = vrt.RayCaster(scene9, (W, H))
    rc.setLightPosition(default_light())
    rc.use_samples, rc.autofocus = True, True
    a = rc.render(cam, spp=3).copy()
    cam.autofocus(scene9)
    assert cam.focal_length != 3.0
    rc2 = vrt.RayCaster(scene9, (W, H))
    rc2.setLightPosition(default_light())
    rc2.use_samples = True
    b = rc2.render(cam, spp=3)
    assert np.array_equal(a, b) and np.array_equal(rc.colors, rc2.colors)


@pytest.mark.parametrize("W,H,spp,chunks,bounces,variant", [(150, 70, 16, 0, 2, 0), (150, 70, 16, 2, 1, 0), (97, 41, 64, 0, 2, 0),
                                                            (64, 36, 5, 0, 2, 3), (64, 36, 70, 0, 2, 3), (33, 9, 24, 3, 1, 3),
                                                            (64, 36, 12, 0, 0, 3), (150, 70, 16, 0, 2, 4), (64, 36, 5, 0, 2, 4),
                                                            (64, 36, 70, 0, 2, 4), (33, 9, 24, 0, 1, 4), (203, 77, 12, 0, 0, 4)])
def test_direction_sorted_kernel_is_exact(vrt, port, terrain9_nodes, textures, W, H, spp, chunks, bounces, variant):
    """K5 regroups the samples of a 32x4 pixel block by GI direction before tracing them: integer sums, same frame,
    same ray statistics.  Ragged frames, uneven runs, more than 64 samples per pixel (forced extra runs), no GI.
    Variant 4 = K6: the sorted lists in global memory, persistent CTAs that help each other with unfinished blocks."""
    c = vrt.Context(0)
    c.set_option("render_variant", variant)
    c.set_option("spp_chunks", chunks)
    s = vrt.LSVO(c, terrain9_nodes, 9)
    s.set_textures(*textures)
    cam = vrt.Camera(position=(256, 200, 256), view_angle=(0.3, -0.35), aperture=0.5, focal_length=60.0)
    rc = vrt.RayCaster(s, (W, H))
    rc.setLightPosition(default_light())
    rc.use_samples, rc.use_gi, rc.gi_bounces = True, bounces > 0, max(1, bounces)
    img = rc.render(cam, spp=spp).copy()
    accum, rgba, stats = port.render(terrain9_nodes, port_params(W, H, 9, cam, default_light(), int(bounces > 0), max(1, bounces), True, spp), *textures)
    assert np.array_equal(rc.colors, accum) and np.array_equal(img, rgba)
    assert rc.last_stats["rays"] == list(stats.rays) and rc.last_stats["complexity"] == list(stats.complexity)
    rc.render(cam, spp=8)                               # progressive: 8 more samples on top
    p2 = port_params(W, H, 9, cam, default_light(), int(bounces > 0), max(1, bounces), True, 8, offset=spp)
    accum2, _, _ = port.render(terrain9_nodes, p2, *textures)
    assert np.array_equal(rc.colors, accum + accum2)
    c.close()


def test_autofocus(vrt, scene9, port, terrain9_nodes):
    cam = vrt.Camera(position=(256, 200, 256), view_angle=(0.0, -0.6))
    f = cam.autofocus(scene9)
    o = cam.position * np.float32(1 / 512.0) + np.float32(1)
    want = port.lsvo_cast(terrain9_nodes, 9, [o], [cam.camera_vec])[0]
    assert want["hit"] and f == np.float32(want["distance"]) * np.float32(512.0)
    assert vrt.Camera(position=(256, 200, 256), view_angle=(0.0, 0.0)).autofocus(scene9) == 100.0   # centre ray misses


@pytest.mark.parametrize("variant,refill", [(0, 8), (1, 1), (1, 12), (1, 32), (2, 8), (3, 8)])
def test_render_kernel_variants_agree(vrt, port, terrain9_nodes, textures, variant, refill):
    c = vrt.Context(0)
    c.set_option("render_variant", variant)
    c.set_option("refill_render", refill)
    s = vrt.LSVO(c, terrain9_nodes, 9)
    s.set_textures(*textures)
    W, H = 203, 77                                   # ragged: not a multiple of the 8x4 tile
    cam = vrt.Camera(position=(256, 200, 256), view_angle=(0.3, -0.35), aperture=0.5, focal_length=60.0)
    rc = vrt.RayCaster(s, (W, H))
    rc.setLightPosition(default_light())
    rc.use_samples, rc.use_gi, rc.gi_bounces = True, True, 2
    img = rc.render(cam, spp=3)
    p = port_params(W, H, 9, cam, default_light(), 1, 2, True, 3)
    accum, rgba, stats = port.render(terrain9_nodes, p, *textures)
    assert np.array_equal(rc.colors, accum) and np.array_equal(img, rgba)
    assert rc.last_stats["rays"] == list(stats.rays) and rc.last_stats["complexity"] == list(stats.complexity)
    c.close()


@pytest.mark.parametrize("chunks", [1, 2, 3, 7])
def test_sample_chunking_is_exact(vrt, port, terrain9_nodes, textures, chunks):
    """K4 may cut a pixel's samples into runs handled by different CTAs (atomic integer sums): same frame."""
    c = vrt.Context(0)
    c.set_option("spp_chunks", chunks)
    s = vrt.LSVO(c, terrain9_nodes, 9)
    s.set_textures(*textures)
    W, H, spp = 150, 70, 7
    cam = vrt.Camera(position=(256, 200, 256), view_angle=(0.3, -0.35), aperture=0.5, focal_length=60.0)
    rc = vrt.RayCaster(s, (W, H))
    rc.setLightPosition(default_light())
    rc.use_samples, rc.use_gi, rc.gi_bounces = True, True, 2
    img = rc.render(cam, spp=spp)
    accum, rgba, stats = port.render(terrain9_nodes, port_params(W, H, 9, cam, default_light(), 1, 2, True, spp), *textures)
    assert np.array_equal(rc.colors, accum) and np.array_equal(img, rgba)
    assert rc.last_stats["rays"] == list(stats.rays)
    c.close()


@pytest.mark.parametrize("q,chunks,spp", [(1, 1, 8), (2, 1, 8), (8, 1, 8), (4, 2, 8), (32, 1, 32), (16, 2, 32), (0, 0, 6)])
def test_lane_mapping_is_exact(vrt, port, terrain9_nodes, textures, q, chunks, spp):
    """K4 may give several lanes of a warp the same pixel (consecutive samples) — integer sums, same frame."""
    c = vrt.Context(0)
    c.set_option("render_variant", 2)
    c.set_option("samples_per_warp", q)
    c.set_option("spp_chunks", chunks)
    s = vrt.LSVO(c, terrain9_nodes, 9)
    s.set_textures(*textures)
    W, H = 70, 42
    cam = vrt.Camera(position=(256, 200, 256), view_angle=(0.3, -0.35), aperture=0.5, focal_length=60.0)
    rc = vrt.RayCaster(s, (W, H))
    rc.setLightPosition(default_light())
    rc.use_samples, rc.use_gi, rc.gi_bounces = True, True, 2
    img = rc.render(cam, spp=spp)
    accum, rgba, stats = port.render(terrain9_nodes, port_params(W, H, 9, cam, default_light(), 1, 2, True, spp), *textures)
    assert np.array_equal(rc.colors, accum) and np.array_equal(img, rgba)
    assert rc.last_stats["rays"] == list(stats.rays) and rc.last_stats["complexity"] == list(stats.complexity)
    c.close()


@pytest.mark.parametrize("spp,use_gi,bounces,roughness", [(2, 1, 2, 0.05), (16, 1, 2, 0.05), (16, 0, 1, 0.0), (3, 1, 1, 0.2)])
def test_lsvo_mirror_reflections_match_oracle(vrt, scene9, port, terrain9_nodes, textures, spp, use_gi, bounces, roughness):
    """Blurry mirror reflections on LSVO frames (extension specified in oracle/port.c shade_sample: the top faces of voxel layer
    y = 240 — the flat valley floors of T(9) — are Cell::Mirror): accumulators, image and ray / loop-trip counts exact against
    the oracle, through K4 (few samples) and K6 (>= 8 samples), with GI and depth of field on top."""
    W, H = 192, 108
    cam = vrt.Camera(position=(256, 200, 256), view_angle=(0.0, -0.6), aperture=0.3, focal_length=80.0)
    rc = vrt.RayCaster(scene9, (W, H))
    rc.setLightPosition(default_light())
    rc.use_samples, rc.use_gi, rc.gi_bounces = True, bool(use_gi), bounces
    rc.mirror_y, rc.roughness, rc.max_bounds = 240, roughness, 4
    img = rc.render(cam, spp=spp)
    p = port_params(W, H, 9, cam, default_light(), use_gi, bounces, True, spp)
    p.mirror_y1, p.roughness, p.max_bounds = 241, roughness, 4
    accum, rgba, stats = port.render(terrain9_nodes, p, *textures)
    assert np.array_equal(rc.colors, accum)
    assert np.array_equal(img, rgba)
    assert rc.last_stats["rays"] == list(stats.rays) and rc.last_stats["complexity"] == list(stats.complexity)
    assert stats.rays[0] > 1.05 * W * H * spp                    # reflection rays were cast (they count as class 0)
    # without the rule the same call is the plain frame
    plain = vrt.RayCaster(scene9, (W, H))
    plain.setLightPosition(default_light())
    plain.use_samples, plain.use_gi, plain.gi_bounces = True, bool(use_gi), bounces
    plain.render(cam, spp=spp)
    assert plain.last_stats["rays"][0] == W * H * spp and not np.array_equal(plain.render_image, img)


def _frame(vrt, scene, size, cam, light, spp, use_gi=True, bounces=2, mirror_y=None):
    rc = vrt.RayCaster(scene, size)
    rc.setLightPosition(light)
    rc.use_samples, rc.use_gi, rc.gi_bounces = True, use_gi, bounces
    rc.mirror_y, rc.roughness = mirror_y, 0.05
    rc.render(cam, spp=spp)
    return rc.colors.copy(), rc.last_stats


@pytest.mark.parametrize("position,view,aperture,focal", [((256, 200, 256), (0.3, -0.35), 0.5, 60.0), ((256, 200, 256), (0.0, 0.0), 0.0, 100.0),
                                                          ((256, 200, 256), (0.0, -0.6), 2.0, 30.0), ((300, 230, 120), (2.1, -0.1), 0.5, 200.0),
                                                          ((256, 262, 256), (0.5, 0.2), 0.5, 80.0), ((256, 200, 256), (0.0, 1.2), 1.0, 5.0)])
def test_beam_floors_do_not_change_frames(vrt, ctx, scene9, position, view, aperture, focal):
    """The per-tile start distances of the camera rays (beam_kernels.cu; on by default in the product, off in this test
    context) must be conservative for every pixel and every lens sample: accumulators, ray counts and the trip counts of all
    secondary rays are identical with tiles of 4, 8, 16 and 32 pixels and without; only the primary rays' trip count shrinks.
    Cameras: the demo's, grazing, looking down / up, wide aperture with a short focal length, close to the ground, under it."""
    cam = vrt.Camera(position=position, view_angle=view, aperture=aperture, focal_length=focal)
    for size, spp, variant in (((200, 113), 8, 2), ((128, 72), 16, 0)):          # K4 (forced) and K6
        ctx.set_option("render_variant", variant)
        ctx.set_option("beam_tile", 0)
        want, st0 = _frame(vrt, scene9, size, cam, default_light(), spp, mirror_y=240)
        assert st0["culled_primary"] == 0
        for tile in (4, 8, 16, 32):
            ctx.set_option("beam_tile", tile)
            got, st = _frame(vrt, scene9, size, cam, default_light(), spp, mirror_y=240)
            assert np.array_equal(got, want), (size, tile)
            assert st["rays"] == st0["rays"] and st["complexity"][1:] == st0["complexity"][1:]
            assert st["complexity"][0] <= st0["complexity"][0]
            # K6 does not start the chains of samples whose tile has an empty frustum: they are counted (rays[0] above, the
            # accumulator's sample counts in `got`), never more of them than camera rays that miss, and only K6 does it
            assert st["culled_primary"] <= st0["rays"][0] and (variant == 0 or st["culled_primary"] == 0)
        # ... and with the walks ending at the scene's bounds instead of the root cube (the product default: both on)
        for tile in (0, 8):
            ctx.set_option("beam_tile", tile)
            ctx.set_option("bounds_exit", 1)
            got, st = _frame(vrt, scene9, size, cam, default_light(), spp, mirror_y=240)
            ctx.set_option("bounds_exit", 0)
            assert np.array_equal(got, want), (size, tile, "bounds")
            assert st["rays"] == st0["rays"] and all(a <= b for a, b in zip(st["complexity"], st0["complexity"]))
            assert sum(st["complexity"]) <= sum(st0["complexity"])
        ctx.set_option("beam_tile", 0)
    ctx.set_option("render_variant", 0)


def test_samples_of_empty_beam_tiles_are_answered_without_a_walk(vrt, ctx, scene9):
    """K6 with beam floors: samples of pixels whose tile has nothing in its frustum do not enter the sample lists — the sort kernel
    counts them.  Horizon view (half sky): the frame (sums AND per-pixel sample counts), the ray counts of every class are those of
    the frame without floors; the culled rays are camera rays that miss (never more than rays[0] - rays[1], no mirrors here so
    every primary hit casts one sun-shadow ray), most of the sky is culled, and their trips are gone from complexity[0]."""
    cam = vrt.Camera(position=(256, 200, 256), view_angle=(0.0, 0.0), aperture=0.5, focal_length=100.0)
    size, spp = (256, 144), 16
    ctx.set_option("render_variant", 0)
    ctx.set_option("beam_tile", 0)
    want, st0 = _frame(vrt, scene9, size, cam, default_light(), spp)
    misses = st0["rays"][0] - st0["rays"][1]
    assert st0["rays"][0] == size[0] * size[1] * spp and misses > st0["rays"][0] // 4 and st0["culled_primary"] == 0
    try:
        for tile in (4, 8, 16):
            ctx.set_option("beam_tile", tile)
            got, st = _frame(vrt, scene9, size, cam, default_light(), spp)
            assert np.array_equal(got, want), tile
            assert (got[..., 3] == spp).all()
            assert st["rays"] == st0["rays"] and st["complexity"][1:] == st0["complexity"][1:]
            assert 0 < st["culled_primary"] <= misses, (tile, st["culled_primary"], misses)
            assert st["culled_primary"] % spp == 0                       # whole pixels
            if tile <= 8:
                assert st["culled_primary"] > misses // 2, (tile, st["culled_primary"], misses)
            assert st["complexity"][0] < st0["complexity"][0]
    finally:
        ctx.set_option("beam_tile", 0)


def test_beam_floors_on_a_random_voxel_scene(vrt, ctx, textures):
    """Thin, scattered geometry (6000 random voxels at depth 6, and a depth-9 scene of isolated voxels and 1-voxel columns): the
    case in which a corner-ray beam optimisation steps over geometry.  Frames identical with and without the floors."""
    g = golden("lsvo_random6.npz")
    rng = np.random.default_rng(3)
    sparse = np.concatenate([rng.integers(0, 512, (3000, 3)), np.stack([np.full(400, 100), np.arange(400), np.full(400, 300)], 1),
                             np.stack([np.arange(60, 460), np.full(400, 250), np.full(400, 257)], 1)]).astype(np.uint32)
    for depth, nodes, cams in ((6, g["nodes"], [((32, 32, -60), (0.0, 0.0)), ((32, 30, -20), (0.2, 0.1)), ((10, 20, 5), (0.7, 0.3))]),
                               (9, vrt.host_build_lsvo_from_voxels(9, sparse), [((256, 256, -100), (0.0, 0.0)), ((20, 20, 20), (0.7, 0.5))])):
        s = vrt.LSVO(ctx, nodes, depth)
        s.set_textures(*textures)
        light = np.float32([-200, -1000, -300]) * np.float32(1.0 / (1 << depth)) + np.float32(1.0)
        seen = 0
        for position, view in cams:
            cam = vrt.Camera(position=position, view_angle=view, aperture=0.7, focal_length=40.0)
            for size, spp, variant in (((160, 90), 8, 2), ((96, 54), 16, 0)):
                ctx.set_option("render_variant", variant)
                ctx.set_option("beam_tile", 0)
                want, st0 = _frame(vrt, s, size, cam, light, spp)
                seen += st0["rays"][1]
                for tile in (4, 8, 16):
                    ctx.set_option("beam_tile", tile)
                    got, st = _frame(vrt, s, size, cam, light, spp)
                    assert np.array_equal(got, want), (depth, position, size, tile)
                    assert st["rays"] == st0["rays"] and st["complexity"][1:] == st0["complexity"][1:]
                ctx.set_option("bounds_exit", 1)
                got, st = _frame(vrt, s, size, cam, light, spp)
                ctx.set_option("bounds_exit", 0)
                assert np.array_equal(got, want) and st["rays"] == st0["rays"], (depth, position, size, "bounds")
                assert all(a <= b for a, b in zip(st["complexity"], st0["complexity"]))
        assert seen > 1000, depth                              # the cameras do look at the voxels
        ctx.set_option("beam_tile", 0)
        ctx.set_option("render_variant", 0)
        s.close()


@pytest.mark.parametrize("aperture,focal,view", [(0.0, 100.0, (0.3, -0.35)), (0.5, 60.0, (0.3, -0.35)), (2.0, 20.0, (0.0, -0.1))])
def test_beam_floors_are_below_every_primary_hit(vrt, ctx, scene9, port, terrain9_nodes, aperture, focal, view):
    """The invariant itself: for every tile, floor <= the hit distance of every camera ray of the tile — checked against the
    ORACLE's primary hits for all pixels and 20 lens samples each — and the floors are
    not trivial: on average they skip more than a third of the way to the nearest hit of the tile."""
    W, H = 160, 90
    cam = vrt.Camera(position=(256, 200, 256), view_angle=view, aperture=aperture, focal_length=focal)
    rc = vrt.RayCaster(scene9, (W, H))
    rc.setLightPosition(default_light())
    p = port_params(W, H, 9, cam, default_light(), 0, 1, True, 1)
    # the oracle's camera rays, sample by sample (Philox lattice)
    o, d = [], []
    for y in range(H):
        for x in range(W):
            for smp in range(20):
                oo, dd = port.camera_ray(p, x, y, smp)
                o.append(oo)
                d.append(dd)
    o, d = np.array(o, np.float32), np.array(d, np.float32)
    hits = port.lsvo_cast(terrain9_nodes, 9, o, d, threads=8)
    t_hit = np.where(hits["hit"] != 0, hits["distance"], np.float32(10.0)).reshape(H, W, 20).min(-1)
    for tile in (4, 8, 16):
        fl = rc.beam_floors(cam, tile)
        per_pixel = np.repeat(np.repeat(fl, tile, 0), tile, 1)[:H, :W]
        assert (per_pixel <= t_hit).all(), (tile, float((per_pixel - t_hit).max()))
        tight = per_pixel[t_hit < 10] / t_hit[t_hit < 10]
        assert tight.mean() > 0.35, (tile, float(tight.mean()))


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_tile_split_slices_add_up_with_product_defaults(vrt, ctx, scene9, world):
    """What the ranks of a multi-GPU tile split render, rendered one after the other on ONE GPU with the product defaults (beam
    floors, bounds exit, K6): the slices' accumulators and ray counts add up to the single-GPU frame's.  Ragged frame (330x187:
    the last 4-row tile and the last 8-pixel beam tile are cut), 8 samples per pixel — the configuration
    tests/multigpu_frame_check.py runs on 4 GPUs."""
    import torch
    from cpuvoxelraycaster_b200.frame import FrameRenderer
    W, H, spp = 330, 187, 8
    cam = vrt.Camera(position=(256, 200, 256), view_angle=(0.3, -0.35), aperture=0.5, focal_length=60.0)
    cs = cam.as_struct()
    ctx.set_option("beam_tile", 8)
    ctx.set_option("bounds_exit", 1)
    try:
        def accumulate(rank, n):
            fr = FrameRenderer(scene9, W, H, rank, n, None, None, None, exchange="nccl")
            fr.use_gi, fr.gi_bounces, fr.light = True, 2, default_light()
            fr.accum.zero_()
            fr.accumulate(cs, fr.params(spp))
            torch.cuda.synchronize()
            return fr.accum.cpu().numpy().astype(np.int64)[: H * W * 4], fr.stats()
        want, st1 = accumulate(0, 1)
        total = np.zeros_like(want)
        rays = [0] * 6
        for r in range(world):
            a, st = accumulate(r, world)
            total += a
            rays = [x + y for x, y in zip(rays, st["rays"])]
        bad = np.flatnonzero((total != want).reshape(-1, 4).any(axis=1))
        assert bad.size == 0, "world %d: %d pixels differ, first at (x, y) = %s" % (world, bad.size, [(int(i % W), int(i // W)) for i in bad[:8]])
        assert rays == list(st1["rays"])
    finally:
        ctx.set_option("beam_tile", 0)
        ctx.set_option("bounds_exit", 0)
