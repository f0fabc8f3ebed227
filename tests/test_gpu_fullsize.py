"""BASELINE.json configurations at (or near) their full sizes: bit-exact against the oracle where the oracle finishes
in seconds, and size-independent properties (kernel variants agree byte for byte, ray-class bookkeeping, partitions)
where it does not."""
import numpy as np
import pytest

from conftest import assert_hits_equal, golden

pytestmark = pytest.mark.gpu


def hit_flag(h):
    return (h["flags"] & 1) != 0


def camera_rays(D, W, H, voxel_units=False):
    S = float(1 << D)
    cam = np.float32([S / 2, S / 2 - 56, S / 2])
    x, y = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    d = np.stack([x / np.float32(H) - np.float32(W / H * 0.5), y / np.float32(H) - np.float32(0.5), np.ones_like(x)], -1).reshape(-1, 3)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = cam if voxel_units else cam / np.float32(S) + np.float32(1)
    return np.broadcast_to(o, d.shape).astype(np.float32).copy(), d.astype(np.float32)


def test_cfg2_grid512_1080p_bit_exact(vrt, ctx, port):
    """configs[1]: dense Grid3D 512^3, coherent primary rays at 1920x1080, hit buffer bit-exact."""
    size = 512
    h = vrt.host_terrain_heights(size)
    hm = np.maximum(16, np.minimum(size, h))
    y = np.arange(size)[None, :, None]
    cells = ((y >= size // 2 + 1) & (y <= (size // 2 + hm - 1)[:, None, :])).astype(np.uint8)
    cells = np.ascontiguousarray(cells[::-1, ::-1, ::-1])          # the world the LSVO shows (point mirrored)
    o, d = camera_rays(9, 1920, 1080, voxel_units=True)
    got = vrt.Grid3D(ctx, cells).cast_rays(o, d)
    want, steps = port.grid_cast(cells, o, d, threads=16)
    assert_hits_equal(got, want, hit_flag(got), "cfg2")
    assert 0.3 < hit_flag(got).mean() < 0.7
    mip = vrt.MipmapGrid3D(ctx, cells, 4).cast_rays(o, d)
    assert np.array_equal(got.view(np.uint8), mip.view(np.uint8))


def test_cfg4_scene_2048_rays_and_variants(vrt, port):
    """configs[3] scene (T(11), built on the GPU): 1080p primary rays — K1 and K1p byte-identical on all 2 M rays,
    bit-exact against the oracle on a 200 k prefix (the oracle builds its own 1.35 GB node array)."""
    c = vrt.Context(0)
    s = vrt.LSVO.from_terrain(c, 11)
    o, d = camera_rays(11, 1920, 1080)
    c.set_option("cast_variant", 0)
    a = s.cast_rays(o, d)
    c.set_option("cast_variant", 1)
    b = s.cast_rays(o, d)
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    nodes = port.build_terrain(11)
    assert len(nodes) == len(s)
    sel = np.arange(0, len(o), len(o) // 200000)[:200000]
    want = port.lsvo_cast(nodes, 11, o[sel], d[sel], threads=16)
    assert_hits_equal(a[sel], want, hit_flag(a[sel]), "cfg4 primary")
    c.close()


def test_cfg4_frame_bookkeeping_and_partitions(vrt, ctx, textures):
    """1080p GI + DOF frame at 2048^3 (8 spp): ray-class counts obey the chain's structure, and the frame is the same
    whether it is rendered in one call, in sample batches, or in interleaved tile partitions."""
    s = vrt.LSVO.from_terrain(ctx, 11)
    s.set_textures(*textures)
    W, H, spp = 1920, 1080, 8
    S = 2048.0
    cam = vrt.Camera(position=(S / 2, S / 2 - 56, S / 2), view_angle=(0, 0), aperture=0.5)
    cam.autofocus(s)
    light = np.float32([-200, -1000, -300]) * np.float32(1 / S) + np.float32(1)

    def rc():
        r = vrt.RayCaster(s, (W, H))
        r.setLightPosition(light)
        r.use_samples, r.use_gi, r.gi_bounces = True, True, 2
        return r
    one = rc()
    one.render(cam, spp)
    rays = one.last_stats["rays"]
    assert rays[0] == W * H * spp                       # one primary ray per sample
    assert rays[1] == rays[2]                           # every primary hit casts one shadow and one GI ray
    # GI hits cast a shadow ray; the second bounce needs a non-zero GI normal (a few GI rays start inside a cell)
    assert rays[5] <= rays[4] <= rays[3] <= rays[2] and rays[4] > 0.99 * rays[3]
    assert int(one.colors[..., 3].min()) == spp == int(one.colors[..., 3].max())
    parts = rc()
    parts.render(cam, 3)
    parts.render(cam, 5)                                # sample batches continue the Philox stream
    assert np.array_equal(one.colors, parts.colors) and np.array_equal(one.render_image, parts.render_image)
    from cpuvoxelraycaster_b200 import capi
    import ctypes as C
    tiled = rc()
    for idx in range(3):                                # what 3 GPUs would each render
        p = tiled.params(spp, 0, H, 0)
        p.tile_step, p.tile_index, p.accum_in = 3, idx, 1
        st = capi.RenderStats()
        capi.check(capi.lib().vrt_render(s.handle, C.byref(cam.as_struct()), C.byref(p), capi.ptr(tiled.render_image),
                                         capi.ptr(tiled.colors), C.byref(st)))
    assert np.array_equal(one.colors, tiled.colors) and np.array_equal(one.render_image, tiled.render_image)
    s.close()


def test_cfg4_headline_frame_is_kernel_independent(vrt, textures):
    """The headline workload itself (1080p, 64 spp, GI 2 bounces, DOF, 2048^3): K4 (lane per pixel), K5 (samples sorted by
    GI direction inside a CTA) and K6 (sorted lists + persistent helping CTAs, the default) give the same accumulators,
    the same image and the same per-class ray and loop-trip counts — 451 M rays, compared exactly."""
    S, W, H, spp = 2048.0, 1920, 1080, 64
    light = np.float32([-200, -1000, -300]) * np.float32(1 / S) + np.float32(1)
    results = []
    for variant in (0, 2, 3):
        c = vrt.Context(0)
        c.set_option("render_variant", variant)
        c.set_option("beam_tile", 0)                     # K5 has no beam floors / bounds exits: compare the reference trip counts
        c.set_option("bounds_exit", 0)
        s = vrt.LSVO.from_terrain(c, 11)
        s.set_textures(*textures)
        cam = vrt.Camera(position=(S / 2, S / 2 - 56, S / 2), view_angle=(0, 0), aperture=0.5)
        cam.autofocus(s)
        r = vrt.RayCaster(s, (W, H))
        r.setLightPosition(light)
        r.use_samples, r.use_gi, r.gi_bounces = True, True, 2
        img = r.render(cam, spp).copy()
        results.append((r.colors.copy(), img, r.last_stats))
        s.close()
        c.close()
    ref = results[0]
    assert sum(ref[2]["rays"]) > 4.4e8 and int(ref[0][..., 3].min()) == spp == int(ref[0][..., 3].max())
    for other in results[1:]:
        assert np.array_equal(ref[0], other[0]) and np.array_equal(ref[1], other[1])
        assert ref[2]["rays"] == other[2]["rays"] and ref[2]["complexity"] == other[2]["complexity"]


def test_cfg5_lsvo4096_random_rays(vrt, port):
    """configs[4]: LSVO 4096^3 (built on the GPU, lsvo.hpp:72 guard lifted), incoherent random rays: the persistent
    regenerating kernel and the one-thread-per-ray kernel agree byte for byte on 4 M rays; a 100 k prefix is bit-exact
    against the oracle walking its own host-built 5.4 GB array."""
    c = vrt.Context(0)
    s = vrt.LSVO.from_terrain(c, 12, guard=-1)
    rng = np.random.default_rng(0xD1CE)
    n = 1 << 22
    o = rng.uniform([1, 1, 1], [2, 1.5 - 96 / 4096.0, 2], (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    c.set_option("cast_variant", 1)
    a = s.cast_rays(o, d)
    c.set_option("cast_variant", 0)
    b = s.cast_rays(o, d)
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    assert 0.15 < hit_flag(a).mean() < 0.4
    nodes = port.build_terrain(12)
    assert len(nodes) == len(s)
    want = port.lsvo_cast(nodes, 12, o[:100000], d[:100000], guard=0, threads=16)
    assert_hits_equal(a[:100000], want, hit_flag(a[:100000]), "cfg5 prefix")
    # with the reference's guard (scale > 12) nothing below level 10 is reachable: every ray misses
    ref_guard = vrt.LSVO.from_terrain(c, 12, guard=0)
    assert not hit_flag(ref_guard.cast_rays(o[:20000], d[:20000])).any()
    c.close()


# ---- round 2: the benchmark configurations pinned against the ORACLE at their own sizes (row crops where the oracle ----
# ---- would need minutes for the whole frame) --------------------------------------------------------------------------
def _cfg4_setup(vrt, ctx, textures):
    s = vrt.LSVO.from_terrain(ctx, 11)
    s.set_textures(*textures)
    S = 2048.0
    cam = vrt.Camera(position=(S / 2, S / 2 - 56, S / 2), view_angle=(0, 0), aperture=0.5)
    cam.autofocus(s)
    light = np.float32([-200, -1000, -300]) * np.float32(1 / S) + np.float32(1)
    return s, cam, light


def _port_cfg4_params(cam, light, W, H, spp, bounces, rows, threads=16):
    from oracle import loader
    pp = loader.PortRenderParams()
    pp.width, pp.height, pp.depth, pp.guard = W, H, 11, 11
    pp.cam_position[:] = [float(x) for x in cam.position]
    pp.rot_mat[:] = [float(x) for x in cam.rot_mat]
    pp.fov, pp.aperture, pp.focal_length = 1.0, cam.aperture, cam.focal_length
    pp.light_position[:] = [float(x) for x in light]
    pp.use_gi, pp.gi_bounces, pp.use_samples, pp.spp = 1, bounces, 1, spp
    pp.seed_lo, pp.seed_hi, pp.threads = 0x5EED, 0, threads
    pp.row_begin, pp.row_end = rows
    return pp


@pytest.mark.parametrize("rows", [(536, 568), (1000, 1032)])
def test_cfg4_headline_frame_crop_vs_oracle(vrt, ctx, port, textures, rows):
    """configs[3] as benchmarked — T(11) = 2048^3, 1920x1080, 64 spp, use_samples + GI (2 bounces) + DOF, camera C(11),
    the default frame kernels (K6) — on 32-row crops (the horizon band, and near terrain) against oracle/port.c:
    integer accumulators, 8-bit image and the per-class ray / loop-trip counts, all exact.  Pins the depth-dependent
    shading constants (SCALE, n_norm) at the benchmark depth; oracle/port.c itself is pinned at this depth against the
    reference's RayCaster rebuilt for depth 11 (tests/test_oracle_pinned.py)."""
    import ctypes as C
    from cpuvoxelraycaster_b200 import capi
    s, cam, light = _cfg4_setup(vrt, ctx, textures)
    W, H, spp = 1920, 1080, 64
    nodes = port.build_terrain(11)
    want_acc, want_img, want_st = port.render(nodes, _port_cfg4_params(cam, light, W, H, spp, 2, rows), *textures)
    r = vrt.RayCaster(s, (W, H))
    r.setLightPosition(light)
    r.use_samples, r.use_gi, r.gi_bounces = True, True, 2
    r.render(cam, spp, rows[0], rows[1])
    y0, y1 = rows
    assert np.array_equal(r.colors[y0:y1], want_acc[y0:y1])
    assert np.array_equal(r.render_image[y0:y1], want_img[y0:y1])
    assert not r.colors[:y0].any() and not r.colors[y1:].any()
    assert r.last_stats["rays"] == list(want_st.rays) and r.last_stats["complexity"] == list(want_st.complexity)
    assert r.last_stats["rays"][0] == W * (y1 - y0) * spp and r.last_stats["rays"][4] > 0
    # the product default: beam floors on — same pixels and ray counts, fewer loop trips on the primary rays only
    for tile, bounds in ((8, 0), (4, 0), (8, 1)):
        ctx.set_option("beam_tile", tile)
        ctx.set_option("bounds_exit", bounds)
        b = vrt.RayCaster(s, (W, H))
        b.setLightPosition(light)
        b.use_samples, b.use_gi, b.gi_bounces = True, True, 2
        b.render(cam, spp, rows[0], rows[1])
        assert np.array_equal(b.colors, r.colors) and np.array_equal(b.render_image, r.render_image)
        assert b.last_stats["rays"] == r.last_stats["rays"]
        if not bounds:
            assert b.last_stats["complexity"][1:] == r.last_stats["complexity"][1:]
        assert all(x <= y for x, y in zip(b.last_stats["complexity"], r.last_stats["complexity"]))
        assert b.last_stats["complexity"][0] < 0.8 * r.last_stats["complexity"][0]
    ctx.set_option("beam_tile", 0)
    ctx.set_option("bounds_exit", 0)
    s.close()


def test_cfg3_mipgrid1024_4k_frame_crop_vs_oracle(vrt, ctx, port, textures):
    """configs[2] at its size: MipmapGrid3D over the 1024^3 terrain with a Cell::Mirror lake, 3840x2160, primary + sun
    shadow + blurry reflections — a 24-row crop across the lake, u8- and accumulator-exact against the oracle's
    specification (vo_grid_render), and the cast records of the crop's primary rays bit-exact against vo_grid_cast."""
    size, W, H, mip = 1024, 3840, 2160, 4
    h = vrt.host_terrain_heights(size)
    hm = np.maximum(16, np.minimum(size, h))
    y = np.arange(size)[None, :, None]
    cells = ((y >= size // 2 + 1) & (y <= (size // 2 + hm - 1)[:, None, :])).astype(np.uint8)
    cells = np.ascontiguousarray(cells[::-1, ::-1, ::-1])
    hh = h[::-1, ::-1]
    top = size // 2 - np.maximum(16, np.minimum(size, hh))
    water = size // 2 - 30
    xs, zs = np.nonzero(top > water)
    for x, z in zip(xs.tolist(), zs.tolist()):
        cells[x, water + 1:top[x, z], z] = 1
    cells[xs, water, zs] = 2
    rows = (1800, 1824)                                             # terrain and lake: most reflection rays of the frame
    o, d = camera_rays(10, W, H, voxel_units=True)
    sel = slice(rows[0] * W, rows[1] * W)
    g = vrt.MipmapGrid3D(ctx, cells, mip)
    got = g.cast_rays(o[sel], d[sel])
    want, _ = port.grid_cast(cells, o[sel], d[sel], threads=16)
    assert_hits_equal(got, want, hit_flag(got), "cfg3 primary crop")
    g.set_textures(*textures)
    from oracle import loader
    cam = vrt.Camera(position=(size / 2, size / 2 - 56, size / 2), view_angle=(0.0, 0.0), focal_length=100.0)
    light = np.float32([-200, -1000, -300]) * np.float32(size / 512.0)
    pp = loader.PortRenderParams()
    pp.width, pp.height, pp.depth, pp.guard = W, H, 10, 10
    pp.cam_position[:] = [float(x) for x in cam.position]
    pp.rot_mat[:] = [float(x) for x in cam.rot_mat]
    pp.fov, pp.aperture, pp.focal_length = 1.0, 0.0, 100.0
    pp.light_position[:] = [float(x) for x in light]
    pp.use_gi, pp.gi_bounces, pp.use_samples, pp.spp = 0, 1, 1, 2
    pp.seed_lo, pp.seed_hi, pp.threads = 0x5EED, 0, 16
    pp.row_begin, pp.row_end = rows
    pp.roughness, pp.max_bounds = 0.06, 4
    want_acc, want_img, want_st = port.grid_render(cells, pp, *textures)
    r = vrt.RayCaster(g, (W, H))
    r.setLightPosition(light)
    r.use_samples, r.roughness, r.max_bounds = True, 0.06, 4
    r.render(cam, 2, rows[0], rows[1])
    assert np.array_equal(r.colors[rows[0]:rows[1]], want_acc[rows[0]:rows[1]])
    assert np.array_equal(r.render_image[rows[0]:rows[1]], want_img[rows[0]:rows[1]])
    assert r.last_stats["rays"][:3] == list(want_st.rays)[:3] and want_st.rays[2] > 0     # reflections happened in the crop
    g.close()


def test_cfg5_lsvo4096_prefix_1e6_vs_oracle(vrt, port):
    """configs[4]: the first 10^6 of the benchmark's random rays at 4096^3, bit-exact against the oracle,
    for both cast kernels."""
    c = vrt.Context(0)
    s = vrt.LSVO.from_terrain(c, 12, guard=-1)
    rng = np.random.default_rng(0xD1CE)
    n = 1_000_000
    o = rng.uniform([1, 1, 1], [2, 1.5 - 96 / 4096.0, 2], (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    nodes = port.build_terrain(12)
    want = port.lsvo_cast(nodes, 12, o, d, guard=0, threads=16)
    del nodes
    for variant in (3, 1, 0, 2):                                       # 3 = automatic (the default): the classifier sends these to K1p
        c.set_option("cast_variant", variant)
        a = s.cast_rays(o, d)
        assert_hits_equal(a, want, hit_flag(a), "cfg5 10^6 prefix, cast_variant %d" % variant)
    c.close()


def test_cfg1_full_frame_known_answers(vrt, ctx, textures):
    """configs[0] at 1280x720 against the values frozen from the reference's own RayCaster (tests/golden/cfg1_full.json):
    468 025 primary hits; image, accumulator, hit-flag, complexity and distance hashes."""
    import hashlib
    import json
    import os
    from conftest import GOLDEN
    g = json.load(open(os.path.join(GOLDEN, "cfg1_full.json")))
    W, H = g["width"], g["height"]
    s = vrt.LSVO(ctx, vrt.host_build_terrain_lsvo(9), 9)
    s.set_textures(*textures)
    cam = vrt.Camera(position=g["cam_position"], view_angle=g["view_angle"], focal_length=g["focal_length"])
    for variant in (0, 2, 4):
        ctx.set_option("render_variant", variant)
        r = vrt.RayCaster(s, (W, H))
        r.setLightPosition(np.float32(g["light"]))
        r.use_samples = True
        r.render(cam, 1)
        assert hashlib.sha256(r.render_image.tobytes()).hexdigest() == g["sha256_image"], variant
        assert hashlib.sha256(r.colors.tobytes()).hexdigest() == g["sha256_samples_u32"], variant
        assert r.last_stats["rays"][:2] == [W * H, g["primary_hits"]] and r.last_stats["complexity"][0] == g["sum_complexity"]
    ctx.set_option("render_variant", 0)
    # the same camera rays through the batched cast: Camera::getRay (camera_controller.hpp:34-49) in numpy, glm's operation order;
    # view (0,0) gives the identity rot_mat
    f32 = np.float32

    def normalize(v):                                              # glm::normalize: v * (1 / sqrt(dot))
        inv = f32(1.0) / np.sqrt((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2], dtype=np.float32)
        return (v * inv[:, None]).astype(np.float32)
    x, y = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    lens = np.stack([x / f32(H) - f32(f32(W) / f32(H) * f32(0.5)), y / f32(H) - f32(0.5), np.ones_like(x)], -1).reshape(-1, 3).astype(np.float32)
    d = normalize((normalize(lens) * f32(g["focal_length"])).astype(np.float32))
    assert np.array_equal(cam.rot_mat, np.eye(3, dtype=np.float32).ravel())
    o = np.broadcast_to(np.float32(g["cam_position"]) * np.float32(1 / 512.0) + np.float32(1), d.shape).copy()
    for variant in (0, 1, 2, 3):                                       # 3 = automatic (the default): the classifier sends these to K1b
        ctx.set_option("cast_variant", variant)
        hits = s.cast_rays(o, d)
        hf = hit_flag(hits)
        assert int(hf.sum()) == g["primary_hits"]
        assert hashlib.sha256(hf.astype(np.uint8).tobytes()).hexdigest() == g["sha256_hit_flags"]
        assert hashlib.sha256(hits["complexity"].astype(np.uint32).tobytes()).hexdigest() == g["sha256_complexity"]
        assert hashlib.sha256(hits["distance"][hf].astype(np.float32).tobytes()).hexdigest() == g["sha256_distance_of_hits"]
    ctx.set_option("cast_variant", 3)
    s.close()
