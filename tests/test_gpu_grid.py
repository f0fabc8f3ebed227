"""K2 / K2m / K3 parity: Grid3D, MipmapGrid3D and (intended) SVO casts on the B200 through the C ABI against the
golden vectors produced by the patched reference build and against the oracle on seeded inputs.  Bit-exact."""
import numpy as np
import pytest

from conftest import assert_hits_equal, golden

pytestmark = pytest.mark.gpu


def hit_flag(h):
    return (h["flags"] & 1) != 0


def test_grid_golden(vrt, ctx):
    g = golden("grid_random5.npz")
    s = vrt.Grid3D(ctx, g["occ"])
    hits = s.cast_rays(g["origin"], g["dir"])
    assert_hits_equal(hits, g["hits"], hit_flag(hits), "grid")
    h = s.castRay(g["origin"][0], g["dir"][0])
    assert h.cell == bool(g["hits"]["hit"][0])


@pytest.mark.parametrize("levels", [1, 2, 3, 5])
def test_mipmap_grid_equals_flat_grid(vrt, ctx, levels):
    g = golden("grid_random5.npz")
    m = vrt.MipmapGrid3D(ctx, g["occ"], mip_levels=levels)
    hits = m.cast_rays(g["origin"], g["dir"])
    assert_hits_equal(hits, g["hits"], hit_flag(hits), "mip grid, %d levels" % levels)


def terrain_grid(vrt, size):
    """T(D) as a dense grid in the reference orientation (main.cpp:63-76 with setCell on the grid)."""
    h = vrt.host_terrain_heights(size)
    hm = np.maximum(16, np.minimum(size, h))
    y = np.arange(size)[None, :, None]
    top = (size // 2 + hm - 1)[:, None, :]
    return ((y >= size // 2 + 1) & (y <= top)).astype(np.uint8)


def test_grid_terrain_vs_oracle(vrt, ctx, port):
    size = 256
    cells = terrain_grid(vrt, size)
    rng = np.random.default_rng(5)
    n = 200000
    o = rng.uniform(0, size, (n, 3)).astype(np.float32)
    o[:, 1] = rng.uniform(0, size // 2, n)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[:100, 0] = 0.0                       # 1/0 = inf lanes of the DDA, integer origins give 0/0 = NaN
    o[:50] = np.floor(o[:50])
    want, steps = port.grid_cast(cells, o, d, threads=8)
    flat = vrt.Grid3D(ctx, cells)
    got = flat.cast_rays(o, d)
    assert_hits_equal(got, want, hit_flag(got), "grid terrain")
    assert flat.last_complexity() == int(steps.sum())          # loop iterations, hits and misses
    m = hit_flag(got)
    assert np.array_equal(got["voxel"][m], want["voxel"][m])
    mip = vrt.MipmapGrid3D(ctx, cells, mip_levels=4)
    got2 = mip.cast_rays(o, d)
    assert np.array_equal(got.view(np.uint8), got2.view(np.uint8))          # identical records, byte for byte
    assert 0.05 < m.mean() < 0.95


def test_grid_edge_cases(vrt, ctx, port):
    cells = np.zeros((5, 7, 3), np.uint8)                        # non-cubic, non power of two
    cells[2, 3, 1] = 1
    cells[4, 6, 2] = 2                                           # Cell::Mirror is non-empty too
    rng = np.random.default_rng(1)
    o = rng.uniform(-1, 8, (5000, 3)).astype(np.float32)         # many origins outside the grid: loop never runs
    d = rng.normal(size=(5000, 3)).astype(np.float32)
    want, _ = port.grid_cast(cells, o, d)
    for scene in (vrt.Grid3D(ctx, cells), vrt.MipmapGrid3D(ctx, cells, mip_levels=2)):
        got = scene.cast_rays(o, d)
        assert_hits_equal(got, want, hit_flag(got), "small grid")
        assert len(scene.cast_rays(np.zeros((0, 3)), np.zeros((0, 3)))) == 0
    empty = vrt.Grid3D(ctx, np.zeros((4, 4, 4), np.uint8))
    assert not hit_flag(empty.cast_rays([[0.5, 0.5, 0.5]], [[0.1, 0.2, 1]]))[0]


def test_svo_golden(vrt, ctx):
    g = golden("svo_random5.npz")
    s = vrt.SVO(ctx, g["occ"])
    hits = s.cast_rays(g["origin"], g["dir"], 1 << 20)
    assert_hits_equal(hits, g["hits"], hit_flag(hits), "svo")
    hits = s.cast_rays(g["origin"], g["dir"], 16)
    assert_hits_equal(hits, g["hits_iter16"], hit_flag(hits), "svo max_iter=16")


def test_svo_terrain_vs_oracle(vrt, ctx, port):
    size = 128
    h = vrt.host_terrain_heights(256)[:size, :size]
    y = np.arange(size)[None, :, None]
    cells = ((y >= 10) & (y <= (10 + np.clip(h, 1, 100))[:, None, :])).astype(np.uint8)
    rng = np.random.default_rng(9)
    n = 100000
    o = rng.uniform(0, size, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    want = port.svo_cast(cells, 7, o, d, 4096, threads=8)
    got = vrt.SVO(ctx, cells).cast_rays(o, d, 4096)
    assert_hits_equal(got, want, hit_flag(got), "svo terrain")
    m = hit_flag(got)
    assert np.array_equal(got["voxel"][m], want["voxel"][m]) and m.mean() > 0.3


def mirror_scene(vrt, size=128):
    """Terrain columns over a flat lake of Cell::Mirror voxels (types: 0 empty, 1 solid, 2 mirror)."""
    h = vrt.host_terrain_heights(256)[:size, :size]
    top = 30 + np.clip(h, 0, 70)
    y = np.arange(size)[None, :, None]
    cells = (y <= top[:, None, :]).astype(np.uint8)
    lake = h <= 18
    xs, zs = np.nonzero(lake)
    cells[xs, 30 + np.clip(h, 0, 70)[lake], zs] = 2
    return cells


@pytest.mark.parametrize("mip,roughness,aperture,spp", [(0, 0.0, 0.0, 1), (3, 0.15, 0.0, 2), (3, 0.3, 0.4, 3)])
def test_grid_frame_with_reflections_matches_oracle(vrt, ctx, port, textures, mip, roughness, aperture, spp):
    from oracle import loader
    cells = mirror_scene(vrt)
    assert (cells == 2).sum() > 100
    scene = vrt.MipmapGrid3D(ctx, cells, mip) if mip else vrt.Grid3D(ctx, cells)
    scene.set_textures(*textures)
    W, H = 200, 120
    cam = vrt.Camera(position=(64.0, 118.0, 20.0), view_angle=(0.0, 0.75), aperture=aperture, focal_length=40.0)
    light = np.float32([300.0, 900.0, -200.0])
    rc = vrt.RayCaster(scene, (W, H))
    rc.setLightPosition(light)
    rc.use_samples, rc.roughness, rc.max_bounds = True, roughness, 4
    img = rc.render(cam, spp=spp)
    p = loader.PortRenderParams()
    p.width, p.height, p.depth, p.guard = W, H, 7, 7
    p.cam_position[:] = [float(x) for x in cam.position]
    p.rot_mat[:] = [float(x) for x in cam.rot_mat]
    p.fov, p.aperture, p.focal_length = cam.fov, cam.aperture, cam.focal_length
    p.light_position[:] = [float(x) for x in light]
    p.use_gi, p.gi_bounces, p.use_samples, p.spp = 0, 1, 1, spp
    p.seed_lo, p.seed_hi, p.threads = 0x5EED, 0, 8
    p.roughness, p.max_bounds = roughness, 4
    accum, rgba, stats = port.grid_render(cells, p, *textures)
    assert np.array_equal(rc.colors, accum) and np.array_equal(img, rgba)
    assert rc.last_stats["rays"][:3] == list(stats.rays)[:3] and rc.last_stats["complexity"][:3] == list(stats.complexity)[:3]
    assert stats.rays[2] > 0 and img[..., :3].max() > 0, "the view must contain reflections and lit terrain"


def test_grid_frame_rejects_gi(vrt, ctx, textures):
    scene = vrt.Grid3D(ctx, np.ones((8, 8, 8), np.uint8))
    scene.set_textures(*textures)
    rc = vrt.RayCaster(scene, (16, 16))
    rc.use_samples, rc.use_gi = True, True
    with pytest.raises(vrt.VrtError):
        rc.render(vrt.Camera(position=(4, 20, 4)), spp=1)


def test_bordered_grid_dda_equals_the_generic_loops(vrt, ctx, port):
    """The default DDA (one linear bit index on the bordered grid, no bounds tests: vrt_context_set_option grid_variant 0)
    against the generic loops (grid_variant 1: flat, and with the fetch-skipping pyramid) and the oracle — random scene,
    origins inside and outside, axis-parallel directions; and a grid long enough for the 2048-iteration cap of
    grid_3d.hpp:70 to bind (a solid cell behind 2100 empty ones is NOT hit)."""
    rng = np.random.default_rng(9)
    cells = (rng.random((40, 24, 56)) < 0.01).astype(np.uint8)
    n = 60000
    o = rng.uniform(-3, 60, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d[:300, 2] = 0.0
    d[300:600, 0] = -0.0
    o[:600] = np.floor(o[:600])
    want, steps = port.grid_cast(cells, o, d, threads=8)
    for make in (lambda: vrt.Grid3D(ctx, cells), lambda: vrt.MipmapGrid3D(ctx, cells, mip_levels=3)):
        scene = make()
        records = []
        for variant in (0, 1):
            ctx.set_option("grid_variant", variant)
            got = scene.cast_rays(o, d)
            assert_hits_equal(got, want, hit_flag(got), "grid_variant %d" % variant)
            assert scene.last_complexity() == int(steps.sum())
            records.append(got.view(np.uint8).copy())
        assert np.array_equal(records[0], records[1])
        scene.close()
    ctx.set_option("grid_variant", 0)
    long_grid = np.zeros((4, 4, 2200), np.uint8)
    long_grid[1, 1, 2100] = 1
    long_grid[2, 2, 2000] = 1
    oo = np.float32([[1.5, 1.5, 0.5], [2.5, 2.5, 0.5], [1.5, 1.5, 100.5]])
    dd = np.float32([[0, 0, 1], [0, 0, 1], [0, 0, 1]])
    want, steps = port.grid_cast(long_grid, oo, dd)
    assert list(want["hit"]) == [0, 1, 1] and int(steps[0]) == 2048
    for variant in (0, 1):
        ctx.set_option("grid_variant", variant)
        s = vrt.Grid3D(ctx, long_grid)
        got = s.cast_rays(oo, dd)
        assert_hits_equal(got, want, hit_flag(got), "cap, grid_variant %d" % variant)
        assert s.last_complexity() == int(steps.sum())
        s.close()
    ctx.set_option("grid_variant", 0)
