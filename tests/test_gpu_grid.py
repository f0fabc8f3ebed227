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
