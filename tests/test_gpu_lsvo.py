"""K1 parity: LSVO<D>::castRay on the B200 through the C ABI, bit-exact against the reference's golden
vectors and against the oracle on seeded inputs (hit flag, position, normal, uv, distance, complexity)."""
import numpy as np
import pytest

from conftest import assert_hits_equal, bits, golden

pytestmark = pytest.mark.gpu


def hit_flag(h):
    return (h["flags"] & 1) != 0


def test_single_voxel_known_answers(vrt, ctx):
    g = golden("lsvo_kat.npz")
    s = vrt.LSVO.from_voxels(ctx, 9, g["voxels"])
    hits = s.cast_rays(g["origin"], g["dir"])
    assert_hits_equal(hits, g["hits"], hit_flag(hits), "kat")
    assert list(hits["voxel"][0]) == [411, 311, 211]
    h = s.castRay(g["origin"][0], g["dir"][0])                  # reference signature, one ray
    assert h.cell and h.complexity == 14 and list(h.normal) == [0.0, 0.0, -4.0]


def test_golden_terrain(vrt, ctx, terrain9_nodes):
    g = golden("lsvo_terrain9.npz")
    s = vrt.LSVO(ctx, terrain9_nodes, 9)
    for key, coef, bias in (("hits_coef0", 0.0, 0.0), ("hits_coef05", 0.5, 0.0), ("hits_bias", 0.25, 0.001)):
        hits = s.cast_rays(g["origin"], g["dir"], coef, bias)
        assert_hits_equal(hits, g[key], hit_flag(hits), key)
        assert s.last_complexity() == int(g[key]["complexity"].sum())


def test_golden_random_scene(vrt, ctx):
    g = golden("lsvo_random6.npz")
    s = vrt.LSVO(ctx, g["nodes"], 6)
    for key, coef in (("hits_coef0", 0.0), ("hits_coef05", 0.5)):
        hits = s.cast_rays(g["origin"], g["dir"], coef, 0.0)
        assert_hits_equal(hits, g[key], hit_flag(hits), key)


def test_oracle_large_seeded(vrt, ctx, port, terrain9_nodes):
    rng = np.random.default_rng(7)
    n = 1 << 20
    o = rng.uniform(1, 2, (n, 3)).astype(np.float32)
    o[:, 1] = rng.uniform(1.0, 1.45, n)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    s = vrt.LSVO(ctx, terrain9_nodes, 9)
    got = s.cast_rays(o, d)
    want = port.lsvo_cast(terrain9_nodes, 9, o, d, threads=8)
    assert_hits_equal(got, want, hit_flag(got), "1M random rays")
    m = hit_flag(got)
    assert np.array_equal(got["voxel"][m], want["voxel"][m]) and np.array_equal(got["face"][m], want["face"][m])
    assert np.array_equal(got["scale"][m], want["scale"][m])


def test_edge_cases(vrt, ctx, port):
    g = golden("lsvo_random6.npz")
    s = vrt.LSVO(ctx, g["nodes"], 6)
    assert len(s.cast_rays(np.zeros((0, 3)), np.zeros((0, 3)))) == 0             # empty batch
    # ragged sizes around the block size, origins outside the cube, zero / denormal direction components
    rng = np.random.default_rng(3)
    for n in (1, 31, 33, 127, 129, 1000):
        o = rng.uniform(-0.5, 3.5, (n, 3)).astype(np.float32)
        d = rng.normal(size=(n, 3)).astype(np.float32)
        d[rng.random((n, 3)) < 0.2] = 0.0
        d[rng.random((n, 3)) < 0.05] = np.float32(1e-40)
        d[rng.random((n, 3)) < 0.05] = np.float32(-0.0)
        got = s.cast_rays(o, d)
        want = port.lsvo_cast(g["nodes"], 6, o, d)
        assert_hits_equal(got, want, hit_flag(got), "ragged n=%d" % n)
    # empty octree (root only)
    e = vrt.LSVO.from_voxels(ctx, 4, np.zeros((0, 3), np.uint32))
    got = e.cast_rays([[1.5, 1.5, 1.1]], [[0, 0, 1]])
    assert not hit_flag(got)[0]
    # full octree: every ray that enters the cube hits at the entry face
    full = np.stack(np.meshgrid(*[np.arange(8)] * 3, indexing="ij"), -1).reshape(-1, 3)
    f = vrt.LSVO.from_voxels(ctx, 3, full)
    fn = vrt.host_build_lsvo_from_voxels(3, full)
    o, d = np.float32([[1.5, 1.5, 0.5], [1.3, 1.7, 1.2]]), np.float32([[0, 0, 1], [0.3, -0.2, 0.5]])
    got = f.cast_rays(o, d)
    assert_hits_equal(got, port.lsvo_cast(fn, 3, o, d), hit_flag(got), "full octree")
    # entering the root's first child without an ADVANCE step leaves the face mask at 0: normal is all zero
    assert hit_flag(got)[0] and got["distance"][0] == 0.5 and not np.any(got["normal"][0])


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_non_finite_rays_are_misses(vrt, port, terrain9_nodes, variant):
    """The reference's loop never ends on a NaN / infinite ray; engine and oracle define a miss of complexity 0."""
    c = vrt.Context(0)
    c.set_option("cast_variant", variant)
    s = vrt.LSVO(c, terrain9_nodes, 9)
    rng = np.random.default_rng(9)
    n = 4096
    o = rng.uniform(1.0, 2.0, (n, 3)).astype(np.float32)
    o[:, 1] = rng.uniform(1.0, 1.3, n)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    bad = rng.random(n) < 0.3
    which = rng.integers(0, 6, n)
    value = rng.choice(np.float32([np.nan, np.inf, -np.inf]), n)
    for i in np.flatnonzero(bad):
        (o if which[i] < 3 else d)[i, which[i] % 3] = value[i]
    got = s.cast_rays(o, d)
    want = port.lsvo_cast(terrain9_nodes, 9, o, d)
    assert_hits_equal(got, want, hit_flag(got), "non-finite rays")
    assert not hit_flag(got)[bad].any() and (got["complexity"][bad] == 0).all() and (got["complexity"][~bad] > 0).all()
    s.close()
    c.close()


@pytest.mark.parametrize("variant,refill", [(0, 8), (1, 0), (1, 1), (1, 8), (1, 20), (1, 32), (2, 0), (3, 0)])
def test_kernel_variants_agree(vrt, port, terrain9_nodes, variant, refill):
    """One-thread-per-ray and persistent/regenerating kernels give byte-identical hit records for every refill
    threshold (scheduling must not change results)."""
    c = vrt.Context(0)
    c.set_option("cast_variant", variant)
    c.set_option("refill_cast", refill)
    s = vrt.LSVO(c, terrain9_nodes, 9)
    rng = np.random.default_rng(11)
    for n in (1, 77, 4097, 300001):
        o = rng.uniform(1, 2, (n, 3)).astype(np.float32)
        o[:, 1] = rng.uniform(1.0, 1.45, n)
        d = rng.normal(size=(n, 3)).astype(np.float32)
        got = s.cast_rays(o, d, 0.0 if n % 2 else 0.5, 0.0)
        want = port.lsvo_cast(terrain9_nodes, 9, o, d, 0.0 if n % 2 else 0.5, 0.0, threads=8)
        assert_hits_equal(got, want, hit_flag(got), "variant %d refill %d n=%d" % (variant, refill, n))
        assert s.last_complexity() == int(want["complexity"].sum())
    c.close()


@pytest.mark.parametrize("depth", [8, 9, 10])
def test_device_scene_construction_is_byte_identical(vrt, ctx, depth):
    """vrt_lsvo_create_terrain builds T(D) on the GPU: the LNode array equals the host builder's (which equals the
    reference's compileSVO, tests/test_host_logic.py) byte for byte."""
    import hashlib
    s = vrt.LSVO.from_terrain(ctx, depth)                     # on_device=True
    got = s.download_nodes()
    want = vrt.host_build_terrain_lsvo(depth)
    assert len(got) == len(want)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    if depth == 9:
        assert hashlib.sha256(got.tobytes()).hexdigest() == str(golden("terrain_heights.npz")["sha_nodes9"])
    s.close()


def test_device_scene_large_depths(vrt, ctx):
    """2048^3 and 4096^3: slot counts match the host builder's closed-form count; a cast works on the result."""
    import ctypes as C
    for depth, guard in ((11, 0), (12, -1)):
        s = vrt.LSVO.from_terrain(ctx, depth, guard)
        n = C.c_uint64(0)
        h = vrt.host_terrain_heights(1 << depth)
        vrt.capi.check(vrt.capi.lib().vrt_host_build_terrain_lsvo(depth, vrt.capi.ptr(h), None, 0, C.byref(n)))
        assert len(s) == n.value
        S = float(1 << depth)
        o = np.float32([[1.5, 1 + (S / 2 - 56) / S, 1.5]])
        got = s.cast_rays(o, np.float32([[0.2, 0.5, 0.8]]))
        assert hit_flag(got)[0] and got["complexity"][0] > depth
        s.close()


@pytest.mark.parametrize("variant", [0, 1])
def test_compact_layout_gives_identical_results(vrt, port, terrain9_nodes, textures, variant):
    """The compact breadth-first node array (live nodes only) changes memory, not results: casts and frames are
    byte-identical to the reference layout, for random voxel scenes too."""
    c = vrt.Context(0)
    c.set_option("cast_variant", variant)
    s = vrt.LSVO(c, terrain9_nodes, 9)
    rng = np.random.default_rng(21)
    n = 200000
    o = rng.uniform(1, 2, (n, 3)).astype(np.float32)
    o[:, 1] = rng.uniform(1.0, 1.45, n)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    ref_layout = s.cast_rays(o, d, 0.25, 0.0)
    s.set_layout(1, l2_persist=True)
    assert s.info()["device_bytes"] < 1.2 * terrain9_nodes.nbytes       # reference copy + ~1/8
    compact = s.cast_rays(o, d, 0.25, 0.0)
    assert np.array_equal(ref_layout.view(np.uint8), compact.view(np.uint8))
    want = port.lsvo_cast(terrain9_nodes, 9, o[:50000], d[:50000], 0.25, 0.0, threads=8)
    assert_hits_equal(compact[:50000], want, hit_flag(compact[:50000]), "compact")
    # a frame through the compact layout
    s.set_textures(*textures)
    cam = vrt.Camera(position=(256, 200, 256), view_angle=(0.3, -0.35), aperture=0.5, focal_length=60.0)
    rc = vrt.RayCaster(s, (160, 90))
    rc.setLightPosition(np.float32([-200, -1000, -300]) * np.float32(1 / 512.0) + np.float32(1))
    rc.use_samples, rc.use_gi, rc.gi_bounces = True, True, 2
    a = rc.render(cam, spp=2).copy()
    s.set_layout(0)
    rc.resetSamples()
    b = rc.render(cam, spp=2)
    assert np.array_equal(a, b) and a[..., :3].max() > 0
    # random voxel scene, shallow depth, empty and full octrees
    g = golden("lsvo_random6.npz")
    r6 = vrt.LSVO(c, g["nodes"], 6)
    r6.set_layout(1)
    hits = r6.cast_rays(g["origin"], g["dir"])
    assert_hits_equal(hits, g["hits_coef0"], hit_flag(hits), "compact random6")
    e = vrt.LSVO.from_voxels(c, 4, np.zeros((0, 3), np.uint32))
    e.set_layout(1)
    assert not hit_flag(e.cast_rays([[1.5, 1.5, 1.1]], [[0, 0, 1]]))[0]
    c.close()


def _heightfield_voxels(heights):
    """The fill rule of main.cpp:70-76 as a voxel list in SVO::setCell coordinates (for the host flattener)."""
    S = heights.shape[0]
    out = []
    for x in range(S):
        for z in range(S):
            hmax = max(16, min(S // 2, int(heights[x, z])))
            ys = np.arange(1, hmax) + S // 2
            out.append(np.stack([np.full(len(ys), x), ys, np.full(len(ys), z)], 1))
    return np.concatenate(out).astype(np.uint32)


@pytest.mark.parametrize("depth", [5, 6, 8])
def test_heightfield_scene_and_edits(vrt, ctx, port, depth):
    """Dynamic scenes: a world given by column heights is flattened on the device byte-identically to the host
    flattener (compileSVO order), also after rectangles of columns were edited; casts stay bit-exact."""
    S = 1 << depth
    rng = np.random.default_rng(depth)
    heights = rng.integers(-5, S // 2 + 8, (S, S)).astype(np.int32)         # some below the 16-voxel floor, some above S/2
    s = vrt.LSVO.from_heightfield(ctx, depth, heights)
    want = vrt.host_build_lsvo_from_voxels(depth, _heightfield_voxels(heights))
    assert np.array_equal(s.download_nodes().view(np.uint64), want.view(np.uint64))
    for k in range(3):
        nx, nz = int(rng.integers(1, S // 2)), int(rng.integers(1, S // 2))
        x0, z0 = int(rng.integers(0, S - nx + 1)), int(rng.integers(0, S - nz + 1))
        patch = rng.integers(0, S // 2, (nx, nz)).astype(np.int32)
        heights[x0:x0 + nx, z0:z0 + nz] = patch
        if k == 1:
            s.set_layout(1)                                # edits rebuild the compact copy as well
        s.edit_heights(x0, z0, patch)
        assert np.array_equal(s.heights(), heights)
        want = vrt.host_build_lsvo_from_voxels(depth, _heightfield_voxels(heights))
        assert np.array_equal(s.download_nodes().view(np.uint64), want.view(np.uint64)), "edit %d" % k
        o = rng.uniform(1.0, 2.0, (2000, 3)).astype(np.float32)
        o[:, 1] = rng.uniform(1.0, 1.2, 2000)
        d = rng.normal(size=(2000, 3)).astype(np.float32)
        got = s.cast_rays(o, d)
        assert_hits_equal(got, port.lsvo_cast(want, depth, o, d), hit_flag(got), "after edit %d" % k)
    with pytest.raises(vrt.VrtError):
        s.edit_heights(S - 1, 0, np.zeros((2, 1), np.int32))       # rectangle out of range
    s.close()
    # the demo terrain as an editable scene equals the terrain scene
    t = vrt.LSVO.from_heightfield(ctx, 8)
    assert np.array_equal(t.download_nodes().view(np.uint64), vrt.host_build_terrain_lsvo(8).view(np.uint64))
    assert np.array_equal(t.heights().reshape(-1), vrt.host_terrain_heights(256).reshape(-1))
    with pytest.raises(vrt.VrtError):
        vrt.LSVO.from_terrain(ctx, 8).edit_heights(0, 0, np.zeros((1, 1), np.int32))   # not a heightfield scene
    t.close()


@pytest.mark.parametrize("depth,n", [(1, 3), (2, 40), (3, 200), (5, 3000), (6, 20000), (8, 150000)])
def test_voxel_set_scene_built_on_device(vrt, ctx, port, depth, n):
    """Arbitrary voxel sets flattened on the GPU (sorted path keys, DFS numbering by binary searches): byte-identical to
    the host flattener (= compileSVO order), duplicates included; casts bit-exact."""
    S = 1 << depth
    rng = np.random.default_rng(100 + depth)
    vox = rng.integers(0, S, (n, 3)).astype(np.uint32)
    vox = np.concatenate([vox, vox[: n // 3]])                              # duplicates, like repeated setCell calls
    s = vrt.LSVO.from_voxels(ctx, depth, vox, on_device=True)
    want = vrt.host_build_lsvo_from_voxels(depth, vox)
    assert np.array_equal(s.download_nodes().view(np.uint64), want.view(np.uint64))
    assert s.voxel_count() == len(np.unique(vox, axis=0))
    o = rng.uniform(0.8, 2.2, (4000, 3)).astype(np.float32)
    d = rng.normal(size=(4000, 3)).astype(np.float32)
    got = s.cast_rays(o, d)
    assert_hits_equal(got, port.lsvo_cast(want, depth, o, d), hit_flag(got), "device-built voxel scene")
    s.close()


def test_voxel_set_edge_cases_and_edits(vrt, ctx, port):
    """Empty and full worlds, and LSVO::setCell given a meaning (a no-op in the reference, lsvo.hpp:26): voxels added and
    removed on the device; after every edit the array equals a fresh flattening of the edited set."""
    depth, S = 5, 32
    e = vrt.LSVO.from_voxels(ctx, depth, np.zeros((0, 3), np.uint32), on_device=True)
    assert np.array_equal(e.download_nodes().view(np.uint64), vrt.host_build_lsvo_from_voxels(depth, np.zeros((0, 3), np.uint32)).view(np.uint64))
    full = np.stack(np.meshgrid(*[np.arange(8)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.uint32)
    f = vrt.LSVO.from_voxels(ctx, 3, full, on_device=True)
    assert np.array_equal(f.download_nodes().view(np.uint64), vrt.host_build_lsvo_from_voxels(3, full).view(np.uint64))
    f.close()
    rng = np.random.default_rng(77)
    world = set()
    for step in range(6):
        if step % 2 == 0:
            vox = rng.integers(0, S, (int(rng.integers(1, 900)), 3)).astype(np.uint32)
            e.set_cells(vox, True)
            world |= set(map(tuple, vox.tolist()))
        else:
            have = np.array(sorted(world), np.uint32)
            gone = have[rng.random(len(have)) < 0.4]
            gone = np.concatenate([gone, rng.integers(0, S, (50, 3)).astype(np.uint32)])   # some were never set
            e.set_cells(gone, False)
            world -= set(map(tuple, gone.tolist()))
        have = np.array(sorted(world), np.uint32).reshape(-1, 3)
        want = vrt.host_build_lsvo_from_voxels(depth, have)
        assert e.voxel_count() == len(have)
        assert np.array_equal(e.download_nodes().view(np.uint64), want.view(np.uint64)), "step %d" % step
        o = rng.uniform(0.8, 2.2, (2000, 3)).astype(np.float32)
        d = rng.normal(size=(2000, 3)).astype(np.float32)
        got = e.cast_rays(o, d)
        assert_hits_equal(got, port.lsvo_cast(want, depth, o, d), hit_flag(got), "after edit %d" % step)
    e.set_cells(np.array(sorted(world), np.uint32), False)                  # remove everything: the empty world again
    assert e.voxel_count() == 0 and len(e.download_nodes()) == 1
    with pytest.raises(vrt.VrtError):
        e.set_cells([[S, 0, 0]], True)                                      # out of range: UB in the reference, an error here
    with pytest.raises(vrt.VrtError):
        vrt.LSVO.from_terrain(ctx, 8).set_cells([[1, 1, 1]], True)          # not a voxel-set scene
    e.close()
