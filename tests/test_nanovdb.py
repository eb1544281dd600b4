"""NanoVDB ingest (.vndb / .nvdb -> LumenRenderer::CreateVolume): the library's reader (csrc/lb_nanovdb.cpp) and its numpy
restatement (tests/nanovdb_tools.py) against golden data produced by the REFERENCE's own vendored NanoVDB
(tests/golden/make_golden_nanovdb.py; the same library PTVolume::Load calls, PT/Framework/PTVolume.cpp:93-98).
Bar: bit-exact — meta data, every sampled voxel value and active state, and sum + bit-fold of every voxel of the index bounding box."""
import os
import struct

import numpy as np
import pytest

import lumenrenderer_b200 as lr
from lumenrenderer_b200 import api, scenes
from lumenrenderer_b200 import nanovdb as nv
import nanovdb_tools as nt
from conftest import GOLDEN, rel_l1

FIXTURES = ["fog5_raw", "fog12_zip", "ls10_zip"]
SPHERE_ASSET = "/root/reference/Lumen_Engine/Sandbox/assets/volume/Sphere.vndb"      # present in the build container only


def fixture_path(name):
    return SPHERE_ASSET if name == "sphere_asset" else os.path.join(GOLDEN, "nanovdb", name + ".vndb")


@pytest.fixture(scope="module")
def gold():
    z = np.load(os.path.join(GOLDEN, "nanovdb_reference.npz"))
    return lambda name, key: z[f"{name}/{key}"]


def names():
    return FIXTURES + [pytest.param("sphere_asset", marks=pytest.mark.skipif(not os.path.exists(SPHERE_ASSET), reason="reference asset not on this machine"))]


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("name", names())
def test_numpy_restatement_matches_the_reference_library(gold, name):
    g = nt.read_grid(open(fixture_path(name), "rb").read())
    assert (g.grid_type, g.grid_class) == (gold(name, "grid_type"), gold(name, "grid_class"))
    assert np.array_equal(np.concatenate([g.index_min, g.index_max]), gold(name, "index_bbox"))
    assert np.array_equal(np.concatenate([g.world_min, g.world_max]), gold(name, "world_bbox"))
    assert np.array_equal(g.voxel_size, gold(name, "voxel_size"))
    assert np.array_equal(g.map_matrix.reshape(-1), gold(name, "map_matrix")) and np.array_equal(g.map_translation, gold(name, "map_translation"))
    assert g.active_voxels == gold(name, "active_voxels") and np.array_equal(np.array(g.node_count, np.uint32), gold(name, "node_count"))
    assert np.array_equal(bits([g.background, g.value_min, g.value_max]), bits(gold(name, "background_min_max")))
    ijk = gold(name, "sample_ijk")
    n = len(ijk) if name != "sphere_asset" else 400
    got = [g.value(int(i), int(j), int(k)) for i, j, k in ijk[:n]]
    assert np.array_equal(bits([v for v, _ in got]), bits(gold(name, "sample_value")[:n]))
    assert np.array_equal(np.array([a for _, a in got]), gold(name, "sample_active")[:n])
    dense = g.dense()
    assert dense.size == gold(name, "dense_count") and float(dense.astype(np.float64).sum()) == gold(name, "dense_sum")
    if name != "sphere_asset":
        assert nt.fold_bits(dense) == gold(name, "dense_fold")
    # the world box of the stored voxels is the grid's worldBBox() for the grids NanoVDB's builders write
    lo, hi = g.volume_box()
    assert np.array_equal(lo, g.world_min.astype(np.float32)) and np.array_equal(hi, g.world_max.astype(np.float32))


@pytest.mark.parametrize("name", names())
def test_library_reader_matches_the_reference_library_and_the_restatement(gold, name):
    path = fixture_path(name)
    with nv.NanoVdbGrid(path) as g:
        i = g.info
        assert (i["grid_type"], i["grid_class"]) == (gold(name, "grid_type"), gold(name, "grid_class"))
        assert np.array_equal(np.concatenate([i["index_min"], i["index_max"]]), gold(name, "index_bbox"))
        assert np.array_equal(np.concatenate([i["world_min"], i["world_max"]]), gold(name, "world_bbox"))
        assert np.array_equal(i["voxel_size"], gold(name, "voxel_size"))
        assert np.array_equal(i["map_matrix"].reshape(-1), gold(name, "map_matrix")) and np.array_equal(i["map_translation"], gold(name, "map_translation"))
        assert i["active_voxels"] == gold(name, "active_voxels") and np.array_equal(np.array(i["node_count"], np.uint32), gold(name, "node_count"))
        assert np.array_equal(bits([i["background"], i["value_min"], i["value_max"]]), bits(gold(name, "background_min_max")))
        assert i["version"][0] == 29 and i["grid_count"] == 1 and i["codec"] == (1 if name.endswith("_zip") else 0)
        v, a = g.values(gold(name, "sample_ijk"))
        assert np.array_equal(bits(v), bits(gold(name, "sample_value"))) and np.array_equal(a, gold(name, "sample_active"))
        dense = g.dense()
        assert dense.size == gold(name, "dense_count") and float(dense.astype(np.float64).sum()) == gold(name, "dense_sum")
        r = nt.read_grid(open(path, "rb").read())
        assert np.array_equal(bits(dense), bits(r.dense()))
        assert np.array_equal(bits(g.dense(as_density=True)), bits(r.density()))
        # same bytes through the in-memory entry point
        with nv.NanoVdbGrid(open(path, "rb").read()) as m:
            assert np.array_equal(bits(m.dense()), bits(dense))


def test_density_convention():
    """Fog volume: values are densities. Level set: interior ramp -v / background clamped to 1, exterior 0 (sdfToFogVolume)."""
    with nv.NanoVdbGrid(fixture_path("fog12_zip")) as g:
        assert g.info["grid_class"] == nv.CLASS_FOG_VOLUME
        d = g.dense(as_density=True)
        assert d.min() == 0.0 and d.max() == 1.0 and np.array_equal(d, np.maximum(g.dense(), 0))
    with nv.NanoVdbGrid(fixture_path("ls10_zip")) as g:
        assert g.info["grid_class"] == nv.CLASS_LEVEL_SET and g.info["background"] == np.float32(1.5)      # 3 voxels x 0.5
        raw, d = g.dense(), g.dense(as_density=True)
        assert (d[raw >= 0] == 0).all() and (d[raw <= -1.5] == 1).all() and d.max() == 1.0
        c = tuple(s // 2 for s in d.shape)
        assert d[c] == 1.0 and d[0, 0, 0] == 0.0                # centre of the sphere is an interior TILE value, the corner is background


def test_multi_grid_files_and_grid_index():
    a = open(fixture_path("fog5_raw"), "rb").read(); b = open(fixture_path("ls10_zip"), "rb").read()
    both = a + b                                                 # two segments (io::writeGrid appends segments the same way)
    with nv.NanoVdbGrid(both, 0) as g0, nv.NanoVdbGrid(both, 1) as g1:
        assert g0.info["name"] == "sphere_fog" and g1.info["name"] == "sphere_ls" and g0.info["grid_count"] == 2
        assert np.array_equal(bits(g1.dense()), bits(nt.read_grid(both, 1).dense()))
    with pytest.raises(nv.NanoVdbError, match="exceeds the grid count"):
        nv.NanoVdbGrid(both, 2)


def test_malformed_files_are_errors_not_crashes(tmp_path):
    raw = bytearray(open(fixture_path("fog5_raw"), "rb").read())
    with pytest.raises(nv.NanoVdbError, match="not a NanoVDB file"):
        nv.NanoVdbGrid(b"glTF" + bytes(60))
    with pytest.raises(nv.NanoVdbError, match="cannot read"):
        nv.NanoVdbGrid(str(tmp_path / "missing.vndb"))
    with pytest.raises(nv.NanoVdbError, match="truncated"):
        nv.NanoVdbGrid(bytes(raw[:len(raw) // 2]))
    with pytest.raises(nv.NanoVdbError, match="truncated"):
        nv.NanoVdbGrid(bytes(raw[:100]))
    old = bytearray(raw); struct.pack_into("<I", old, 8, 28 << 21)                       # file version 28: the reference rejects it too
    with pytest.raises(nv.NanoVdbError, match="ABI 28"):
        nv.NanoVdbGrid(bytes(old))
    blosc = bytearray(raw); struct.pack_into("<H", blosc, 14, 2)
    with pytest.raises(nv.NanoVdbError, match="BLOSC") as e:
        nv.NanoVdbGrid(bytes(blosc))
    assert e.value.code == -5
    name_size, = struct.unpack_from("<I", raw, 16 + 136)
    grid0 = 16 + 160 + name_size
    dbl = bytearray(raw); struct.pack_into("<I", dbl, grid0 + 628, 2)                    # GridType::Double
    with pytest.raises(nv.NanoVdbError, match="only float grids") as e:
        nv.NanoVdbGrid(bytes(dbl))
    assert e.value.code == -5
    # corrupt child offsets / tile child ids: every lookup stays inside the buffer or raises
    rng = np.random.default_rng(3)
    for trial in range(40):
        bad = bytearray(raw)
        for _ in range(8):
            at = grid0 + 672 + int(rng.integers(0, 400 if trial % 2 else len(raw) - grid0 - 676)) & ~3
            struct.pack_into("<I", bad, at, int(rng.integers(0, 2 ** 32)))
        try:
            with nv.NanoVdbGrid(bytes(bad)) as g:
                if np.prod(np.maximum(g.info["index_max"].astype(np.int64) - g.info["index_min"] + 1, 0)) < 1 << 24:
                    g.dense()
                g.values(rng.integers(-60, 60, (64, 3)))
        except nv.NanoVdbError:
            pass
    zbad = bytearray(open(fixture_path("fog12_zip"), "rb").read()); zbad[-20] ^= 0xFF
    try:
        nv.NanoVdbGrid(bytes(zbad))
    except nv.NanoVdbError:
        pass


def test_create_volume_from_file_rejects_other_containers_without_touching_the_device():
    import ctypes as C
    b = lr.bindings()
    out = C.c_int32()
    fake = C.c_void_p(1)                                        # never dereferenced: the extension check comes first
    assert b.volume_create_file(fake, b"bunny.vdb", C.byref(out)) == -5 and b"OpenVDB" in b.nanovdb_last_error()
    assert b.volume_create_file(fake, b"cloud.txt", C.byref(out)) == -5
    assert b.volume_create_file(None, b"x.vndb", C.byref(out)) == -1


def _room_without_volumes():
    s = scenes.fog_room(grid=8)
    s.volumes = []
    return s


@pytest.mark.gpu
@pytest.mark.parametrize("name,mode", [("ls10_zip", lr.VOLUME_DELTA), ("fog12_zip", lr.VOLUME_DELTA), ("fog5_raw", lr.VOLUME_COMPAT)])
def test_volume_from_nanovdb_file_renders_like_the_oracle(oracle, name, mode):
    """CreateVolume(path) on the GPU == the oracle fed the numpy restatement's density box: identical volume bounds (primary volume
    hits), radiance within the media tolerance of tests/test_gpu_frame.py."""
    path = fixture_path(name)
    st = lr.Settings(width=160, height=96, depth=3, restir=False, volume_mode=mode)
    g = lr.Renderer(st); c = api.Renderer(oracle, st)
    scene = _room_without_volumes()
    g.load_scene(scene); c.load_scene(scene)
    ref = nt.read_grid(open(path, "rb").read())
    lo, hi = ref.volume_box()
    centre = 0.5 * (lo.astype(np.float64) + hi); size = float((hi - lo).max())
    s = 9.0 / size                                               # the grid's box becomes 9 units wide, in front of the two boxes
    m = np.array([[s, 0, 0, 0.0 - s * centre[0]], [0, s, 0, 9.0 - s * centre[1]], [0, 0, s, 7.0 - s * centre[2]], [0, 0, 0, 1]], np.float32)
    c0 = api.Renderer(oracle, st); c0.load_scene(scene); c0.render_frames(2); empty = c0.read_hdr()[..., :3].copy(); c0.close()
    hv = nv.create_volume_from_file(g, path)
    g.add_volume_instance(hv, m, 1.5)
    c.add_volume_instance(c.create_volume(ref.density(), lo, hi), m, 1.5)
    for _ in range(2):
        g.render_frames(1); c.render_frames(1)
    hg, hc = g.read_hdr()[..., :3], c.read_hdr()[..., :3]
    vg, vc = g.read_channel(lr.CHANNEL_VOLUMETRIC), c.read_channel(lr.CHANNEL_VOLUMETRIC)
    assert np.isfinite(hg).all() and rel_l1(hc, empty) > 0.1, "the medium must be visible in this framing"
    assert rel_l1(hg, hc) < 2e-3, rel_l1(hg, hc)
    assert rel_l1(vg, vc) < 5e-3, rel_l1(vg, vc)
    g.close(); c.close()
