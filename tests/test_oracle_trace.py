"""Known-answer and property tests of the oracle's ray/triangle + BVH code (parity for this part is UNPINNED by the
reference — OptiX is closed — so it is pinned analytically here; the CUDA traversal is then compared against it)."""
import numpy as np
import pytest

from lumenrenderer_b200 import api, scenes


def _single_triangle(oracle, verts):
    r = api.Renderer(oracle, api.Settings(width=8, height=8, depth=1, restir=False))
    m = r.create_material(api.MaterialData(metallic_factor=0.0))
    p = r.create_primitive(np.asarray(verts, np.float32), [0, 1, 2], m)
    r.add_mesh_instance(r.create_mesh([p]))
    return r


def test_single_triangle_known_answers(oracle):
    r = _single_triangle(oracle, [[0, 0, -2], [1, 0, -2], [0, 1, -2]])
    o = np.array([[0.25, 0.25, 0], [0.1, 0.7, 0], [0.9, 0.9, 0], [0.25, 0.25, -4]], np.float32)
    d = np.array([[0, 0, -1], [0, 0, -1], [0, 0, -1], [0, 0, 1]], np.float32)
    h = r.trace_closest(o, d, 0.01, 100.0)
    assert h["t"][0] == 2.0 and h["u"][0] == 0.25 and h["v"][0] == 0.25          # u, v weight vertices 1 and 2
    assert h["t"][1] == 2.0 and np.isclose(h["u"][1], 0.1) and np.isclose(h["v"][1], 0.7)
    assert h["t"][2] == -1.0                                                      # outside
    assert h["t"][3] == 2.0                                                       # no back-face culling (OPTIX_RAY_FLAG_NONE)
    assert r.trace_any(o, d, [100, 1.5, 100, 100]).tolist() == [1, 0, 0, 1]        # tmax clips the second ray
    assert r.trace_closest(o[:1], d[:1], 2.0, 100.0)["t"][0] == -1.0              # t must be > tmin strictly
    r.close()


def test_watertight_shared_edges_and_vertices(oracle):
    """Rays through shared edges / vertices of a fan never slip through (Woop et al. watertightness)."""
    r = api.Renderer(oracle, api.Settings(width=8, height=8, depth=1, restir=False))
    m = r.create_material(api.MaterialData(metallic_factor=0.0))
    n = 24
    ang = np.linspace(0, 2 * np.pi, n, endpoint=False)
    verts = np.concatenate([[[0.013, -0.007, -3.0]], np.stack([np.cos(ang), np.sin(ang), np.full(n, -3.0)], 1)]).astype(np.float32)
    idx = np.array([[0, 1 + k, 1 + (k + 1) % n] for k in range(n)], np.uint32)
    r.add_mesh_instance(r.create_mesh([r.create_primitive(verts, idx, m)]))
    rng = np.random.default_rng(1)
    t = (0.95 * rng.random((4000, 1))).astype(np.float32)          # stay off the rim: points beyond it are legitimately outside
    k = rng.integers(0, n, 4000)
    on_edge = verts[0] * (1 - t) + verts[1 + k] * t                      # points on the spokes
    o = np.array([[0.3, -0.2, 1.0]], np.float32).repeat(4000, 0)
    d = on_edge - o; d /= np.linalg.norm(d, axis=1, keepdims=True)
    h = r.trace_closest(o, d.astype(np.float32), 0.01, 100.0)
    assert (h["t"] > 0).all()
    o2 = np.array([[0.0, 0.0, 0.0]], np.float32); d2 = (verts[0] / np.linalg.norm(verts[0]))[None].astype(np.float32)
    assert r.trace_closest(o2, d2, 0.01, 100.0)["t"][0] > 0                # through the hub vertex
    r.close()


@pytest.mark.parametrize("name", ["cornell", "gallery"])
def test_bvh_equals_brute_force(oracle, name):
    import ctypes as C
    scene = scenes.SCENES[name]()
    r = api.Renderer(oracle, api.Settings(width=8, height=8, depth=1, restir=False)); r.load_scene(scene)
    rng = np.random.default_rng(2)
    n = 3000
    o = (rng.random((n, 3)) * [4, 3, 4] - [2, 0.5, 2]).astype(np.float32)
    d = rng.normal(size=(n, 3)); d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    a = r.trace_closest(o, d)
    rays = np.ascontiguousarray(np.concatenate([o, d], 1)); b = np.empty(n, api.HIT_DTYPE)
    fn = oracle.lib.lo_debug_trace_closest_brute
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_float, C.c_float, C.c_void_p]
    assert fn(r._h, rays.ctypes.data, n, 0.01, 5000.0, b.ctypes.data) == 0
    assert np.array_equal(a, b)
    assert (a["t"] > 0).mean() > 0.15
    r.close()


def test_tie_break_prefers_smaller_ids(oracle):
    """Two coincident triangles: the documented tie rule returns the smaller (instance, primitive)."""
    r = api.Renderer(oracle, api.Settings(width=8, height=8, depth=1, restir=False))
    m = r.create_material(api.MaterialData(metallic_factor=0.0))
    v = np.array([[0, 0, -2], [1, 0, -2], [0, 1, -2]], np.float32)
    p = r.create_primitive(np.concatenate([v, v]), [3, 4, 5, 0, 1, 2], m)
    mesh = r.create_mesh([p]); r.add_mesh_instance(mesh); r.add_mesh_instance(mesh)
    h = r.trace_closest([[0.2, 0.2, 0]], [[0, 0, -1]])
    assert (h["instance"][0], h["primitive"][0]) == (0, 0)
    r.close()


def test_empty_scene_and_degenerate_inputs(oracle):
    r = api.Renderer(oracle, api.Settings(width=16, height=8, depth=2, restir=True))
    r.render_frames(1)
    assert np.all(r.read_hdr() == 0) and (r.read_primary_hits()["t"] == -1).all()
    m = r.create_material(api.MaterialData(metallic_factor=0.0))
    p = r.create_primitive(np.zeros((3, 3), np.float32), [0, 1, 2], m)            # zero-area triangle
    r.add_mesh_instance(r.create_mesh([p]))
    r.render_frames(1)
    assert np.isfinite(r.read_hdr()).all()
    with pytest.raises(api.LumenError):
        r.create_primitive(np.zeros((3, 3), np.float32), [0, 1, 5], m)            # index out of range
    with pytest.raises(api.LumenError):
        r.create_material(api.MaterialData(roughness_factor=0.0))                 # reference asserts roughness > 0
    r.close()
