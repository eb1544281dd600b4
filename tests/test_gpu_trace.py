"""Extend / any-hit parity: the CUDA compressed-BVH8 traversal against the oracle's binary BVH on identical rays.
Hit ids, t and barycentrics must be bit-exact (the accepted set is a pure function of ray and triangle)."""
import numpy as np
import pytest

import lumenrenderer_b200 as lr
from lumenrenderer_b200 import api, scenes

pytestmark = pytest.mark.gpu


def _rays(rng, n, lo, hi):
    o = (rng.random((n, 3)) * (np.asarray(hi) - np.asarray(lo)) + np.asarray(lo)).astype(np.float32)
    d = rng.normal(size=(n, 3)); d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    k = n // 8                                    # axis-aligned and near-axis directions exercise the slab / shear corner cases
    d[:k] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, k)] * rng.choice([-1.0, 1.0], (k, 1)).astype(np.float32)
    d[k:2 * k] = d[:k] + (rng.normal(size=(k, 3)) * 1e-6).astype(np.float32)
    d[k:2 * k] /= np.linalg.norm(d[k:2 * k], axis=1, keepdims=True)
    return o, d


@pytest.mark.parametrize("name,n,lo,hi", [("cornell", 200000, (-1.2, -0.2, -1.2), (1.2, 2.2, 2.2)),
                                           ("gallery", 200000, (-4, 0, -3), (4, 4, 5))])
def test_closest_and_any_hit_bit_exact(oracle, name, n, lo, hi):
    scene = scenes.SCENES[name]()
    st = lr.Settings(width=8, height=8, depth=1, restir=False)
    g = lr.Renderer(st); c = api.Renderer(oracle, st)
    g.load_scene(scene); c.load_scene(scene)
    o, d = _rays(np.random.default_rng(3), n, lo, hi)
    hg, hc = g.trace_closest(o, d), c.trace_closest(o, d)
    assert (hc["t"] > 0).mean() > 0.3
    assert np.array_equal(hg, hc), f"{(hg != hc).sum()} of {n} closest hits differ"
    tmax = (np.random.default_rng(4).random(n) * 6).astype(np.float32)
    assert np.array_equal(g.trace_any(o, d, tmax), c.trace_any(o, d, tmax))
    g.close(); c.close()


def test_atrium_c2_geometry_bit_exact(oracle):
    """The C2 scene itself (262 K triangles, instanced + override materials), 300 K random rays + the 1440p primary-ray corner cases."""
    scene = scenes.atrium(texture_size=64)
    st = lr.Settings(width=8, height=8, depth=1, restir=False)
    g = lr.Renderer(st); c = api.Renderer(oracle, st)
    g.load_scene(scene); c.load_scene(scene)
    o, d = _rays(np.random.default_rng(5), 300000, (-17, 0.2, -7.5), (17, 13.5, 7.5))
    hg, hc = g.trace_closest(o, d), c.trace_closest(o, d)
    assert np.array_equal(hg, hc), f"{(hg != hc).sum()} closest hits differ"
    assert (hc["t"] > 0).mean() > 0.9
    tmax = (np.random.default_rng(6).random(o.shape[0]) * 30).astype(np.float32)
    assert np.array_equal(g.trace_any(o, d, tmax), c.trace_any(o, d, tmax))
    lg, lc = g.read_lights(), c.read_lights()
    assert lg[0].shape[0] >= 1000 and np.array_equal(lg[0], lc[0]) and np.array_equal(lg[1], lc[1])
    fc = g.frame_counters()
    assert fc["triangles"] == scene.triangle_count() == c.frame_counters()["triangles"]
    assert fc["stack_overflows"] == 0 and fc["bvh_levels"] + 2 <= 64        # the traversal stack never dropped a group
    g.close(); c.close()


def test_ties_edges_and_empty(oracle):
    st = lr.Settings(width=8, height=8, depth=1, restir=False)
    g = lr.Renderer(st); c = api.Renderer(oracle, st)
    for r in (g, c):
        m = r.create_material(lr.MaterialData(metallic_factor=0.0))
        assert (r.trace_closest([[0, 0, 0]], [[0, 0, -1]])["t"] == -1).all()      # empty scene
        v = np.array([[0, 0, -2], [1, 0, -2], [0, 1, -2], [1, 1, -2]], np.float32)
        p = r.create_primitive(np.concatenate([v, v]), [4, 5, 6, 0, 1, 2, 1, 3, 2, 5, 7, 6], m)   # coincident duplicates + shared edge
        mesh = r.create_mesh([p]); r.add_mesh_instance(mesh); r.add_mesh_instance(mesh)
    rng = np.random.default_rng(7)
    t = rng.random((5000, 1)).astype(np.float32)
    pts = np.array([1, 0, -2], np.float32) * (1 - t) + np.array([0, 1, -2], np.float32) * t            # on the shared diagonal
    o = np.zeros((5000, 3), np.float32) + np.array([0.3, 0.3, 1.0], np.float32)
    d = pts - o; d /= np.linalg.norm(d, axis=1, keepdims=True)
    hg, hc = g.trace_closest(o, d.astype(np.float32)), c.trace_closest(o, d.astype(np.float32))
    assert (hc["t"] > 0).all() and np.array_equal(hg, hc)
    assert hg["instance"].max() == 0                                              # tie rule: smaller instance wins
    g.close(); c.close()
