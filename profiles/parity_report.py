"""Prints the radiance parity (relative L1 against the CPU oracle) of the loaded library on the frame-test scenes — the numbers
the -m gpu tests bound, reported so that builds with different arithmetic flags can be compared."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lumenrenderer_b200 as lr
from lumenrenderer_b200 import api, scenes
import __graft_entry__ as ge


def rel_l1(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).sum() / max(np.abs(b).sum(), 1e-30))


orc = ge.oracle_bindings()
for name, scene, kw, frames in (
        ("cornell nee 256", scenes.cornell_box(), dict(width=256, height=256, depth=2, restir=False), 1),
        ("cornell restir 256 x4", scenes.cornell_box(), dict(width=256, height=256, depth=3, restir=True), 4),
        ("gallery restir 192x128 x3", scenes.material_gallery(), dict(width=192, height=128, depth=4, restir=True), 3)):
    st = lr.Settings(**kw)
    with lr.Renderer(st) as g, api.Renderer(orc, st) as c:
        g.load_scene(scene); c.load_scene(scene)
        errs = []
        for f in range(frames):
            g.render_frames(1); c.render_frames(1)
            errs.append(rel_l1(g.read_hdr()[..., :3], c.read_hdr()[..., :3]))
        a, b = g.read_hdr()[..., :3], c.read_hdr()[..., :3]
        bad = (np.abs(a - b).sum(-1) > 1e-3 * np.maximum(np.abs(b).sum(-1), 1e-3)).mean()
        print(f"{name}: rel-L1 per frame {' '.join('%.2e' % e for e in errs)}; pixels off by >1e-3 rel: {bad:.2e}")
