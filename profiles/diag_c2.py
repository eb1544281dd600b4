"""Where does a full-size C2 frame differ from the oracle? Per-channel relative L1 error, share of differing pixels, reservoir agreement."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lumenrenderer_b200 as lr
from lumenrenderer_b200 import api, scenes
import __graft_entry__ as entry

def rel(a, b):
    return float(np.abs(a.astype(np.float64) - b).sum() / max(np.abs(b.astype(np.float64)).sum(), 1e-30))

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2560, 1440)
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 1
for restir in (False, True):
    st = lr.Settings(width=W, height=H, depth=4, restir=restir)
    scene = scenes.atrium(detail=0.78, texture_size=256)
    g = lr.Renderer(st); c = api.Renderer(entry.oracle_bindings(), st)
    g.load_scene(scene); c.load_scene(scene)
    for f in range(frames):
        g.render_frames(1); c.render_frames(1)
        hg, hc = g.read_hdr()[..., :3], c.read_hdr()[..., :3]
        bad = ~np.isclose(hg, hc, rtol=1e-3, atol=1e-6).all(axis=-1)
        print(f"restir={restir} frame {f}: hdr rel-L1 {rel(hg, hc):.3e}, pixels off {bad.mean():.3e}, max abs diff {np.abs(hg - hc).max():.3e}, sum gpu {hg.sum():.6e} oracle {hc.sum():.6e}")
        for ch, name in ((0, "direct"), (1, "indirect")):
            a, b = g.read_channel(ch)[..., :3], c.read_channel(ch)[..., :3]
            d = ~np.isclose(a, b, rtol=1e-3, atol=1e-6).all(axis=-1)
            print(f"   {name}: rel-L1 {rel(a, b):.3e}, pixels off {d.mean():.3e}, top diffs {np.sort(np.abs(a - b).reshape(-1))[-3:]}")
        if restir:
            rg, rc = g.read_reservoirs(), c.read_reservoirs()
            print(f"   reservoirs: count differs {(rg[..., 2] != rc[..., 2]).mean():.3e}, weightSum rel-L1 {rel(rg[..., 0], rc[..., 0]):.3e}, weight rel-L1 {rel(rg[..., 1], rc[..., 1]):.3e}, sample position differs {(np.abs(rg[..., 4:7] - rc[..., 4:7]).max(axis=-1) > 1e-4).mean():.3e}")
        cg, cc = g.frame_counters(), c.frame_counters()
        print("   rays", {k: (cg[k], cc[k]) for k in ("extend_rays", "shadow_rays", "visibility_rays")})
    g.close(); c.close()
