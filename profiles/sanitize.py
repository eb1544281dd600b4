#!/usr/bin/env python
"""Workload of the compute-sanitizer runs (profiles/r02_sanitizer.sh): the small configurations of the parity suite — C1 Cornell (NEE and ReSTIR),
the material gallery (alpha pass-through, glass, clear coat) and the fog room in both volume modes — two frames each, with the side-stream
overlap off (mode 0) and on (mode 5: shadow rays of wave d under the extend of wave d + 1, late waves beside the ReSTIR chain), plus one scene
re-commit (BVH rebuild, light list) and the debug traces. Sizes are small because memcheck / racecheck run the kernels 10-100x slower."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lumenrenderer_b200 as lr
from lumenrenderer_b200 import scenes

W, H = (int(x) for x in (sys.argv[1:3] if len(sys.argv) > 2 else (96, 64)))
cases = [("cornell_nee", scenes.cornell_box(), dict(depth=3, restir=False)), ("cornell_restir", scenes.cornell_box(), dict(depth=4, restir=True)),
         ("gallery", scenes.material_gallery(), dict(depth=4, restir=True)), ("fog_compat", scenes.fog_room(16), dict(depth=3, restir=True)),
         ("fog_delta", scenes.fog_room(16), dict(depth=3, restir=True, volume_mode=lr.api.VOLUME_DELTA))]
for name, scene, kw in cases:
    for overlap in (0, 5, 13):
        r = lr.Renderer(lr.Settings(width=W, height=H, **kw)); r.load_scene(scene); r.set_overlap(overlap)
        r.render_frames(2)
        hdr = r.read_hdr(); fc = r.frame_counters()
        assert np.isfinite(hdr).all() and fc["stack_overflows"] == 0
        if overlap == 0 and name == "gallery":
            rng = np.random.default_rng(1); o = rng.uniform(-1, 1, (2000, 3)).astype(np.float32); d = rng.normal(size=(2000, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
            r.trace_closest(o, d); r.trace_any(o, d, np.full(2000, 5.0, np.float32)); r.read_gbuffer(); r.read_ldr()
        print(f"{name} overlap={overlap}: rays {fc['extend_rays']}+{fc['shadow_rays']}+{fc['visibility_rays']} mean {hdr[..., :3].mean():.4f}", flush=True)
        r.close()
# the optional TMA-staged spatial pass (mbarrier + cp.async.bulk.tensor), and a transform-only change (refit of both hierarchies) followed by a
# change that forces the rebuild (device-driven PLOC rounds, single-block tail, batched collapse)
os.environ["LB_SPATIAL_TMA"] = "1"
r = lr.Renderer(lr.Settings(width=W, height=H, depth=3, restir=True)); r.load_scene(scenes.material_gallery()); r.render_frames(2)
assert np.isfinite(r.read_hdr()).all(); r.close(); del os.environ["LB_SPATIAL_TMA"]
print("gallery with LB_SPATIAL_TMA=1", flush=True)
r = lr.Renderer(lr.Settings(width=W, height=H, depth=3, restir=True)); r.load_scene(scenes.cornell_box()); r.render_frames(1)
m = np.eye(4, dtype=np.float32); m[0, 3] = 0.05
r.set_instance_transform(1, m); r.render_frames(1); fc = r.frame_counters(); assert fc["bvh_refits"] == 1, fc
r.set_instance_emissiveness(1, 2, (1.0, 0.5, 0.2), 3.0); r.render_frames(1); fc = r.frame_counters(); assert fc["bvh_refits"] == 0 and fc["stack_overflows"] == 0, fc
assert np.isfinite(r.read_hdr()).all(); r.close()
print("cornell refit + rebuild", flush=True)
print("sanitize workload done")
