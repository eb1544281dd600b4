#!/usr/bin/env python
"""The reference's own Sponza asset (Sandbox/assets/models/Sponza/Sponza.gltf: 262 267 triangles, 103 primitives, 25 materials, 69 textures of
which 65 are JPEG) through the native ingest (lb_gltf_*, PNG + JPEG decoded in the library) into the B200 renderer — parity against the oracle
at reduced resolution, then ms/frame at 2560x1440.  python profiles/sponza_real.py <Sponza.gltf> [out.json]

The asset is not part of this repository and /root/reference does not exist on the GPU box: for the run recorded in profiles/r02_sponza_real.json
a copy was placed under tmp_assets/ (git-ignored) for the duration of one gpurun call. Sponza has no emissive material, so — as SURVEY 8d C2
prescribes — lamps are added as override-emissive spheres (EmissionMode::OVERRIDE). The reference renders this asset UNSCALED (its root mesh
node keeps an identity world matrix, DESIGN.md "glTF ingest"), so the scene is ~3 700 units long and the lamps / camera are placed in those units."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as ge
import lumenrenderer_b200 as lr
from lumenrenderer_b200 import api, scenes
from lumenrenderer_b200.gltf import GltfDocument

path = sys.argv[1] if len(sys.argv) > 1 else "tmp_assets/Sponza/Sponza.gltf"
out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/r02_sponza_real.json"
scene, cam_pos, cam_rot, info, t_load = scenes.real_sponza(path)
res = {"asset": os.path.basename(path), "gltf_info": info, "load_and_decode_s": t_load, "triangles": scene.triangle_count()}

# ---- parity against the oracle, 480x270, depth 4, ReSTIR, 2 frames
st = lr.Settings(width=480, height=270, depth=4, restir=True)
with lr.Renderer(st) as g, api.Renderer(ge.oracle_bindings(), st) as c:
    for r in (g, c):
        r.load_scene(scene); r.set_camera(cam_pos, cam_rot)
    errs = []
    for frame in range(2):
        g.render_frames(1); c.render_frames(1)
        hg, hc = g.read_primary_hits(), c.read_primary_hits()
        same_hits = all(np.array_equal(hg[f], hc[f]) for f in ("instance", "primitive", "t", "u", "v"))
        same_surface = bool(np.array_equal(g.read_surface(), c.read_surface()))
        a, b = g.read_hdr()[..., :3].astype(np.float64), c.read_hdr()[..., :3].astype(np.float64)
        errs.append(float(np.abs(a - b).sum() / max(np.abs(b).sum(), 1e-30)))
        assert same_hits and same_surface, (frame, same_hits, same_surface)
    lg, lc = g.read_lights(), c.read_lights()
    fg, fc = g.frame_counters(), c.frame_counters()
    res["parity_480x270"] = {"primary_hits_bit_exact": True, "surface_records_bit_exact": True, "hit_fraction": float((hc["t"] > 0).mean()), "radiance_rel_l1_per_frame": errs,
                             "lights": int(len(lg[0])), "lights_bit_exact": bool(np.array_equal(lg[0], lc[0]) and np.array_equal(lg[1], lc[1])),
                             "rays_gpu": [fg["extend_rays"], fg["shadow_rays"], fg["visibility_rays"]], "rays_oracle": [fc["extend_rays"], fc["shadow_rays"], fc["visibility_rays"]],
                             "stack_overflows": fg["stack_overflows"], "mean_radiance": float(b.mean())}
    assert max(errs) < 2e-3 and fg["stack_overflows"] == 0, errs

# ---- 2560x1440 on the GPU
import torch
st = lr.Settings(width=2560, height=1440, depth=4, restir=True)
with lr.Renderer(st) as g:
    g.load_scene(scene); g.set_camera(cam_pos, cam_rot)
    g.render_frames(11); g.synchronize()
    t0 = time.time(); g.render_frames(20); g.synchronize(); ms = (time.time() - t0) / 20 * 1e3
    fc = g.frame_counters()
    rays = fc["extend_rays"] + fc["shadow_rays"] + fc["visibility_rays"]
    g.set_overlap(0); g.render_frames(1); g.render_frames(1)
    res["c2_real_sponza_1440p"] = {"ms_per_frame": ms, "fps": 1e3 / ms, "mrays_per_s": rays / ms / 1e3, "rays_per_frame": rays, "bvh_build_ms": fc["bvh_build_us"] / 1e3,
                                   "bvh_levels": fc["bvh_levels"], "stage_ms": {k: v / 1e3 for k, v in g.frame_stats().items()}, "finite": bool(np.isfinite(g.read_hdr()).all())}
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res))
