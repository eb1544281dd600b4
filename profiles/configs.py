"""BASELINE.json's five configurations measured on one B200 (C2 is bench.py's own line; the others are parity-test cases, reported here
once for completeness): ms/frame (device, wall clock around N frames + synchronize after warm-up), rays/s, BVH build.
Usage (GPU box): python profiles/configs.py > gpurun_out/configs.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lumenrenderer_b200 as lr
from lumenrenderer_b200 import scenes


def run(name, scene, frames, warm, extra=None, **kw):
    g = lr.Renderer(lr.Settings(**kw))
    t = time.time(); g.load_scene(scene)
    if extra:
        extra(g)
    g.render_frames(warm); g.synchronize(); setup = time.time() - t
    t = time.time(); g.render_frames(frames); g.synchronize(); dt = (time.time() - t) / frames
    c = g.frame_counters(); st = g.frame_stats()
    rays = c["extend_rays"] + c["shadow_rays"] + c["visibility_rays"]
    out = {"config": name, "resolution": [kw["width"], kw["height"]], "depth": kw["depth"], "restir": bool(kw.get("restir", True)),
           "triangles": c["triangles"], "lights": c["lights"], "bvh_bytes": c["bvh_bytes"], "bvh_build_ms_both_hierarchies": c["bvh_build_us"] / 1e3,
           "ms_per_frame": dt * 1e3, "fps": 1.0 / dt, "rays_per_frame": rays, "mrays_per_s": rays / dt / 1e6,
           "extend_rays": c["extend_rays"], "shadow_rays": c["shadow_rays"], "visibility_rays": c["visibility_rays"],
           "stage_us_last_frame": {k: round(v, 1) for k, v in st.items()}, "setup_s_incl_scene_build_and_warmup": setup,
           "finite": bool(np.isfinite(g.read_hdr()).all())}
    g.close()
    print(json.dumps(out), flush=True)


run("C1 Cornell 256x256 depth 2 no ReSTIR", scenes.cornell_box(), 200, 10, width=256, height=256, depth=2, restir=False)
room = scenes.fog_room(grid=256)
run("C3 fog room 1440p depth 3 delta tracking (homogeneous box + 256^3 grid)", room, 20, 5, width=2560, height=1440, depth=3, restir=True, volume_mode=lr.VOLUME_DELTA)
run("C4 64 instances x 156K triangles (10 M) 1440p depth 5 ReSTIR", scenes.instanced_field(8, 280), 10, 3, width=2560, height=1440, depth=5, restir=True)
run("C5 atrium 3840x2160 64 spp progressive (one GPU; N GPUs: bench.py --gpus N --width 3840 --height 2160)", scenes.atrium(detail=0.78), 64, 3,
    extra=lambda g: g.set_blend_mode(True), width=3840, height=2160, depth=4, restir=True)
