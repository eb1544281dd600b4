"""Dynamic-scene cost: wall time of a frame that follows an instance move. Round 1: flatten + BOTH hierarchies rebuilt + light
list (+ the frame); round 2: flatten + both hierarchies REFITTED (bvh_refit) + light list. C2 (262 K triangles) and C4 (10 M triangles, 64 instances).
LB_REFIT_MAX=0 forces the round-1 behaviour (rebuild on every move).   python profiles/dynamic_commit.py [c2|c4]"""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json, time, numpy as np, lumenrenderer_b200 as lr
from lumenrenderer_b200 import scenes
which = sys.argv[1] if len(sys.argv) > 1 else "c2"
scene = scenes.atrium(detail=0.78, texture_size=64) if which == "c2" else scenes.instanced_field()
moved = 3 if which == "c2" else 5
g = lr.Renderer(lr.Settings(width=640, height=360, depth=3, restir=True))
g.load_scene(scene); g.render_frames(2); g.synchronize()
fc = g.frame_counters()
out = {"scene": which, "triangles": fc["triangles"], "first_build_ms": fc["bvh_build_us"] / 1e3, "refit_max": os.environ.get("LB_REFIT_MAX")}
g.render_frames(3); g.synchronize()
t = time.time(); g.render_frames(5); g.synchronize(); out["static_frame_wall_ms"] = (time.time() - t) / 5 * 1e3
base = np.asarray(scene.instances[moved].get("transform", np.eye(4)), np.float32).reshape(4, 4)
walls, refits, builds = [], [], []
for k in range(8):
    m = base.copy(); m[:3, 3] += np.float32(0.01 * (k + 1))
    t = time.time(); g.set_instance_transform(moved, m); g.render_frames(1); g.synchronize(); walls.append((time.time() - t) * 1e3)
    fc = g.frame_counters(); refits.append(fc["bvh_refit_us"] / 1e3); builds.append(fc["bvh_build_us"] / 1e3)
out.update(move_frame_wall_ms=walls, refit_ms=refits, build_ms_counter=builds, bvh_refits=fc["bvh_refits"], move_cost_ms=float(np.median(walls) - out["static_frame_wall_ms"]))
print(json.dumps(out))
