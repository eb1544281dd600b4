"""Dynamic-scene cost: wall time of a frame that follows an instance move, i.e. flatten + both hierarchies + light list + frame."""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time, numpy as np, lumenrenderer_b200 as lr
from lumenrenderer_b200 import scenes
g = lr.Renderer(lr.Settings(width=640, height=360, depth=3, restir=True))
s = scenes.atrium(detail=0.78, texture_size=64)
g.load_scene(s); g.render_frames(2); g.synchronize()
print('first build ms', g.frame_counters()['bvh_build_us']/1e3)
ts=[]
for k in range(6):
    t=time.time(); g.set_instance_transform(3, scenes.translate(0.01*k, 0, 0)); g.render_frames(1); g.synchronize(); ts.append((time.time()-t)*1e3)
    print('recommit frame wall ms %.1f build ms %.1f' % (ts[-1], g.frame_counters()['bvh_build_us']/1e3))
