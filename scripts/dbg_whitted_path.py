"""Every ray the oracle casts for one camera sample, re-traced on the device (lgb_trace_rays): where do the two disagree?"""
import sys
import numpy as np
sys.path.insert(0, ".")
from lasgun_b200 import _native as N, scenes
from oracle import pyoracle as po
which = tuple(sys.argv[1].split(","))
px, py = int(sys.argv[2]), int(sys.argv[3])
depth = int(sys.argv[4]) if len(sys.argv) > 4 else 3
sc, (w, h) = scenes.mixed4k(res=(480, 270), supersampling=0, whitted=which)
sc.set_max_recursion_depth(depth)
o = po.OracleScene(sc)
ctx = N.Context(0)
dev = N.DeviceScene(ctx, N.FlatScene(sc))
ray = o.camera_sample(px, py, w, h)[0]
li, tr = o.debug_li(ray)
out = dev.capture_aov(w, h, li=True)
print("oracle li", li.tolist(), "device li", out["li"][py * w + px].tolist())
ids, t, ng, ns = dev.trace_rays(tr[:, 2:8])
for k in range(len(tr)):
    oid = po.MISS if tr[k, 8] < 0 else int(tr[k, 8])
    shadow = tr[k, 0] == 1
    if shadow:      # occluded iff closest t < 1 (point.rs:48-49): the device answers that question with an any-hit to t < 1 in the frame, with a closest hit here
        agree = (tr[k, 9] < 1.0) == (t[k] < 1.0)
    else:
        agree = oid == int(ids[k]) and (tr[k, 9] == t[k] or oid == po.MISS)
    print("shadow" if shadow else "path  ", "depth", int(tr[k, 1]), "oracle", oid, tr[k, 9], "device", int(ids[k]), t[k], "" if agree else "   <-- DISAGREE", "o", tr[k, 2:5].tolist(), "d", tr[k, 5:8].tolist())
