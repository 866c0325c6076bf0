import os, sys, time
sys.path.insert(0, ".")
from lasgun_b200 import _native as N, scenes
sc, (w, h) = scenes.mesh1m(n=140, res=128)
ctx = N.Context(0)
flat = N.FlatScene(sc)
dev = N.DeviceScene(ctx, flat)
print("verify", dev.verify())
rgba, st = dev.capture(w, h)
print("render ok", st["render_ms"])
