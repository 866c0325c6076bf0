"""Per-sample radiance of a glass / mirror scene, device vs oracle: which samples differ, and what their primary hit is."""
import sys
import numpy as np
sys.path.insert(0, ".")
from lasgun_b200 import _native as N, scenes
from oracle import pyoracle as po
which = tuple(sys.argv[1].split(",")) if len(sys.argv) > 1 else ("mirror",)
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sc, (w, h) = scenes.mixed4k(res=(480, 270), supersampling=0, whitted=which)
sc.set_max_recursion_depth(depth)
o = po.OracleScene(sc)
ref = o.capture(w, h, aov=True, li=True)
ctx = N.Context(0)
dev = N.DeviceScene(ctx, N.FlatScene(sc))
out = dev.capture_aov(w, h, li=True)
d = np.abs(out["li"] - ref["li"]).max(axis=1)
scale = np.maximum(np.abs(ref["li"]).max(axis=1), 1e-3)
bad = np.nonzero(d > 1e-9 * scale)[0]
print(which, "depth", depth, "samples", len(d), "differing beyond 1e-9 rel", len(bad), "max abs", float(d.max()), "ids equal", bool((out["prim_id"] == ref["prim_id"]).all()))
print("rel diff histogram (log10):", np.histogram(np.log10(np.maximum(d / scale, 1e-18)), bins=[-18, -15, -13, -11, -9, -6, -3, 0, 3])[0].tolist())
for i in bad[:10]:
    x, y = int(i % w), int(i // w)
    rays = o.camera_sample(x, y, w, h)
    print("  sample", int(i), "pixel", (x, y), "prim", int(ref["prim_id"][i]), "t", float(ref["t"][i]), "dev li", out["li"][i].tolist(), "oracle li", ref["li"][i].tolist())
    print("     ray", rays[0].tolist())
