"""Film of a large glass / mirror scene against the oracle: how many bytes differ, by how much, in both Whitted modes."""
import sys
import numpy as np
sys.path.insert(0, ".")
from lasgun_b200 import _native as N, parity, scenes
from oracle import pyoracle as po
res = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (640, 360)
ss = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx = N.Context(0)
for which in (("glass",), ("mirror",), ("ground",), ("metal",), ("matte",), ("glass", "mirror", "ground")):
    sc, (w, h) = scenes.mixed4k(res=res, supersampling=ss, whitted=which)
    if len(sys.argv) > 4:
        sc.set_max_recursion_depth(int(sys.argv[4]))
    ref = po.OracleScene(sc).capture(w, h)["rgba"]
    dev = N.DeviceScene(ctx, N.FlatScene(sc))
    film, st = dev.capture(w, h)
    dev.destroy()
    d = np.abs(film.astype(int) - ref.astype(int)).max(axis=2)
    ys, xs = np.nonzero(d > 1)
    print(which, "pixels differing", int((d > 0).sum()), "by >1", int((d > 1).sum()), "max", int(d.max()), "secondary", st["secondary_rays"], "first", list(zip(xs[:6].tolist(), ys[:6].tolist())))
