import sys
import numpy as np
sys.path.insert(0, ".")
from lasgun_b200 import _native as N, parity, scenes
from oracle import pyoracle as po
name = sys.argv[1] if len(sys.argv) > 1 else "cornell"
sc, (w, h) = {"cornell": lambda: scenes.cornell_groups((192, 192), 2), "nested": lambda: scenes.nested_groups((256, 192), 1),
              "nested_root": lambda: scenes.nested_groups((160, 120), 1, transformed_root=True)}[name]()
o = po.OracleScene(sc)
ref = o.capture(w, h, aov=True)
ctx = N.Context(0)
dev = N.DeviceScene(ctx, N.FlatScene(sc))
out = dev.capture_aov(w, h)
spp = sc.camera.num_samples()
def retest(pid, i):
    rays = o.camera_sample((i // spp) % w, (i // spp) // w, w, h)
    return o.retest(pid, rays[i % spp, :3], rays[i % spp, 3:])
a = parity.aov_report(out, ref, retest)
print({k: v for k, v in a.items() if k != "examples"})
for e in a["examples"][:12]:
    print(e)
ids_d, ids_r = out["prim_id"].reshape(-1), ref["prim_id"].reshape(-1)
td, tr = out["t"].reshape(-1), ref["t"].reshape(-1)
bad = np.nonzero((ids_d != ids_r) | (td != tr))[0]
print("n bad", len(bad))
for i in bad[:16]:
    px = i // spp
    print(i, "pixel", px % w, px // w, "s", i % spp, "dev", ids_d[i], repr(td[i]), "ref", ids_r[i], repr(tr[i]))
f = parity.film_report(out["rgba"], ref["rgba"]); print(f)
occ = (out["occl"].reshape(-1) != ref["occl"].reshape(-1)); print("occl diff", occ.sum())
