"""First-contact GPU check: known answers, per-sample AOV parity and film parity vs the oracle."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from lasgun_b200 import _native as N, parity, scenes  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

ctx = N.Context(0)
print("ceilings", ctx.measure())

CASES = [
    ("simple_b", lambda: scenes.simple("b", 2, 256)),
    ("cornell", lambda: scenes.cornell((480, 270), 1)),
    ("mesh", lambda: scenes.mesh1m(n=120, res=256)),
    ("spheres", lambda: scenes.spheres1m(count=50000, res=256)),
    ("mixed", lambda: scenes.mixed4k(mesh_n=80, nspheres=8000, res=(384, 216), supersampling=1)),
]
only = sys.argv[1:] or None
for name, mk in CASES:
    if only and name not in only:
        continue
    sc, (w, h) = mk()
    o = po.OracleScene(sc)
    t0 = time.time(); ref = o.capture(w, h, aov=True, counters=True); t_ref = time.time() - t0
    for resplit in (False,):
        flat = N.FlatScene(sc)
        dev = N.DeviceScene(ctx, flat)
        out = dev.capture_aov(w, h)
        rgba, st = dev.capture(w, h)
        assert np.array_equal(rgba, out["rgba"]), "AOV film differs from plain film"
        spp = flat.spp
        def retest(pid, i):
            px = i // spp
            rays = o.camera_sample(px % w, px // w, w, h)
            return o.retest(pid, rays[i % spp, :3], rays[i % spp, 3:])
        a = parity.aov_report(out, ref, retest)
        f = parity.film_report(rgba, ref["rgba"])
        print(f"{name} resplit={resplit} {w}x{h} spp={spp}: render {st['render_ms']:.2f} ms ({(st['primary_rays']+st['shadow_rays'])/st['render_ms']/1e3:.1f} Mrays/s), oracle {t_ref*1e3:.0f} ms")
        print("   aov ", {k: v for k, v in a.items() if k != "examples"})
        if a["examples"]: print("   ex  ", a["examples"][:4])
        print("   film", f)
        print("   stats", {k: st[k] for k in ("primary_hits", "shadow_rays", "shadow_rays_traced", "shadow_occluded")}, "aov-stats", {k: out['stats'][k] for k in ("node_tests", "filter_tests", "exact_tests")}, "oracle", ref["counters"])
        dev.destroy()
