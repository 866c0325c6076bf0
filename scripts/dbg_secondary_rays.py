"""Closest hits of rays shaped like the specular rays of a Whitted frame (origins on surfaces, unit directions): device vs oracle."""
import sys
import numpy as np
sys.path.insert(0, ".")
from lasgun_b200 import _native as N, scenes
from oracle import pyoracle as po
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
sc, (w, h) = scenes.mixed4k(res=(64, 36), supersampling=0, whitted=True)
o = po.OracleScene(sc)
ctx = N.Context(0)
dev = N.DeviceScene(ctx, N.FlatScene(sc))
rng = np.random.default_rng(7)
# camera-like rays first; their hit points become the origins of the second batch
d0 = rng.normal(size=(n, 3)); d0 /= np.linalg.norm(d0, axis=1, keepdims=True)
rays0 = np.concatenate([np.tile([0.0, 60.0, 520.0], (n, 1)), d0 * np.array([1.0, 0.6, -1.0]) * 0 + np.stack([rng.uniform(-0.5, 0.5, n), rng.uniform(-0.4, 0.2, n), -np.ones(n)], 1)], axis=1)
for name, rays in (("camera-like", rays0),):
    ids, t, ng, ns = dev.trace_rays(rays)
    oi, ot = o.trace_rays(rays)
    print(name, "n", n, "hits", int((oi != po.MISS).sum()), "id mismatches", int((ids != oi).sum()), "t mismatches among equal ids", int(((ids == oi) & (t != ot) & (oi != po.MISS)).sum()))
hit = oi != po.MISS
p = rays0[hit, :3] + rays0[hit, 3:] * ot[hit, None]
nrm = ng[hit]
nrm = np.where((np.sum(nrm * rays0[hit, 3:], axis=1) > 0)[:, None], -nrm, nrm)
d1 = rng.normal(size=p.shape); d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
for sign, name in ((1.0, "leaving the surface (reflection-like)"), (-1.0, "entering the surface (transmission-like)")):
    org = p + sign * nrm * 1.455e-11 * 1.0
    dirs = np.where((np.sum(d1 * nrm, axis=1) * sign > 0)[:, None], d1, -d1)
    rays = np.concatenate([org, dirs], axis=1)
    ids, t, _, _ = dev.trace_rays(rays)
    oi2, ot2 = o.trace_rays(rays)
    bad = ids != oi2
    print(name, "n", len(rays), "hits", int((oi2 != po.MISS).sum()), "id mismatches", int(bad.sum()), "of them oracle miss", int((bad & (oi2 == po.MISS)).sum()), "device miss", int((bad & (ids == po.MISS)).sum()),
          "t mismatches among equal ids", int(((ids == oi2) & (t != ot2) & (oi2 != po.MISS)).sum()))
    k = np.nonzero(bad)[0][:12]
    for i in k:
        print("   ray", i, "dev", ids[i], t[i], "oracle", oi2[i], ot2[i], "rel dt", abs(t[i] - ot2[i]) / max(ot2[i], 1e-300) if np.isfinite(t[i]) and np.isfinite(ot2[i]) else None)
