"""Known answers for the lobes of SURVEY 8f item 4 (matte sigma > 0, metal, glass, mirror; Whitted recursion).

The reference holds no unit test, golden vector or fixture for src/core/bxdf/, src/material/ or src/integrate/ (its 17 tests
cover shapes and one surface interaction), so the oracle's restatement of these files is pinned on closed forms instead:
textbook Fresnel / Snell / Trowbridge-Reitz identities re-derived here in numpy, the reference's own formula where it departs
from the textbook (OrenNayar::new, diffuse.rs:31), and energy / reciprocity properties of the rendered result."""
import ctypes as C
import math

import numpy as np
import pytest

from lasgun_b200 import scenes
from lasgun_b200.api import MAT_GLASS, MAT_MATTE, MAT_METAL, MAT_MIRROR, MAT_PLASTIC


def _d(v):
    return (C.c_double * len(v))(*[float(x) for x in v])


def fresnel(oracle, kind, cos_i, p):
    out = (C.c_double * 3)()
    oracle.lib().orc_test_fresnel(kind, cos_i, _d(p), out)
    return np.array(out[:])


FRAME = [0, 0, 1, 0, 0, 1, 1, 0, 0]          # ng, ns, dpdu: the shading frame is the world frame


def bsdf(oracle, kind, kd, ks, rough, rough2, wo, wi, frame=FRAME):
    f, r, t = (C.c_double * 3)(), (C.c_double * 7)(), (C.c_double * 7)()
    oracle.lib().orc_test_bsdf(kind, _d(kd), _d(ks), rough, rough2, _d(frame), _d(wo), _d(wi), f, r, t)
    return np.array(f[:]), np.array(r[:]), np.array(t[:])


def unit(v):
    v = np.asarray(v, float)
    return v / np.linalg.norm(v)


def test_dielectric_fresnel(oracle):             # fresnel.rs:37-64
    assert fresnel(oracle, 0, 1.0, [1.0, 1.5]) == pytest.approx([0.04] * 3, abs=1e-15)          # ((n1 - n2) / (n1 + n2))^2
    assert fresnel(oracle, 0, 0.0, [1.0, 1.5]) == pytest.approx([1.0] * 3)                       # grazing (cos 0 counts as exiting: TIR)
    assert fresnel(oracle, 0, -0.5, [1.0, 1.5]) == pytest.approx([1.0] * 3)                      # from inside beyond the critical angle
    for c in (0.9, 0.5, 0.2):                     # unpolarised Fresnel equations
        st = math.sqrt(1 - c * c) / 1.5
        ct = math.sqrt(1 - st * st)
        rp = (1.5 * c - ct) / (1.5 * c + ct)
        rs = (c - 1.5 * ct) / (c + 1.5 * ct)
        assert fresnel(oracle, 0, c, [1.0, 1.5])[0] == pytest.approx(0.5 * (rp * rp + rs * rs), rel=1e-13)
        assert fresnel(oracle, 0, -ct, [1.0, 1.5])[0] == pytest.approx(0.5 * (rp * rp + rs * rs), rel=1e-12)    # reciprocity


def test_conductor_fresnel(oracle):              # fresnel.rs:68-90
    for c in (1.0, 0.7, 0.3):                     # k = 0: a conductor is a dielectric
        assert fresnel(oracle, 1, c, [1, 1, 1, 1.5, 1.5, 1.5, 0, 0, 0])[0] == pytest.approx(fresnel(oracle, 0, c, [1.0, 1.5])[0], rel=1e-12)
    n, k = 0.2, 3.9                               # normal incidence: ((n - 1)^2 + k^2) / ((n + 1)^2 + k^2)
    assert fresnel(oracle, 1, 1.0, [1, 1, 1, n, n, n, k, k, k])[0] == pytest.approx(((n - 1) ** 2 + k * k) / ((n + 1) ** 2 + k * k), rel=1e-12)
    assert fresnel(oracle, 2, 0.3, [])[0] == 1.0  # Substance::NoOp


def test_specular_samples(oracle):               # specular.rs:17-25, 44-66 through BSDF::sample_f (bsdf.rs:94-140)
    wo = unit([0.3, -0.2, 0.8])
    f, r, t = bsdf(oracle, MAT_MIRROR, [0.5, 0.6, 0.7], [0, 0, 0], 0, 0, wo, wo)
    assert np.all(f == 0)                                                     # specular lobes scatter only through sample_f
    assert r[3:6] == pytest.approx([-wo[0], -wo[1], wo[2]]) and r[6] == 1.0
    assert r[:3] == pytest.approx(np.minimum(np.array([0.5, 0.6, 0.7]) / wo[2], 1.0))      # kr / |cos|, clamped (bsdf.rs:122)
    assert t[6] == 0.0                                                        # a mirror has no transmission lobe
    eta = 1.5
    f, r, t = bsdf(oracle, MAT_GLASS, [0.9, 0.9, 0.9], [0.8, 0.7, 0.6], eta, 0, wo, wo)
    sin_i = math.sqrt(1 - wo[2] ** 2)
    sin_t = np.linalg.norm(t[3:5])
    assert sin_t == pytest.approx(sin_i / eta, rel=1e-13) and t[5] < 0        # Snell, other side
    assert wo[0] * t[4] - wo[1] * t[3] == pytest.approx(0, abs=1e-15)           # in the plane of incidence
    F = fresnel(oracle, 0, t[5], [1.0, eta])[0]
    assert t[:3] == pytest.approx(np.minimum(np.array([0.8, 0.7, 0.6]) * (1 - F) / abs(t[5]), 1.0), rel=1e-13)
    assert r[:3] == pytest.approx(np.minimum(0.9 * fresnel(oracle, 0, wo[2], [1.0, eta])[0] / wo[2], 1.0), rel=1e-13)
    wo_in = unit([0.8, 0.0, -0.6])                                            # from inside, beyond the critical angle: no refraction
    f, r, t = bsdf(oracle, MAT_GLASS, [0.9, 0.9, 0.9], [0.8, 0.7, 0.6], eta, 0, wo_in, wo_in)
    assert t[6] == 0.0 and np.all(r[:3] == 1.0)                               # total internal reflection: F = 1, 0.9 / 0.6 clamped


def test_oren_nayar(oracle):                     # diffuse.rs:28-56
    sigma = 25.0 * math.pi / 180.0
    s2 = sigma * sigma
    A = 1.0 - (s2 / 2.0 * (s2 + 0.33))            # as the reference writes it (diffuse.rs:31), not the textbook's s2 / (2 (s2 + 0.33))
    B = 0.45 * s2 / (s2 + 0.09)
    kd = np.array([0.7, 0.5, 0.3])
    n = [0, 0, 1]
    assert bsdf(oracle, MAT_MATTE, kd, [0, 0, 0], 25.0, 0, n, n)[0] == pytest.approx(kd / math.pi * A, rel=1e-14)
    wo, wi = unit([0.5, 0.1, 0.6]), unit([0.3, 0.4, 0.4])
    so, si = math.sqrt(1 - wo[2] ** 2), math.sqrt(1 - wi[2] ** 2)
    dcos = max(0.0, (wo[0] * wi[0] + wo[1] * wi[1]) / (so * si))
    sin_alpha, tan_beta = (so, si / wi[2]) if wi[2] > wo[2] else (si, so / wo[2])
    want = kd / math.pi * (A + B * dcos * sin_alpha * tan_beta)
    assert bsdf(oracle, MAT_MATTE, kd, [0, 0, 0], 25.0, 0, wo, wi)[0] == pytest.approx(want, rel=1e-12)
    assert bsdf(oracle, MAT_MATTE, kd, [0, 0, 0], 25.0, 0, wi, wo)[0] == pytest.approx(want, rel=1e-12)      # reciprocal
    assert bsdf(oracle, MAT_MATTE, kd, [0, 0, 0], 0.0, 0, wo, wi)[0] == pytest.approx(kd / math.pi)          # sigma = 0: Lambertian
    assert np.all(bsdf(oracle, MAT_MATTE, kd, [0, 0, 0], 25.0, 0, wo, [wi[0], wi[1], -wi[2]])[0] == 0)       # other side of ng


def tr_reference(ax, ay, wo, wi, F):
    """Torrance-Sparrow with an anisotropic Trowbridge-Reitz distribution, textbook form."""
    wh = unit(wo + wi)

    def lam(w):
        t2 = (1 - w[2] ** 2) / w[2] ** 2
        s2 = 1 - w[2] ** 2
        a2 = (w[0] ** 2 * ax * ax + w[1] ** 2 * ay * ay) / s2
        return (math.sqrt(1 + a2 * t2) - 1) / 2

    c2 = wh[2] ** 2
    e = (wh[0] ** 2 / (ax * ax) + wh[1] ** 2 / (ay * ay)) / c2
    D = 1 / (math.pi * ax * ay * c2 * c2 * (1 + e) ** 2)
    G = 1 / (1 + lam(wo) + lam(wi))
    return D * G * F / (4 * wo[2] * wi[2])


def test_metal_and_plastic_microfacet(oracle):   # microfacet.rs:31-66, 101-115; metal.rs; plastic.rs
    wo, wi = unit([0.4, 0.2, 0.7]), unit([-0.3, 0.5, 0.6])
    eta, k = [0.2, 0.9, 1.1], [3.9, 2.4, 2.2]
    wh = unit(wo + wi)
    F = np.array([fresnel(oracle, 1, float(np.dot(wi, wh)), [1, 1, 1] + eta + k)[c] for c in range(3)])
    got = bsdf(oracle, MAT_METAL, eta, k, 0.08, 0.45, wo, wi)[0]
    assert got == pytest.approx(tr_reference(0.08, 0.45, wo, wi, F), rel=1e-11)
    assert bsdf(oracle, MAT_METAL, eta, k, 0.08, 0.45, wi, wo)[0] == pytest.approx(got, rel=1e-11)           # reciprocal
    rot = [0, 0, 1, 0, 0, 1, 0, 1, 0]             # dpdu along y: the anisotropy turns with the frame (bsdf.rs:33-36)
    sw = lambda v: np.array([v[1], -v[0], v[2]])
    assert bsdf(oracle, MAT_METAL, eta, k, 0.08, 0.45, wo, wi, rot)[0] == pytest.approx(tr_reference(0.08, 0.45, sw(wo), sw(wi), F), rel=1e-11)
    Fd = fresnel(oracle, 0, float(np.dot(wi, wh)), [1.0, 1.5])[0]
    ks = np.array([0.5, 0.7, 0.5])
    want = np.array([0.2, 0.3, 0.4]) / math.pi + ks * tr_reference(0.25, 0.25, wo, wi, Fd)
    assert bsdf(oracle, MAT_PLASTIC, [0.2, 0.3, 0.4], ks, 0.25, 0, wo, wi)[0] == pytest.approx(want, rel=1e-11)


def test_whitted_recursion_properties(oracle):   # integrate.rs:69-132
    """Depth 0 switches the specular rays off: glass and mirror then show only their (zero) direct term; each further level can
    only add light; a scene without specular lobes does not depend on the depth."""
    def film(scene_fn, **kw):
        sc, (w, h) = scene_fn(**kw)
        return oracle.OracleScene(sc).capture(w, h, li=True)["li"]
    l0 = film(scenes.simplereflect, supersampling=0, res=48, recursion=0)
    l1 = film(scenes.simplereflect, supersampling=0, res=48, recursion=1)
    l4 = film(scenes.simplereflect, supersampling=0, res=48, recursion=4)
    assert np.all(l1 >= l0) and np.all(l4 >= l1 - 1e-12) and (l4 > l1 + 1e-6).any() and (l1 > l0 + 1e-6).any()
    sc, (w, h) = scenes.simplereflect(0, 48, 0)
    r = oracle.OracleScene(sc).capture(w, h, aov=True, li=True)
    hit = r["prim_id"] != oracle.MISS
    assert np.all(r["li"][hit] == 0.0)            # every object there is glass or mirror: BSDF::f is zero (bxdf/mod.rs:172)
    a = film(scenes.simple, supersampling=0, res=48)
    sc, (w, h) = scenes.simple("b", 0, 48); sc.set_max_recursion_depth(0)
    assert np.array_equal(a, oracle.OracleScene(sc).capture(w, h, li=True)["li"])
