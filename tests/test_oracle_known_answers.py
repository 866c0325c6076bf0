"""The reference's own unit-test vectors against the CPU oracle.
Exact equality, as the Rust tests assert (`assert_eq!` on f64)."""
import numpy as np

from lasgun_b200.api import parse_obj_text

PLANE = """o plane
v -1 0 -1
v 1 0 -1
v 1 0 1
v -1 0 1

f 1 2 3
f 1 3 4
"""

UNIT = ([-1.0, -1.0, -1.0], [1.0, 1.0, 1.0])
WIDE = ([-1.1, -1.1, -1.0], [1.1, 1.1, 1.0])


def veq(a, b):
    return all(float(x) == float(y) for x, y in zip(a, b))   # -0.0 == 0.0, as Rust's PartialEq


def test_sphere_straight_on(oracle):          # sphere.rs:137
    hit, t, ng, _ = oracle.test_sphere([0, 0, 0], 1.0, [0, 0, 2], [0, 0, -1])
    assert hit and t == 1.0 and veq(ng, [0, 0, 1])


def test_sphere_inside(oracle):               # sphere.rs:149
    hit, t, ng, _ = oracle.test_sphere([0, 0, 0], 1.0, [0, 0, 0], [0, 0, 1])
    assert hit and t == 1.0 and veq(ng, [0, 0, -1])


def test_sphere_behind(oracle):               # sphere.rs:160 (rounded, as the reference does)
    hit, t, ng, _ = oracle.test_sphere([0, 0, 0], 1.0, [0, 0, -2], [0, 0, 1])
    assert hit and t == 1.0 and veq(np.round(ng), [0, 0, -1])


def test_cuboid_vectors(oracle):              # cuboid.rs:137-245
    cases = [
        (UNIT, [0, 0, -2], [0, 0, 1], 1.0, ("ng", [0, 0, -1])),
        (WIDE, [0, 0, -2], [1, 0, 1], 1.0, ("ng", [0, 0, -1])),
        (WIDE, [0, 0, -2], [1, 1, 1], 1.0, ("ng", [0, 0, -1])),
        (UNIT, [0, 0, 0], [0, 0, 1], 1.0, None),
        (UNIT, [0, 0, 0], [0, -1, 0], 1.0, None),
        (UNIT, [0.5, 0.5, 0.5], [1, 0, 1], None, None),
        (UNIT, [0, 0, 2], [0, 0, -1], 1.0, ("ng", [0, 0, 1])),
        (UNIT, [0, 2, 0], [0, -1, 0], 1.0, ("ns", [0, 1, 0])),
        (UNIT, [0, -2, 0], [0, 1, 0], 1.0, ("ns", [0, -1, 0])),
        (UNIT, [0, 2, 2], [0, -0.5, -1], 2.0, ("ng", [0, 1, 0])),
    ]
    for (mn, mx), o, d, t_exp, normal in cases:
        hit, t, ng, ns = oracle.test_cuboid(mn, mx, o, d)
        assert hit, (o, d)
        if t_exp is not None:
            assert t == t_exp, (o, d, t)
        if normal:
            assert veq(ng if normal[0] == "ng" else ns, normal[1]), (o, d, ng, ns)


def test_triangle_plane(oracle):              # triangle.rs:411 and :434
    obj = parse_obj_text(PLANE)
    hit, t, ng, _, which = oracle.test_mesh(obj, [0, 1, 0], [0, -1, 0])
    assert hit and t == 1.0 and veq(ng, [0, 1, 0])
    assert which == 0          # the ray lies on the shared diagonal: first triangle wins, second rejects t >= isect.t


def test_surface_interaction(oracle):         # surface.rs:194
    assert veq(oracle.test_surface(1.0, [1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 0, -1]), [0, 0, 1])


def test_quantisation_and_camera_shapes(oracle):
    """Film bytes: alpha is always 255 and every pixel is written (img.rs:56-61, lib.rs:152-161)."""
    from lasgun_b200 import scenes
    sc, _ = scenes.simple("a", 1, 64)
    out = oracle.OracleScene(sc).capture(64, 48, threads=2)
    assert out["rgba"].shape == (48, 64, 4) and (out["rgba"][..., 3] == 255).all()
    # capture_subset(k, n) partitions the film (lib.rs:114-141)
    o = oracle.OracleScene(sc)
    full = o.capture(64, 48)["rgba"]
    parts = np.zeros_like(full)
    for k in range(3):
        sub = o.capture(64, 48, subset=(3, k, 1))["rgba"].reshape(-1, 4)
        idx = np.arange(k, 64 * 48, 3)
        assert (np.delete(sub, idx, axis=0) == 0).all()
        parts.reshape(-1, 4)[idx] = sub[idx]
    assert np.array_equal(parts, full)


def test_oracle_reproduces_golden_films(oracle):
    """tests/golden/films.npz (made by tests/golden/make_golden.py) pins the oracle's own output for a small version of every
    config: an edit of the oracle that changes a single byte of any film fails here."""
    import importlib.util
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    gold = np.load(os.path.join(here, "golden", "films.npz"))
    assert sorted(gold.files) == sorted(mg.CASES)
    for name, mk in mg.CASES.items():
        sc, (w, h) = mk()
        assert np.array_equal(oracle.OracleScene(sc).capture(w, h)["rgba"], gold[name]), name
