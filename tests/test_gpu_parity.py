"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle
(SURVEY.md Appendix E).  Bars: closest-hit ids identical except documented near-ties within 1e-6
of t, zero hit/miss flips; RGBA8 within 1 LSB on >= 99.9% of pixels, alpha identical."""
import numpy as np
import pytest

from lasgun_b200 import parity, scenes
from lasgun_b200.api import Film, Material, Scene, parse_obj_text

pytestmark = pytest.mark.gpu

PLAST = Material.plastic([0.5, 0.5, 0.5], [0.3, 0.3, 0.3], 0.25)


def radiance_outliers(dev_li, ref_li, rel=1e-9):
    """Fraction of samples whose radiance (li, integrate.rs:23; f64 on both sides) differs from the oracle's by more than `rel`
    of its magnitude: a far finer comparison than the bytes of the film."""
    d = np.abs(dev_li - ref_li).max(axis=1)
    scale = np.maximum(np.abs(ref_li).max(axis=1), 1e-3)
    return float((d > rel * scale).mean())


def one_prim_scene(kind, *args):
    sc = Scene()
    if kind == "sphere":
        sc.root.add_sphere(args[0], args[1], PLAST)
    elif kind == "box":
        sc.root.add_box(args[0], args[1], PLAST)
    elif kind == "obj":
        sc.root.add_obj_of(sc.parse_obj(args[0]), PLAST)
    return sc


def trace(native, gpu_ctx, sc, o, d):
    dev = native.DeviceScene(gpu_ctx, native.FlatScene(sc))
    ids, t, ng, ns = dev.trace_rays(np.array([list(o) + list(d)], np.float64))
    dev.destroy()
    return int(ids[0]), float(t[0]), ng[0], ns[0]


def veq(a, b):
    return all(float(x) == float(y) for x, y in zip(a, b))


def test_known_answers_on_device(native, gpu_ctx):
    """The reference's own unit-test vectors, exact as the Rust asserts."""
    sph = one_prim_scene("sphere", [0, 0, 0], 1.0)
    i, t, ng, _ = trace(native, gpu_ctx, sph, [0, 0, 2], [0, 0, -1]); assert i == 0 and t == 1.0 and veq(ng, [0, 0, 1])       # sphere.rs:137
    i, t, ng, _ = trace(native, gpu_ctx, sph, [0, 0, 0], [0, 0, 1]); assert i == 0 and t == 1.0 and veq(ng, [0, 0, -1])       # sphere.rs:149
    i, t, ng, _ = trace(native, gpu_ctx, sph, [0, 0, -2], [0, 0, 1]); assert i == 0 and t == 1.0 and veq(np.round(ng), [0, 0, -1])   # sphere.rs:160
    unit = one_prim_scene("box", [-1, -1, -1], [1, 1, 1]); wide = one_prim_scene("box", [-1.1, -1.1, -1.0], [1.1, 1.1, 1.0])
    cases = [(unit, [0, 0, -2], [0, 0, 1], 1.0, ("ng", [0, 0, -1])), (wide, [0, 0, -2], [1, 0, 1], 1.0, ("ng", [0, 0, -1])),
             (wide, [0, 0, -2], [1, 1, 1], 1.0, ("ng", [0, 0, -1])), (unit, [0, 0, 0], [0, 0, 1], 1.0, None),
             (unit, [0, 0, 0], [0, -1, 0], 1.0, None), (unit, [0.5, 0.5, 0.5], [1, 0, 1], None, None),
             (unit, [0, 0, 2], [0, 0, -1], 1.0, ("ng", [0, 0, 1])), (unit, [0, 2, 0], [0, -1, 0], 1.0, ("ns", [0, 1, 0])),
             (unit, [0, -2, 0], [0, 1, 0], 1.0, ("ns", [0, -1, 0])), (unit, [0, 2, 2], [0, -0.5, -1], 2.0, ("ng", [0, 1, 0]))]   # cuboid.rs:137-245
    for sc, o, d, t_exp, normal in cases:
        i, t, ng, ns = trace(native, gpu_ctx, sc, o, d)
        assert i == 0, (o, d)
        if t_exp is not None:
            assert t == t_exp, (o, d, t)
        if normal:
            assert veq(ng if normal[0] == "ng" else ns, normal[1]), (o, d, ng, ns)
    plane = one_prim_scene("obj", "v -1 0 -1\nv 1 0 -1\nv 1 0 1\nv -1 0 1\nf 1 2 3\nf 1 3 4\n")
    i, t, ng, _ = trace(native, gpu_ctx, plane, [0, 1, 0], [0, -1, 0])                                                          # triangle.rs:411,434
    assert i == 0 and t == 1.0 and veq(ng, [0, 1, 0])     # shared diagonal: first triangle wins


SMALL = {
    "simple_b_9spp": lambda: scenes.simple("b", 2, 192),
    "simple_a_1spp": lambda: scenes.simple("a", 0, 192),
    "cornell_4spp": lambda: scenes.cornell((320, 180), 1),
    "mesh": lambda: scenes.mesh1m(n=100, res=192),
    "spheres": lambda: scenes.spheres1m(count=40000, res=192),
    "mixed_4spp": lambda: scenes.mixed4k(mesh_n=64, nspheres=6000, res=(256, 144), supersampling=1),
    "ragged_edges": lambda: scenes.simple("b", 1, 100)[0:1] + ((101, 67),),   # film not a multiple of the tile size
    # nested BVH levels with transforms and swap_backface (bvh.rs:462-518, transform.rs:243-305)
    "cornell_groups_9spp": lambda: scenes.cornell_groups((192, 192), 2, eye=(0.21, 0.13, 5.0)),
    "nested_groups": lambda: scenes.nested_groups((256, 192), 1),
    "nested_groups_root_4spp": lambda: scenes.nested_groups((160, 120), 1, transformed_root=True)[0:1] + ((160, 120),),
}


def test_shading_fast_reciprocals(native, gpu_ctx):
    """The shading kernel replaces IEEE 1/x and 1/sqrt(x) by a MUFU seed + one third-order step (csrc/lgb_math.cuh): within a few
    ulp over the whole range shading meets (they only colour: no hit, shadow or sign decision goes through them)."""
    rng = np.random.default_rng(7)
    x = np.concatenate([np.exp(rng.uniform(np.log(1e-30), np.log(1e30), 200000)), rng.uniform(0.5, 2.0, 200000),
                        np.array([1.0, 2.0, 4.0, 0.25, 3.0, 1e-300, 1e300, 2.0 - 2.0 ** -52])])
    r, q = gpu_ctx.fastmath(x)
    assert np.max(np.abs(r * x - 1.0)) <= 6 * 2.0 ** -53
    assert np.max(np.abs(q * q * x - 1.0)) <= 12 * 2.0 ** -53


@pytest.mark.parametrize("name", sorted(SMALL))
def test_small_config_parity(native, oracle, gpu_ctx, name):
    sc, (w, h) = SMALL[name]()
    o = oracle.OracleScene(sc)
    ref = o.capture(w, h, aov=True, li=True)
    dev = native.DeviceScene(gpu_ctx, native.FlatScene(sc))
    out = dev.capture_aov(w, h, li=True)
    rgba, st = dev.capture(w, h)
    dev.destroy()
    spp = sc.camera.num_samples()
    assert radiance_outliers(out["li"], ref["li"]) <= 1e-4          # per-sample radiance to 1e-9 of its magnitude

    def retest(pid, i):
        rays = o.camera_sample((i // spp) % w, (i // spp) // w, w, h)
        return o.retest(pid, rays[i % spp, :3], rays[i % spp, 3:])

    a = parity.aov_report(out, ref, retest)
    assert a["mismatches"] == 0 and a["hit_miss_flips"] == 0, a
    assert a["near_ties"] <= 1e-4 * a["samples"], a
    assert a["t_bit_equal"] == a["t_compared"], a          # exact f64 tests: t is the reference's, bit for bit
    assert a["occl_diff"] == 0, a
    f = parity.film_report(rgba, ref["rgba"])
    assert f["alpha_equal"] and f["within_1_frac"] >= 0.999, f
    assert f["identical_frac"] >= 0.9999, f                # stronger than the bar: shading is f64 end to end
    assert np.array_equal(rgba, out["rgba"])               # all-shadow-rays (AOV) mode renders the same film
    assert st["primary_rays"] == w * h * spp and st["stack_overflow"] == 0
    assert st["shadow_rays"] == len(sc.lights) * st["primary_hits"]


@pytest.mark.parametrize("name", ["simple_b_9spp", "cornell_4spp", "mesh", "spheres", "mixed_4spp", "ragged_edges"])
def test_light_grids(native, oracle, gpu_ctx, name):
    """Shadow rays through the per-light cube-map grids (csrc/lgb_grid.cu) forced ON, also on scenes far too small for the automatic
    choice: occlusion bits, ids, t and film must be the oracle's, and the film the one the BVH shadow rays render."""
    if name not in SMALL:
        pytest.skip("no such small config")
    sc, (w, h) = SMALL[name]()
    ref = oracle.OracleScene(sc).capture(w, h, aov=True)
    films = {}
    try:
        for mode in (1, 0):
            gpu_ctx.set_light_grids(mode)
            dev = native.DeviceScene(gpu_ctx, native.FlatScene(sc))
            out = dev.capture_aov(w, h)
            films[mode], st = dev.capture(w, h)
            dev.destroy()
            a = parity.aov_report(out, ref)
            assert a["mismatches"] == 0 and a["hit_miss_flips"] == 0 and a["id_mismatch"] == 0, a
            assert a["occl_diff"] == 0, (mode, a)
            assert np.array_equal(films[mode], out["rgba"])
            assert st["shadow_rays"] == len(sc.lights) * st["primary_hits"]
    finally:
        gpu_ctx.set_light_grids(-1)
    assert np.array_equal(films[1], films[0])
    f = parity.film_report(films[1], ref["rgba"])
    assert f["alpha_equal"] and f["identical_frac"] >= 0.9999, f


@pytest.mark.parametrize("name", ["simple_b_9spp", "simple_a_1spp", "cornell_4spp", "mesh", "spheres", "mixed_4spp", "ragged_edges"])
def test_camera_grid(native, oracle, gpu_ctx, name):
    """Primary rays through the camera grid (pixel tiles listing the primitives their samples can see, csrc/lgb_grid.cu) forced ON:
    ids, t, occlusion and film are the oracle's and the film is the one the BVH traversal renders; also through capture_subset and
    a two-rank tile split."""
    sc, (w, h) = SMALL[name]()
    ref = oracle.OracleScene(sc).capture(w, h, aov=True)
    spp = sc.camera.num_samples()
    films = {}
    try:
        for mode in (1, 0):
            gpu_ctx.set_camera_grid(mode)
            dev = native.DeviceScene(gpu_ctx, native.FlatScene(sc))
            out = dev.capture_aov(w, h)
            films[mode], st = dev.capture(w, h)
            a = parity.aov_report(out, ref)
            assert a["id_mismatch"] == 0 and a["t_bit_equal"] == a["t_compared"] and a["occl_diff"] == 0, (mode, a)
            assert np.array_equal(films[mode], out["rgba"]) and st["primary_rays"] == w * h * spp
            if mode == 1:
                sub = np.zeros((h, w, 4), np.uint8)
                for k in range(3):
                    dev.capture_subset(k, 3, w, h, sub)
                assert np.array_equal(sub, films[1])
                import torch
                film = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
                for r in range(2):
                    dev.capture_device(w, h, film.data_ptr(), rank=r, ranks=2, want_stats=True)
                assert np.array_equal(film.cpu().numpy(), films[1])
            dev.destroy()
    finally:
        gpu_ctx.set_camera_grid(-1)
    assert np.array_equal(films[1], films[0])


def _columns_scene():
    """Spheres stacked behind one another as seen from the eye AND from a light beside it: a camera-grid tile and a light-grid cell with
    more than 1024 entries (the segmented sort), a second column of a few hundred (one warp and shared memory), scattered ones (a warp's
    lanes).  Every cell list must come out nearest first: the walks stop at the first entry that starts behind the best hit."""
    sc = Scene()
    sc.set_ambient_light([0.1, 0.1, 0.1])
    sc.set_solid_background([0.05, 0.05, 0.1])
    cam = sc.set_perspective_camera(40.0)
    cam.look_at([0.0, 0.0, 30.0], [0.0, 0.0, 0.0], [0.0, 1.0, 0.0])
    cam.set_supersampling(1)
    sc.add_point_light([0.4, 0.3, 31.0], [0.8, 0.8, 0.8], [1.0, 0.0, 0.0])
    sc.add_point_light([25.0, 20.0, 10.0], [0.5, 0.5, 0.5], [1.0, 0.0, 0.0])
    rng = np.random.default_rng(0x5EED0C01)
    cols = []
    for n, x0, y0, r in ((1500, 0.0, 0.0, 0.30), (300, 4.0, 1.0, 0.25)):
        k = np.arange(n)
        c = np.stack([x0 + 0.05 * rng.standard_normal(n), y0 + 0.05 * rng.standard_normal(n), 10.0 - 0.11 * k], axis=1)
        cols.append((c, np.full(n, r)))
    scat = np.concatenate([-8.0 + 16.0 * rng.random((2000, 2)), -30.0 + 35.0 * rng.random((2000, 1))], axis=1)
    cols.append((scat, 0.15 + 0.2 * rng.random(2000)))
    centers = np.concatenate([c for c, _ in cols]); radii = np.concatenate([r for _, r in cols])
    order = rng.permutation(len(centers))                   # (so that the columns do not arrive nearest first by construction)
    pal = [Material.plastic([0.8, 0.3, 0.2], [0.4, 0.4, 0.4], 0.2), Material.plastic([0.2, 0.7, 0.3], [0.3, 0.3, 0.3], 0.3), Material.plastic([0.3, 0.4, 0.9], [0.5, 0.5, 0.5], 0.1)]
    sc.root.add_spheres(centers[order], radii[order], pal, np.arange(len(centers)) % 3)
    return sc, (96, 96)


def test_grid_cells_of_every_length(native, oracle, gpu_ctx):
    """k_sort_cells / k_sort_medium / the segmented sort (csrc/lgb_grid.cu) through the frames that depend on them: with both grids
    forced on, ids, t, occlusion bits and film are the oracle's and the film is the one the BVH kernels render."""
    sc, (w, h) = _columns_scene()
    ref = oracle.OracleScene(sc).capture(w, h, aov=True)
    films = {}
    try:
        for mode in (1, 0):
            gpu_ctx.set_camera_grid(mode); gpu_ctx.set_light_grids(mode)
            dev = native.DeviceScene(gpu_ctx, native.FlatScene(sc))
            out = dev.capture_aov(w, h)
            films[mode], st = dev.capture(w, h)
            dev.destroy()
            a = parity.aov_report(out, ref)
            assert a["mismatches"] == 0 and a["hit_miss_flips"] == 0 and a["id_mismatch"] == 0 and a["t_bit_equal"] == a["t_compared"], (mode, a)
            assert a["occl_diff"] == 0, (mode, a)
            assert np.array_equal(films[mode], out["rgba"]) and st["primary_rays"] == w * h * sc.camera.num_samples()
    finally:
        gpu_ctx.set_camera_grid(-1); gpu_ctx.set_light_grids(-1)
    assert np.array_equal(films[1], films[0])
    f = parity.film_report(films[1], ref["rgba"])
    assert f["alpha_equal"] and f["identical_frac"] >= 0.9999, f


@pytest.mark.parametrize("name", ["simple_b_9spp", "mixed_4spp", "nested_groups", "whitted"])
def test_frames_in_bands(native, gpu_ctx, name):
    """A memory budget far below what the frame's per-sample buffers need (LGB_OPT_WAVE_BUDGET_MB): the frame is rendered band after
    band on the same buffers and must be the film of the single pass, through capture, capture_subset and the AOV entry."""
    sc, (w, h) = WHITTED["materials_4spp"]() if name == "whitted" else SMALL[name]()
    dev = native.DeviceScene(gpu_ctx, native.FlatScene(sc))
    try:
        ref, st1 = dev.capture(w, h)
        aov1 = dev.capture_aov(w, h)
        gpu_ctx.set_wave_budget_mb(1)
        out, stn = dev.capture(w, h)
        assert stn["bands"] > 1 and st1["bands"] == 1
        assert np.array_equal(out, ref)
        for k in ("primary_rays", "primary_hits", "shadow_rays_traced", "shadow_occluded", "secondary_rays"):
            assert stn[k] == st1[k], k
        aovn = dev.capture_aov(w, h)
        for k in ("rgba", "prim_id", "t", "occl"):
            assert np.array_equal(aovn[k], aov1[k]), k
        sub = np.zeros((h, w, 4), np.uint8)
        for k in range(2):
            dev.capture_subset(k, 2, w, h, sub)
        assert np.array_equal(sub, ref)
    finally:
        gpu_ctx.set_wave_budget_mb(16384)
        dev.destroy()


def test_deferred_bvh(native, oracle, gpu_ctx):
    """LGB_OPT_LAZY_BVH: a large plastic scene is created without its device BVH (grids serve its rays) and renders the film of the
    scene created with it; the first entry point that walks a tree (here: a frame with the camera grid off, lgb_trace_rays) builds
    it, re-orders the primitives and rebuilds the grids -- the film, ids and t stay those of the oracle."""
    sc, (w, h) = scenes.mixed4k(mesh_n=200, nspheres=40000, res=(320, 180), supersampling=1)
    ref = oracle.OracleScene(sc).capture(w, h, aov=True)
    try:
        gpu_ctx.set_lazy_bvh(0)
        eager = native.DeviceScene(gpu_ctx, native.FlatScene(sc))
        film_eager, _ = eager.capture(w, h)
        eager.destroy()
        gpu_ctx.set_lazy_bvh(1)
        dev = native.DeviceScene(gpu_ctx, native.FlatScene(sc))
        film0, st0 = dev.capture(w, h)                       # no tree yet
        out0 = dev.capture_aov(w, h)
        a = parity.aov_report(out0, ref)
        assert a["id_mismatch"] == 0 and a["t_bit_equal"] == a["t_compared"] and a["occl_diff"] == 0, a
        assert np.array_equal(film0, film_eager) and np.array_equal(out0["rgba"], film0)
        gpu_ctx.set_camera_grid(0)                           # this frame walks the BVH: built now
        film1, st1 = dev.capture(w, h)
        gpu_ctx.set_camera_grid(-1)
        film2, _ = dev.capture(w, h)                         # and the grids were rebuilt over the re-ordered primitives
        out2 = dev.capture_aov(w, h)
        a = parity.aov_report(out2, ref)
        assert a["id_mismatch"] == 0 and a["t_bit_equal"] == a["t_compared"] and a["occl_diff"] == 0, a
        assert np.array_equal(film1, film_eager) and np.array_equal(film2, film_eager)
        assert dev.verify()["boxes_ok"] == 1
        dev.destroy()
        dev = native.DeviceScene(gpu_ctx, native.FlatScene(sc))
        rays = oracle.OracleScene(sc).camera_sample(w // 2, h // 2, w, h)
        ids, ts, _, _ = dev.trace_rays(rays)                 # lgb_trace_rays on a scene without a tree
        k = (h // 2 * w + w // 2) * sc.camera.num_samples()
        assert np.array_equal(ids, ref["prim_id"].reshape(-1)[k:k + len(ids)])
        dev.destroy()
    finally:
        gpu_ctx.set_lazy_bvh(-1); gpu_ctx.set_camera_grid(-1)


# matte(sigma > 0), metal, glass, mirror and the Whitted recursion of integrate.rs:69-132
WHITTED = {
    "simplereflect_9spp": lambda: scenes.simplereflect(2, 160),                 # src/examples/simplereflect.rs, depth 4
    "simplereflect_depth1": lambda: scenes.simplereflect(0, 160, recursion=1),
    "simplereflect_depth0": lambda: scenes.simplereflect(0, 96, recursion=0),
    "materials_4spp": lambda: scenes.materials((256, 192), 1),
    "materials_grouped": lambda: scenes.materials((256, 192), 0, grouped=True),
    "materials_depth8": lambda: scenes.materials((160, 120), 0, recursion=8),
    "mixed_whitted_4spp": lambda: scenes.mixed4k(mesh_n=64, nspheres=6000, res=(192, 108), supersampling=1, whitted=True),
}


@pytest.mark.parametrize("name", sorted(WHITTED))
def test_materials_and_whitted_parity(native, oracle, gpu_ctx, name):
    sc, (w, h) = WHITTED[name]()
    o = oracle.OracleScene(sc)
    ref = o.capture(w, h, aov=True, li=True)
    dev = native.DeviceScene(gpu_ctx, native.FlatScene(sc))
    out = dev.capture_aov(w, h, li=True)
    # per-sample radiance: 1e-6 (a reflection off a small sphere multiplies the last-bit difference between the reference's
    # trigonometric sphere normal and the device's), and rare outliers where the reference's own result is rounding noise --
    # a secondary hit 400 units down a ray lands up to 1e-11 off the sphere (cancellation in sphere.rs:36-41), more than the
    # 1.5e-11 offset of surface.rs:168, so whether it shadows itself is decided by the last bit of the incoming ray
    assert radiance_outliers(out["li"], ref["li"], 1e-6) <= 2e-3
    rgba, st = dev.capture(w, h)                                 # the ray trees level by level (default) ...
    gpu_ctx.set_whitted(False)
    try:
        rgba_t, st_t = dev.capture(w, h)                         # ... and one thread per tree (k_secondary): same rays, same film
    finally:
        gpu_ctx.set_whitted(True)
    dev.destroy()
    ft = parity.film_report(rgba_t, rgba)
    assert ft["within_1_frac"] >= 0.999 and ft["identical_frac"] >= 0.999, ft
    assert st_t["secondary_rays"] == st["secondary_rays"], (st_t["secondary_rays"], st["secondary_rays"])
    ft = parity.film_report(rgba_t, ref["rgba"])
    assert ft["alpha_equal"] and ft["within_1_frac"] >= 0.999 and ft["identical_frac"] >= 0.999, ft
    spp = sc.camera.num_samples()

    def retest(pid, i):
        rays = o.camera_sample((i // spp) % w, (i // spp) // w, w, h)
        return o.retest(pid, rays[i % spp, :3], rays[i % spp, 3:])

    a = parity.aov_report(out, ref, retest)                      # the primary hits, as for every other scene
    assert a["mismatches"] == 0 and a["hit_miss_flips"] == 0 and a["near_ties"] <= 1e-4 * a["samples"], a
    assert a["t_bit_equal"] == a["t_compared"] and a["occl_diff"] == 0, a
    f = parity.film_report(rgba, ref["rgba"])                    # ... and the film, which carries every ray below them
    assert f["alpha_equal"] and f["within_1_frac"] >= 0.999, f
    assert f["identical_frac"] >= 0.999, f
    assert np.array_equal(rgba, out["rgba"])
    assert (st["secondary_rays"] > 0) == (sc.recursion > 0), st


def test_whitted_through_every_entry_point(native, oracle, gpu_ctx):
    """The specular ray trees behind the other ways into the path: capture_subset (compact output), tile ranks, forced pixel beams,
    the lazy reference tree of `capture(scene, film)` (taken eagerly for such scenes), the C++ mirror, a scene without lights."""
    import ctypes as C

    import torch

    import lasgun_b200
    sc, (w, h) = scenes.materials((200, 150), 1, grouped=False)
    ref = oracle.OracleScene(sc).capture(w, h)["rgba"]
    dev = native.DeviceScene(gpu_ctx, native.FlatScene(sc))
    try:
        full, st = dev.capture(w, h)
        f = parity.film_report(full, ref)
        assert f["alpha_equal"] and f["within_1_frac"] >= 0.999 and f["identical_frac"] >= 0.999, f
        sub = np.zeros((h, w, 4), np.uint8); dev.capture_subset(2, 5, w, h, sub)
        assert np.array_equal(sub.reshape(-1, 4)[2::5], full.reshape(-1, 4)[2::5]) and not sub.reshape(-1, 4)[0::5].any()
        acc = torch.zeros((h, w, 4), dtype=torch.int32, device="cuda")
        for r in range(3):
            film = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
            dev.capture_device(w, h, film.data_ptr(), rank=r, ranks=3, want_stats=True)      # with stats the call waits for the frame
            acc += film.to(torch.int32)
        assert np.array_equal(acc.cpu().numpy().astype(np.uint8), full)
    finally:
        dev.destroy()
    sc2, (w2, h2) = scenes.simplereflect(1, 96)                          # no transformed group: the primary rays can go through beams
    dev = native.DeviceScene(gpu_ctx, native.FlatScene(sc2))
    try:
        plain, st_p = dev.capture(w2, h2)
        gpu_ctx.set_beams(1)
        beamed, st_b = dev.capture(w2, h2)
        assert st_b["beams"] == 1 and st_p["beams"] == 0 and np.array_equal(beamed, plain) and st_b["secondary_rays"] == st_p["secondary_rays"]
    finally:
        gpu_ctx.set_beams(-1)
        dev.destroy()
    film = Film(w, h)
    lasgun_b200.capture(sc, film, ctx=gpu_ctx)                           # flattens lazily; glass and mirrors make the device ask for the tree at once
    assert np.array_equal(film.pixels(), full)
    host = native.HostScene(sc)
    rgba = np.zeros((h, w, 4), np.uint8)
    assert native.lib().lgh_capture(host.h, w, h, rgba.ctypes.data_as(C.POINTER(C.c_uint8))) == 0, native.lib().lgh_last_error()
    assert np.array_equal(rgba, full)
    sc.lights.clear()                                                    # no light: ambient, background and the ray trees only
    ref = oracle.OracleScene(sc).capture(w, h)["rgba"]
    dev = native.DeviceScene(gpu_ctx, native.FlatScene(sc))
    dark, st = dev.capture(w, h)
    dev.destroy()
    f = parity.film_report(dark, ref)
    assert f["alpha_equal"] and f["within_1_frac"] >= 0.999 and f["identical_frac"] >= 0.999 and st["secondary_rays"] > 0, (f, st)


def test_reference_cornell_grazing_rays(native, oracle, gpu_ctx):
    """src/examples/cornell.rs as shipped (on-axis eye): the rays of the image diagonals run exactly along the edges where
    two transformed walls meet.  There the reference's own f64 node test (cuboid.rs:104-121) can reject, by one rounding, a
    box whose primitive its exact test would hit, so which of two equal-t walls it reports -- or a miss -- depends on its
    tree; the device's boxes are conservative and it reports the watertight hit.  Bound: a handful of samples, each an
    exact-t tie or such an edge, never a different t; the film stays within the north-star tolerance."""
    sc, (w, h) = scenes.cornell_groups((192, 192), 2)
    o = oracle.OracleScene(sc)
    ref = o.capture(w, h, aov=True)
    dev = native.DeviceScene(gpu_ctx, native.FlatScene(sc))
    out = dev.capture_aov(w, h)
    dev.destroy()
    spp = sc.camera.num_samples()

    def retest(pid, i):
        rays = o.camera_sample((i // spp) % w, (i // spp) // w, w, h)
        return o.retest(pid, rays[i % spp, :3], rays[i % spp, 3:])

    a = parity.aov_report(out, ref, retest)
    assert a["mismatches"] == 0, a                               # every differing id is an exact-t tie ...
    assert a["near_ties"] + a["hit_miss_flips"] <= 1e-4 * a["samples"], a   # ... or an edge the reference's box test dropped
    assert a["t_bit_equal"] == a["t_compared"], a
    for i in np.nonzero(out["prim_id"].reshape(-1) != ref["prim_id"].reshape(-1))[0]:
        if ref["prim_id"].reshape(-1)[i] == parity.MISS:           # the device's hit is a real one: the oracle's exact test agrees
            assert np.isfinite(retest(int(out["prim_id"].reshape(-1)[i]), int(i)))
    f = parity.film_report(out["rgba"], ref["rgba"])
    assert f["alpha_equal"] and f["within_1_frac"] >= 0.999, f


FULL = {
    "C1_simple": (lambda: scenes.simple("b", 2), 97),
    "C2_mesh1m": (scenes.mesh1m, 211),
    "C3_cornell": (scenes.cornell, 389),
    "C4_spheres1m": (scenes.spheres1m, 1999),
    "C5_mixed4k": (scenes.mixed4k, 7919),          # 3840x2160 at 16 spp: the oracle renders ~1000 pixels of it (a few seconds per thread)
}


@pytest.mark.parametrize("name", sorted(FULL))
def test_full_size_sampled_against_oracle(native, oracle, gpu_ctx, name):
    """BASELINE.json's full sizes: the oracle renders every n-th pixel (capture_subset(0, n),
    lib.rs:110) of the full-size scene; the device film must agree at those pixels."""
    mk, n = FULL[name]
    sc, (w, h) = mk()
    flat = native.FlatScene(sc)
    dev = native.DeviceScene(gpu_ctx, flat)
    rgba, st = dev.capture(w, h)
    dev.destroy()
    assert (rgba[..., 3] == 255).all() and st["primary_rays"] == w * h * sc.camera.num_samples()
    ref = oracle.OracleScene(sc).capture(w, h, subset=(n, 0, 1))["rgba"].reshape(-1, 4)
    idx = np.arange(0, w * h, n)
    f = parity.film_report(rgba.reshape(-1, 4)[idx][None], ref[idx][None])
    assert f["alpha_equal"] and f["within_1_frac"] >= 0.999 and f["identical_frac"] >= 0.999, f


@pytest.mark.parametrize("name", ["mesh", "spheres", "mixed"])
def test_device_built_bvh(native, gpu_ctx, monkeypatch, name):
    """Scenes of >= 32768 primitives get their BVH built ON the device (lgb_gpubuild.cu): every primitive in exactly one
    leaf, boxes conservative and nested, and the film byte-identical to the one rendered from the host-built tree."""
    sc, (w, h) = {"mesh": lambda: scenes.mesh1m(n=150, res=160), "spheres": lambda: scenes.spheres1m(count=60000, res=160),
                  "mixed": lambda: scenes.mixed4k(mesh_n=120, nspheres=12000, res=(240, 136), supersampling=1)}[name]()
    flat = native.FlatScene(sc)
    dev = native.DeviceScene(gpu_ctx, flat)
    v = dev.verify()
    assert v["ranks_ok"] == 1, "expected the device-side builder for this many primitives"
    assert v["boxes_ok"] == 1 and v["max_leaf"] <= 4 and v["max_depth"] < 64, v
    film_dev, _ = dev.capture(w, h)
    dev.destroy()
    monkeypatch.setenv("LGB_HOST_BUILD", "1")
    host = native.DeviceScene(gpu_ctx, flat)
    vh = host.verify()
    assert vh["ranks_ok"] == 0 and vh["boxes_ok"] == 1
    assert vh["nodes"] == v["nodes"] and abs(vh["sah_cost"] - v["sah_cost"]) <= 1e-6 * vh["sah_cost"]     # the same tree
    film_host, _ = host.capture(w, h)
    host.destroy()
    assert np.array_equal(film_dev, film_host)


def test_device_built_bvh_degenerate(native, gpu_ctx, monkeypatch):
    """The device builder where the SAH has nothing to go on: 12 000 positions, each holding a sphere, the same sphere again and a
    cube (coincident centroids: nodes halved by position, leaves split by type, in the level loop and in k_subtrees alike).  The tree
    must be well formed and the film -- exact-t ties between the twin spheres included -- the one the host-built tree renders."""
    rng = np.random.default_rng(0x5EED0C02)
    sc = Scene()
    sc.set_ambient_light([0.1, 0.1, 0.1])
    cam = sc.set_perspective_camera(45.0)
    cam.look_at([0.0, 0.0, 160.0], [0.0, 0.0, 0.0], [0.0, 1.0, 0.0])
    sc.add_point_light([80.0, 120.0, 150.0], [0.8, 0.8, 0.8], [1.0, 0.0, 0.0])
    pos = -50.0 + 100.0 * rng.random((12000, 3)); rad = 0.6 + 0.6 * rng.random(12000)
    pal = [Material.plastic([0.8, 0.3, 0.2], [0.4, 0.4, 0.4], 0.2), Material.plastic([0.2, 0.7, 0.3], [0.3, 0.3, 0.3], 0.3)]
    sc.root.add_spheres(np.concatenate([pos, pos]), np.concatenate([rad, rad]), pal, np.arange(24000) % 2)
    for p, r in zip(pos, rad):
        sc.root.add_cube((p - 0.4 * r).tolist(), float(0.8 * r), pal[0])
    w = h = 128
    flat = native.FlatScene(sc)
    assert flat.prim_count == 36000
    dev = native.DeviceScene(gpu_ctx, flat)
    v = dev.verify()
    assert v["ranks_ok"] == 1, "expected the device-side builder for this many primitives"
    assert v["boxes_ok"] == 1 and v["max_leaf"] <= 4 and v["max_depth"] < 64, v
    out_dev = dev.capture_aov(w, h)
    film_dev, _ = dev.capture(w, h)
    dev.destroy()
    monkeypatch.setenv("LGB_HOST_BUILD", "1")
    host = native.DeviceScene(gpu_ctx, flat)
    vh = host.verify()
    assert vh["ranks_ok"] == 0 and vh["boxes_ok"] == 1
    out_host = host.capture_aov(w, h)
    film_host, _ = host.capture(w, h)
    host.destroy()
    assert np.array_equal(out_dev["prim_id"], out_host["prim_id"]) and np.array_equal(out_dev["t"], out_host["t"])
    assert np.array_equal(film_dev, film_host)


def test_golden_films(native, gpu_ctx):
    """The committed oracle films of tests/golden/films.npz (a small version of every config + two nested-group scenes): the device
    film must equal them byte for byte -- no oracle is run here."""
    import importlib.util
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    gold = np.load(os.path.join(here, "golden", "films.npz"))
    for name, mk in mg.CASES.items():
        sc, (w, h) = mk()
        dev = native.DeviceScene(gpu_ctx, native.FlatScene(sc))
        rgba, _ = dev.capture(w, h)
        dev.destroy()
        if name.startswith("F4_"):      # rays below specular hits: the throughput form reorders f64 products (k_secondary), a byte may move by 1
            f = parity.film_report(rgba, gold[name])
            assert f["alpha_equal"] and f["within_1_frac"] >= 0.999 and f["identical_frac"] >= 0.999, (name, f)
        else:
            assert np.array_equal(rgba, gold[name]), name


def _variant(name):
    """Corners of the configuration space the five configs do not reach."""
    if name == "orthographic_9spp":                      # camera.rs:126-128, 168-173: origins move across the image plane
        sc, _ = scenes.mixed4k(mesh_n=48, nspheres=3000, res=(96, 54), supersampling=2)
        cam = sc.set_orthographic_camera(700.0)
        cam.look_at([40.0, 90.0, 520.0], [0.0, 0.0, 0.0], [0.0, 1.0, 0.0]); cam.set_supersampling(2)
        return sc, (96, 54)
    if name == "spp_289":                                # more samples than a block holds: radiance buffer + k_resolve
        sc, _ = scenes.simple("b", 0, 16)
        sc.camera.set_supersampling(16)
        return sc, (20, 12)
    if name == "lights_6":
        sc, _ = scenes.spheres1m(count=4000, res=64)
        for i in range(3):
            sc.add_point_light([300.0 * (i - 1), -700.0, 900.0 - 200.0 * i], [0.2, 0.3, 0.25], [1.0, 0.0, 0.0])
        sc.camera.set_supersampling(1)
        return sc, (80, 64)
    raise KeyError(name)


@pytest.mark.parametrize("beams", [0, 1])
@pytest.mark.parametrize("name", ["orthographic_9spp", "spp_289", "lights_6"])
def test_variants(native, oracle, gpu_ctx, name, beams):
    sc, (w, h) = _variant(name)
    dev = native.DeviceScene(gpu_ctx, native.FlatScene(sc))
    try:
        gpu_ctx.set_beams(beams)
        out = dev.capture_aov(w, h)
        rgba, st = dev.capture(w, h)
    finally:
        gpu_ctx.set_beams(-1)
        dev.destroy()
    ref = oracle.OracleScene(sc).capture(w, h, aov=True)
    a = parity.aov_report(out, ref)
    assert a["id_mismatch"] == 0 and a["t_bit_equal"] == a["t_compared"] and a["occl_diff"] == 0, a
    f = parity.film_report(rgba, ref["rgba"])
    assert f["alpha_equal"] and f["within_1_frac"] >= 0.999 and f["identical_frac"] >= 0.999, f
    assert np.array_equal(rgba, out["rgba"]) and st["beams"] == (1 if beams and sc.camera.num_samples() >= 4 else 0)


@pytest.mark.parametrize("name", ["mixed_16spp", "simple_9spp", "cornell_4spp", "coincident_4spp", "spheres_9spp_ragged"])
def test_pixel_beams(native, oracle, gpu_ctx, name):
    """LGB_OPT_BEAMS: the sample rays of a pixel share one bundle traversal (k_beam) and walk its leaf list (k_leafp), with a
    per-ray fallback.  Ids, t, occlusion bits and film must be the oracle's -- and identical to the per-ray path."""
    sc, (w, h) = {"mixed_16spp": lambda: scenes.mixed4k(mesh_n=64, nspheres=6000, res=(96, 54), supersampling=3),
                  "simple_9spp": lambda: scenes.simple("b", 2, 96),
                  "cornell_4spp": lambda: scenes.cornell((160, 90), 1),
                  "coincident_4spp": lambda: scenes.coincident_planes(nspheres=3000, res=(96, 72), supersampling=1),
                  "spheres_9spp_ragged": lambda: scenes.spheres1m(count=30000, res=77)[0:1] + ((77, 45),)}[name]()
    if name == "spheres_9spp_ragged":
        sc.camera.set_supersampling(2)
    dev = native.DeviceScene(gpu_ctx, native.FlatScene(sc))
    try:
        gpu_ctx.set_beams(1)
        out = dev.capture_aov(w, h)
        film_on, st_on = dev.capture(w, h)
        sub_on = np.zeros((h, w, 4), np.uint8); dev.capture_subset(1, 3, w, h, sub_on)        # capture_subset through the beams too
        gpu_ctx.set_beams(0)
        film_off, st_off = dev.capture(w, h)
        sub_off = np.zeros((h, w, 4), np.uint8); dev.capture_subset(1, 3, w, h, sub_off)
    finally:
        gpu_ctx.set_beams(-1)
        dev.destroy()
    assert np.array_equal(film_on, film_off) and np.array_equal(out["rgba"], film_on)
    assert np.array_equal(sub_on, sub_off) and np.array_equal(sub_on.reshape(-1, 4)[1::3], film_on.reshape(-1, 4)[1::3]) and not sub_on.reshape(-1, 4)[0::3].any()
    for k in ("primary_rays", "primary_hits", "shadow_rays", "shadow_rays_traced", "shadow_occluded"):
        assert st_on[k] == st_off[k], k
    ref = oracle.OracleScene(sc).capture(w, h, aov=True)
    a = parity.aov_report(out, ref)
    assert a["id_mismatch"] == 0 and a["t_bit_equal"] == a["t_compared"] and a["occl_diff"] == 0, a
    assert np.array_equal(film_on, ref["rgba"])


@pytest.mark.parametrize("case", ["no_ties", "ties_listed", "ties_whole_frame", "small_scene"])
def test_lazy_reference_tree(native, oracle, gpu_ctx, monkeypatch, case):
    """Lazy reference tree (lasgun_b200.h): the reference BVH is not built unless a closest-hit ray meets two primitives at
    bit-identical t; then the device fetches it through the callback, builds the rank tables and re-traces the tied
    slots.  Ids, t and film must be the oracle's either way."""
    if case == "no_ties":
        sc, (w, h) = scenes.mixed4k(mesh_n=130, nspheres=8000, res=(192, 108), supersampling=1)
    elif case == "small_scene":
        sc, (w, h) = scenes.simple("b", 1, 96)                 # below the device-builder threshold: the tree is fetched at creation
    else:
        sc, (w, h) = scenes.coincident_planes()
    if case == "ties_whole_frame":
        monkeypatch.setenv("LGB_TIE_CAP", "16")
    flat = native.FlatScene(sc, lazy=True)
    assert not flat.tree_built
    dev = native.DeviceScene(gpu_ctx, flat)
    assert flat.tree_built == (case == "small_scene")
    out = dev.capture_aov(w, h)
    rgba, _ = dev.capture(w, h)
    assert flat.tree_built == (case != "no_ties")             # built exactly when a tie (or a small scene) needed it
    dev.destroy()
    ref = oracle.OracleScene(sc).capture(w, h, aov=True)
    a = parity.aov_report(out, ref)
    assert a["id_mismatch"] == 0 and a["t_bit_equal"] == a["t_compared"] and a["occl_diff"] == 0, a
    assert np.array_equal(out["rgba"], ref["rgba"]) and np.array_equal(rgba, ref["rgba"])
    if case.startswith("ties"):                                # the ties are real: an eager scene resolves thousands of them too
        assert (ref["prim_id"].reshape(-1) < 4).sum() > 1000


@pytest.mark.parametrize("name", ["plain", "instanced"])
def test_scene_export_import(native, gpu_ctx, name):
    """Multi-GPU replication path on one GPU: the exported arena, copied byte for byte into memory the 'other rank' owns,
    imports into a scene that renders the same film (lgb_scene_export / lgb_scene_import)."""
    import torch
    from lasgun_b200 import multi
    sc, (w, h) = scenes.mixed4k(mesh_n=64, nspheres=5000, res=(192, 108), supersampling=1) if name == "plain" else scenes.nested_groups((192, 144), 1)
    dev = native.DeviceScene(gpu_ctx, native.FlatScene(sc))
    film, _ = dev.capture(w, h)
    layout, ptr, nbytes = dev.export()
    assert len(layout) == native.lib().lgb_scene_layout_bytes() and 0 < nbytes <= dev.device_bytes      # (device_bytes also counts the light grids)
    arena = torch.as_tensor(multi._DevicePointer(ptr, nbytes), device="cuda").clone()
    dev.destroy()
    twin = native.DeviceScene.adopt(gpu_ctx, layout, arena.data_ptr(), dev.spp, keep=arena)
    film2, _ = twin.capture(w, h)
    twin.destroy()
    assert np.array_equal(film, film2)
    with pytest.raises(native.LasgunError):
        native.DeviceScene.adopt(gpu_ctx, b"\0" * len(layout), arena.data_ptr(), 1)


def test_capture_subset_union_equals_capture(native, gpu_ctx):
    """capture_subset(k, n) for k in 0..n tiles the film exactly (lib.rs:114-141) and leaves other pixels untouched."""
    sc, (w, h) = scenes.simple("b", 1, 160)
    dev = native.DeviceScene(gpu_ctx, native.FlatScene(sc))
    full, _ = dev.capture(w, h)
    film = np.full((h, w, 4), 7, np.uint8)
    n = 5
    for k in range(n):
        before = film.copy()
        dev.capture_subset(k, n, w, h, film)
        idx = np.arange(k, w * h, n)
        mask = np.ones(w * h, bool); mask[idx] = False
        assert np.array_equal(film.reshape(-1, 4)[mask], before.reshape(-1, 4)[mask])
    dev.destroy()
    assert np.array_equal(film, full)


def test_tile_ranks_partition_the_film(native, gpu_ctx):
    """Multi-GPU sharding on one GPU: rank r of n renders disjoint macro tiles whose union is the 1-rank film."""
    import torch
    sc, (w, h) = scenes.mixed4k(mesh_n=48, nspheres=3000, res=(300, 170), supersampling=1)
    dev = native.DeviceScene(gpu_ctx, native.FlatScene(sc))
    full, _ = dev.capture(w, h)
    for ranks in (2, 4, 8):
        acc = torch.zeros((h, w, 4), dtype=torch.int32, device="cuda")
        for r in range(ranks):
            film = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
            st = dev.capture_device(w, h, film.data_ptr(), rank=r, ranks=ranks, want_stats=True)
            assert st["stack_overflow"] == 0
            assert int(((film[..., 3] != 0) & (acc[..., 3] != 0)).sum()) == 0      # disjoint
            acc += film.to(torch.int32)
        assert np.array_equal(acc.cpu().numpy().astype(np.uint8), full)
    dev.destroy()


@pytest.mark.parametrize("spp_root", [1, 3])
def test_capture_readback_in_slices(native, gpu_ctx, monkeypatch, spp_root):
    """lgb_capture of a film of >= 8 MB: the shade kernel runs in slices and the finished rows travel to the host behind each
    (run_capture), through a pinned bounce buffer when the caller's film is pageable.  The film must be the one the device holds,
    byte for byte -- pageable and pinned destination, a height that is no multiple of the 32-row macro tiles, 1 and 9 spp (9: a
    block's pixels do not divide a macro tile) -- and the one the unsliced read-back (LGB_NO_FILM_OVERLAP) delivers."""
    import torch
    w, h = 2048, 1100
    sc, _ = scenes.mixed4k(mesh_n=48, nspheres=3000, res=(w, h), supersampling=spp_root)
    dev = native.DeviceScene(gpu_ctx, native.FlatScene(sc))
    film = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    dev.capture_device(w, h, film.data_ptr(), want_stats=True)
    want = film.cpu().numpy()
    pageable = np.full((h, w, 4), 7, np.uint8)
    _, st = dev.capture(w, h, out=pageable)
    assert np.array_equal(pageable, want)
    assert st["kernel_launches"] >= 4            # the slices are launches of their own
    pinned = torch.full((h, w, 4), 9, dtype=torch.uint8).pin_memory()
    dev.capture(w, h, out=pinned.numpy())
    assert np.array_equal(pinned.numpy(), want)
    monkeypatch.setenv("LGB_NO_FILM_OVERLAP", "1")
    plain = np.full((h, w, 4), 5, np.uint8)
    dev.capture(w, h, out=plain)
    assert np.array_equal(plain, want)
    monkeypatch.setenv("LGB_NO_FILM_BOUNCE", "1")
    plain2 = np.full((h, w, 4), 3, np.uint8)
    dev.capture(w, h, out=plain2)
    assert np.array_equal(plain2, want)
    dev.destroy()


def test_public_api_capture(native, oracle, gpu_ctx):
    """The call a user makes: lasgun_b200.capture(scene, film) == oracle film."""
    import lasgun_b200
    sc, (w, h) = scenes.cornell((240, 136), 1)
    film = Film(w, h)
    lasgun_b200.capture(sc, film, ctx=gpu_ctx)
    ref = oracle.OracleScene(sc).capture(w, h)["rgba"]
    f = parity.film_report(film.pixels(), ref)
    assert f["alpha_equal"] and f["within_1_frac"] >= 0.999, f
    # C++ mirror entry (include/lasgun_host.hpp: lasgun::capture)
    host = native.HostScene(sc)
    rgba = np.zeros((h, w, 4), np.uint8)
    import ctypes as C
    rc = native.lib().lgh_capture(host.h, w, h, rgba.ctypes.data_as(C.POINTER(C.c_uint8)))
    assert rc == 0, native.lib().lgh_last_error()
    assert np.array_equal(rgba, film.pixels())


def test_abi_rejects_what_the_path_does_not_cover(native, gpu_ctx):
    import ctypes as C
    sc, _ = scenes.simple("a", 0, 32)
    flat = native.FlatScene(sc)
    L = native.lib()
    h = C.c_void_p()
    abi = flat.desc.abi_version
    flat.desc.abi_version = 99
    assert L.lgb_scene_create(gpu_ctx.h, C.byref(flat.desc), C.byref(h)) == native.LGB_ERR_INVALID
    flat.desc.abi_version = abi
    kind_addr = flat.desc.materials + 8 * 8               # lgb_material.kind
    old = C.c_uint32.from_address(kind_addr).value
    C.c_uint32.from_address(kind_addr).value = 7          # no such Material variant
    assert L.lgb_scene_create(gpu_ctx.h, C.byref(flat.desc), C.byref(h)) == native.LGB_ERR_INVALID
    assert b"material" in L.lgb_last_error(gpu_ctx.h)
    C.c_uint32.from_address(kind_addr).value = 3          # glass, deeper than the per-ray stack of k_secondary
    flat.desc.recursion = 13
    assert L.lgb_scene_create(gpu_ctx.h, C.byref(flat.desc), C.byref(h)) == native.LGB_ERR_UNSUPPORTED
    flat.desc.recursion = 3
    C.c_uint32.from_address(kind_addr).value = old
    assert L.lgb_scene_create(gpu_ctx.h, C.byref(flat.desc), C.byref(h)) == 0
    L.lgb_scene_destroy(h)
