"""Host HLBVH build + flatten (product, C++) against the oracle's restatement of bvh.rs:135-453:
node bounds, (leaf, a, b) and order[] must be identical, level by level."""
import numpy as np
import pytest

from lasgun_b200 import scenes
from lasgun_b200.api import Aggregate, Material, Scene


def oracle_levels(o, path=()):
    yield path
    _, _, order = o.bvh_dump(path)
    for i in range(len(order)):
        try:
            o.bvh_dump(path + (i,))
        except KeyError:
            continue
        yield from oracle_levels(o, path + (i,))


CASES = {
    "simple": lambda: scenes.simple()[0],
    "cornell": lambda: scenes.cornell()[0],
    "mesh": lambda: scenes.mesh1m(n=90)[0],
    "spheres": lambda: scenes.spheres1m(count=30000)[0],
    "mixed": lambda: scenes.mixed4k(mesh_n=70, nspheres=6000)[0],
    "cornell_groups": lambda: scenes.cornell_groups()[0],                 # transformed groups: transform_bounds feeds the parent's tree
    "nested_groups": lambda: scenes.nested_groups()[0],
    "nested_groups_root": lambda: scenes.nested_groups(transformed_root=True)[0],
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_reference_tree_identical(native, oracle, name):
    sc = CASES[name]()
    flat = native.FlatScene(sc, keep_levels=True)
    o = oracle.OracleScene(sc)
    assert flat.prim_count == o.prim_count
    paths = list(oracle_levels(o))
    assert flat.level_count() == len(paths)
    for li, path in enumerate(paths):
        b, meta, order = o.bvh_dump(path)
        fb, fmeta, forder, _ = flat.level(li)
        assert np.array_equal(b, fb) and np.array_equal(meta, fmeta) and np.array_equal(order, forder), (name, path)


def _walk(nodes, refs, i, depth, out):
    n = nodes[i]
    if n["b"] & 0x80000000:
        cnt = int(n["b"] & 0x7FFFFFFF)
        out["leaves"].append((int(n["a"]), cnt))
        out["depth"] = max(out["depth"], depth)
        return
    assert n["b"] <= 2 and n["a"] > i + 1
    for c in (i + 1, int(n["a"])):
        assert (nodes[c]["lo"] >= n["lo"]).all() and (nodes[c]["hi"] <= n["hi"]).all()
        _walk(nodes, refs, c, depth + 1, out)


def test_flat_scene_structure(native):
    """Every primitive is referenced exactly once; child boxes nest; leaves keep the reference's size bound."""
    sc = scenes.mixed4k(mesh_n=40, nspheres=3000)[0]
    flat = native.FlatScene(sc)
    nodes, refs = flat.nodes(), flat.prim_refs()
    d = flat.desc
    assert len(refs) == d.n_spheres + d.n_cuboids + d.n_triangles + d.n_instances
    assert len(np.unique(refs)) == len(refs)
    out = {"leaves": [], "depth": 0}
    _walk(nodes, refs, 0, 0, out)
    assert max(c for _, c in out["leaves"]) <= 254          # bvh.rs:289


def test_reference_hazards_are_errors(native):
    """Q9: empty aggregates -> loud errors.  Every material variant flattens."""
    sc = Scene()
    with pytest.raises(native.LasgunError):
        native.FlatScene(sc)
    sc = Scene(); sc.root.add_sphere([0, 0, 0], 1.0, Material.glass([1, 1, 1], [1, 1, 1], 1.5))
    sc.root.add_sphere([3, 0, 0], 1.0, Material.metal([0.2, 0.9, 1.1], [3.9, 2.4, 2.2], 0.08, 0.45)); sc.set_max_recursion_depth(5)
    flat = native.FlatScene(sc)
    assert flat.desc.n_materials == 2 and flat.desc.recursion == 5
    sc = Scene(); g = Aggregate(); g.translate([1, 0, 0]); sc.root.add_group(g)      # empty nested aggregate
    with pytest.raises(native.LasgunError):
        native.FlatScene(sc)


def expected_ids(sc):
    """Canonical primitive ids: a running counter in construction order -- `from_aggregate` visits `contents` in order,
    a mesh contributes its triangles in file order, a nested group its own contents at the place it was added."""
    spheres, boxes, tri_first = {}, {}, {}
    nxt = [0]

    def walk(agg):
        for item in agg.contents:
            k = item[0]
            if k == "sphere":
                spheres.setdefault((tuple(np.asarray(item[1], float)), float(item[2])), []).append(nxt[0]); nxt[0] += 1
            elif k == "spheres":
                for c, r in zip(np.asarray(item[1], float), np.asarray(item[2], float)):
                    spheres.setdefault((tuple(c), float(r)), []).append(nxt[0]); nxt[0] += 1
            elif k == "cube":
                o = np.asarray(item[1], float)
                boxes.setdefault((tuple(o), tuple(o + item[2])), []).append(nxt[0]); nxt[0] += 1
            elif k == "box":
                a, b = np.asarray(item[1], float), np.asarray(item[2], float)
                boxes.setdefault((tuple(np.minimum(a, b)), tuple(np.maximum(a, b))), []).append(nxt[0]); nxt[0] += 1
            elif k == "mesh":
                tri_first.setdefault(id(item), []).append(nxt[0]); nxt[0] += len(sc.meshes[item[1].index].faces)
            elif k == "group":
                walk(item[1])
    walk(sc.root)
    return spheres, boxes, sorted(v for vs in tri_first.values() for v in vs), nxt[0]


def many_chunks_scene():
    """More contents than one flatten chunk holds (8192), with groups and a mesh in between: the parallel id / slot assignment
    must hand the nested levels their id ranges at the right places."""
    sc = Scene()
    sc.set_perspective_camera(45.0).look_at([0, 0, 60], [0, 0, 0], [0, 1, 0])
    rng = np.random.default_rng(5)
    mat = Material.plastic([0.5, 0.5, 0.5], [0.2, 0.2, 0.2], 0.3)
    mesh = sc.add_obj(scenes.mesh_grid(6, 1.0))
    for i in range(30000):
        c = rng.uniform(-20, 20, 3)
        if i % 7 == 3:
            sc.root.add_cube(list(c), 0.3 + (i % 5) * 0.01, mat)
        else:
            sc.root.add_sphere(list(c), 0.1 + (i % 11) * 0.01, mat)
        if i in (5, 8191, 8192, 12000, 24575, 29999):
            g = Aggregate(); g.translate([float(i % 13), 0.0, 0.0])
            g.add_sphere([0.0, 0.0, float(i)], 0.5, mat); g.add_obj_of(mesh, mat); g.add_cube([1.0, 2.0, float(i)], 0.25, mat)
            sc.root.add_group(g)
    return sc


@pytest.mark.parametrize("name", ["nested_groups", "mixed", "many_chunks"])
def test_canonical_ids_follow_construction_order(native, name):
    import ctypes as C
    sc = many_chunks_scene() if name == "many_chunks" else CASES[name]()
    flat = native.FlatScene(sc, lazy=(name != "nested_groups"))
    d = flat.desc
    spheres, boxes, tri_first, total = expected_ids(sc)
    assert flat.prim_count == total

    def arr(ptr, n, ty, width=1):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ty)), shape=(n * width,)).reshape(n, width) if n else np.zeros((0, width))
    sp = arr(d.spheres, d.n_spheres, C.c_double, 4); sid = arr(d.sphere_id, d.n_spheres, C.c_uint32)[:, 0]
    for row, i in zip(sp, sid):
        assert int(i) in spheres[(tuple(row[:3]), float(row[3]))]          # (the same sphere may be added twice: a list of ids)
    cb = arr(d.cuboids, d.n_cuboids, C.c_double, 6); cid = arr(d.cuboid_id, d.n_cuboids, C.c_uint32)[:, 0]
    for row, i in zip(cb, cid):
        assert int(i) in boxes[(tuple(row[:3]), tuple(row[3:]))]
    tid = arr(d.triangle_id, d.n_triangles, C.c_uint32)[:, 0]
    used = set(sid.tolist()) | set(cid.tolist()) | set(tid.tolist())
    assert len(used) == total and used == set(range(total))                 # every id exactly once
    for first in tri_first:                                                   # a mesh's triangles take consecutive ids from its place in the order
        assert first in set(tid.tolist())


def test_mesh_index_out_of_range_is_an_error(native):
    """api.ObjData refuses bad indices; arrays that went bad afterwards are caught by the C++ flatten (in its parallel loop)."""
    from lasgun_b200.api import ObjData
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    with pytest.raises(ValueError):
        ObjData(pos, np.array([[0, 1, 5]], np.uint32))
    for what in ("face", "normal"):
        sc = Scene()
        m = ObjData(pos, np.array([[0, 1, 2]], np.uint32), np.array([[0, 0, 1]], np.float32), np.array([[0, 0, 0]], np.uint32))
        sc.root.add_obj_of(sc.add_obj(m), Material.default())
        (m.faces if what == "face" else m.normal_faces)[0, 2] = 7
        with pytest.raises(native.LasgunError):
            native.FlatScene(sc)
