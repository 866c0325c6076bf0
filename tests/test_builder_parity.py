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
