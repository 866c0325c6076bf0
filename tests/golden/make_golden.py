"""Generates tests/golden/films.npz: small films of every config rendered by the CPU oracle (oracle/lasgun_oracle.cpp).

    python tests/golden/make_golden.py

The Rust reference cannot be built in this image (no cargo), so these are ORACLE outputs, pinned here so that (a) a change
of the oracle that alters a pixel is noticed (tests/test_oracle_known_answers.py::test_oracle_reproduces_golden_films) and
(b) the device path is checked against bytes that travel with the repository (tests/test_gpu_parity.py::test_golden_films).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lasgun_b200 import scenes  # noqa: E402

CASES = {
    "C1_simple_9spp": lambda: scenes.simple("b", 2, 64),
    "C2_mesh": lambda: scenes.mesh1m(n=60, res=64),
    "C3_cornell_4spp": lambda: scenes.cornell((80, 45), 1),
    "C4_spheres": lambda: scenes.spheres1m(count=20000, res=64),
    "C5_mixed_16spp": lambda: scenes.mixed4k(mesh_n=40, nspheres=4000, res=(64, 36), supersampling=3),
    "F1_nested_groups": lambda: scenes.nested_groups((80, 60), 1),
    "F1_cornell_groups_offaxis_9spp": lambda: scenes.cornell_groups((48, 48), 2, eye=(0.21, 0.13, 5.0)),
    "F4_simplereflect_depth4_4spp": lambda: scenes.simplereflect(1, 64),
    "F4_materials": lambda: scenes.materials((96, 72), 0),
    "F4_materials_grouped_4spp": lambda: scenes.materials((64, 48), 1, grouped=True),
}

if __name__ == "__main__":
    from oracle import pyoracle as po
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "films.npz")
    old = dict(np.load(path)) if os.path.exists(path) else {}
    out = {}
    for name, mk in CASES.items():
        sc, (w, h) = mk()
        out[name] = po.OracleScene(sc).capture(w, h)["rgba"]
        print(name, out[name].shape, int(out[name].astype(np.uint64).sum()), "unchanged" if name in old and np.array_equal(old[name], out[name]) else "NEW/CHANGED")
    np.savez_compressed(path, **out)
