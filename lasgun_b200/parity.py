"""Parity report between a device capture and a reference capture of the same scene
(SURVEY.md Appendix E).  Pure numpy on caller-supplied arrays: the checker (oracle) is passed in
by tests / bench, never imported here."""
from __future__ import annotations

import numpy as np

MISS = 0xFFFFFFFF


def film_report(dev_rgba, ref_rgba):
    d = np.abs(dev_rgba.astype(np.int16) - ref_rgba.astype(np.int16))
    rgb = d[..., :3].max(axis=-1)
    npx = rgb.size
    return {
        "pixels": int(npx),
        "alpha_equal": bool(np.array_equal(dev_rgba[..., 3], ref_rgba[..., 3])),
        "identical_frac": float((rgb == 0).sum() / npx),
        "within_1_frac": float((rgb <= 1).sum() / npx),
        "max_abs": int(rgb.max()) if npx else 0,
        "hist": {int(k): int(v) for k, v in zip(*np.unique(rgb, return_counts=True))},
    }


def aov_report(dev, ref, retest=None, rays=None, tie_tol=1e-6):
    """dev/ref: dicts with prim_id, t (per sample) and optionally occl.  `retest(prim_id, sample_index)`
    returns the f64 t of the DEVICE's primitive for that sample's ray (oracle arithmetic)."""
    did, rid = dev["prim_id"], ref["prim_id"]
    dt, rt = dev["t"], ref["t"]
    n = did.size
    same = did == rid
    rep = {"samples": int(n), "id_equal": int(same.sum()), "id_mismatch": int((~same).sum())}
    both_hit = same & (rid != MISS)
    rep["t_bit_equal"] = int((dt[both_hit] == rt[both_hit]).sum())
    rep["t_compared"] = int(both_hit.sum())
    with np.errstate(invalid="ignore"):
        rel = np.abs(dt[both_hit] - rt[both_hit]) / np.maximum(1.0, np.abs(rt[both_hit]))
    rep["t_max_rel"] = float(rel.max()) if rel.size else 0.0
    bad = np.nonzero(~same)[0]
    near, flips, wrong = 0, 0, 0
    examples = []
    for i in bad[:10000]:
        if did[i] == MISS or rid[i] == MISS:
            flips += 1
            kind = "hit/miss flip"
        else:
            t_dev64 = retest(int(did[i]), int(i)) if retest else dt[i]
            if np.isfinite(t_dev64) and abs(t_dev64 - rt[i]) <= tie_tol * max(1.0, abs(rt[i])):
                near += 1
                kind = "near-tie"
            else:
                wrong += 1
                kind = "MISMATCH"
        if len(examples) < 8:
            examples.append((int(i), int(did[i]), int(rid[i]), float(dt[i]), float(rt[i]), kind))
    rep.update(near_ties=near, hit_miss_flips=flips, mismatches=wrong, examples=examples)
    if dev.get("occl") is not None and ref.get("occl") is not None:
        rep["occl_diff"] = int((dev["occl"][same] != ref["occl"][same]).sum())
    return rep
