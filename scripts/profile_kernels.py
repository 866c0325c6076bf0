"""Per-kernel times (lgb_capture_profile) of one workload, timing pass + counting pass."""
import sys
import torch
sys.path.insert(0, ".")
from lasgun_b200 import _native as N, scenes
name = sys.argv[1] if len(sys.argv) > 1 else "mixed4k"
sc, (w, h) = scenes.CONFIGS[name]()
ctx = N.Context(0)
dev = N.DeviceScene(ctx, N.FlatScene(sc))
film = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
for _ in range(2):
    dev.capture_device(w, h, film.data_ptr(), want_stats=True)
best = None
for _ in range(3):
    k, st = dev.capture_profile(w, h, film.data_ptr())
    if best is None or sum(x["ms"] for x in k) < sum(x["ms"] for x in best):
        best = k
ctx.set_count_work(True)
kc, stc = dev.capture_profile(w, h, film.data_ptr())
ctx.set_count_work(False)
tot = sum(x["ms"] for x in best)
for a, b in zip(best, kc):
    print(f"{a['name']:28s} {a['ms']:7.3f} ms {100 * a['ms'] / tot:5.1f}%  nodes {b['node_tests']:>12d} filt {b['filter_tests']} exact {b['exact_tests']} prim {b['primary_rays']} shad {b['shadow_rays']} occl {b['shadow_occluded']}")
print("sum", round(tot, 3), "frame (overlapped)", round(dev.capture_device(w, h, film.data_ptr(), want_stats=True)["render_ms"], 3))
