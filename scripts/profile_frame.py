"""Renders a few frames of one workload (target for ncu)."""
import sys
import torch
sys.path.insert(0, ".")
from lasgun_b200 import _native as N, scenes
name = sys.argv[1] if len(sys.argv) > 1 else "mixed4k"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2
scale = sys.argv[3] if len(sys.argv) > 3 else "full"
if name == "mixed4k" and scale != "full":
    sc, (w, h) = scenes.mixed4k(res=(1920, 1080), supersampling=1)
else:
    sc, (w, h) = scenes.CONFIGS[name]()
import os
ctx = N.Context(0)
if os.environ.get("LGB_SIDE") == "0":
    ctx.set_side_streams(False)
if os.environ.get("LGB_WHITTED") == "0":
    ctx.set_whitted(False)       # one thread per specular ray tree (k_secondary) instead of the level-by-level wavefront
dev = N.DeviceScene(ctx, N.FlatScene(sc))
film = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
for i in range(frames):
    st = dev.capture_device(w, h, film.data_ptr(), want_stats=True)
    print(name, w, h, "frame", i, "render_ms", round(st["render_ms"], 3), {k: st[k] for k in ("primary_hits", "shadow_rays", "shadow_rays_traced", "shadow_occluded", "shadow_cache_hits", "secondary_rays")})
