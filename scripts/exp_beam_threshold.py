"""Where do the pixel beams (primary + shadow) pay?  ms/frame with LGB_OPT_BEAMS off / on over sample counts and BVH sizes."""
import sys
sys.path.insert(0, ".")
import torch
from lasgun_b200 import _native as N, scenes
CASES = (("simple 9spp", lambda: scenes.simple("b", 2)), ("cornell 4spp", scenes.cornell),
         ("mixed4k 4spp", lambda: scenes.mixed4k(supersampling=1)), ("mixed4k 9spp", lambda: scenes.mixed4k(supersampling=2)),
         ("mixed 1080p 16spp, 3 k nodes", lambda: scenes.mixed4k(mesh_n=40, nspheres=1500, res=(1920, 1080))),
         ("mixed 1080p 16spp, 30 k nodes", lambda: scenes.mixed4k(mesh_n=120, nspheres=15000, res=(1920, 1080))))
for name, mk in CASES:
    sc, (w, h) = mk()
    ctx = N.Context(0); dev = N.DeviceScene(ctx, N.FlatScene(sc))
    film = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    out = []
    for b in (0, 1):
        ctx.set_beams(b)
        t = [dev.capture_device(w, h, film.data_ptr(), want_stats=True)["render_ms"] for _ in range(4)]
        out.append(round(min(t[1:]), 3))
    print(f"{name:32s} beams off {out[0]:8.3f}  on {out[1]:8.3f}  nodes {N.lib().lgb_scene_node_count(dev.h)}")
    dev.destroy(); ctx.close()
