"""A frame far beyond one band: mixed4k's scene at 7680x4320 and 64 spp (2.1 G samples), rendered in bands of the default budget."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from lasgun_b200 import _native as N, scenes
res = (7680, 4320) if len(sys.argv) < 2 else tuple(int(v) for v in sys.argv[1].split("x"))
ss = int(sys.argv[2]) if len(sys.argv) > 2 else 7          # supersampling base: (7 + 1)^2 = 64 spp
sc, (w, h) = scenes.mixed4k(res=res, supersampling=ss)
ctx = N.Context(0)
dev = N.DeviceScene(ctx, N.FlatScene(sc))
film = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
for i in range(2):
    t0 = time.time()
    st = dev.capture_device(w, h, film.data_ptr(), want_stats=True)
    print(f"{w}x{h} spp {sc.camera.num_samples()} frame {i}: render_ms {st['render_ms']:.1f} bands {st['bands']} primary {st['primary_rays']} hits {st['primary_hits']} "
          f"shadow {st['shadow_rays_traced']} Mrays/s {(st['primary_rays'] + st['shadow_rays']) / st['render_ms'] / 1e3:.0f} wall {time.time() - t0:.2f}s "
          f"mem {torch.cuda.mem_get_info()[0] / 2**30:.1f} GiB free")
a = film.cpu().numpy()
print("alpha all 255:", bool((a[..., 3] == 255).all()), "mean rgb", a[..., :3].mean(axis=(0, 1)).round(2).tolist())
# the centre 3840x2160 window of the 8K frame at 64 spp vs the 4K frame? different sampling: just report
