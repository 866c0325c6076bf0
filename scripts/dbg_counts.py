import os, sys
sys.path.insert(0, ".")
import torch
from lasgun_b200 import _native as N, scenes
name = sys.argv[1] if len(sys.argv) > 1 else "mixed4k"
sc, (w, h) = scenes.CONFIGS[name]()
ctx = N.Context(0)
dev = N.DeviceScene(ctx, N.FlatScene(sc))
film = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
ctx.set_count_work(True)
st = dev.capture_device(w, h, film.data_ptr(), want_stats=True)
pr = st["primary_rays"]
print(name, "beams" if os.environ.get("LGB_BEAMS") else "plain", "per primary ray: node fetches %.1f, filter tests %s, exact %s" % (
    st["primary_node_tests"] / pr, [round(v / pr, 2) for v in st["primary_filter_tests"]], [round(v / pr, 3) for v in st["primary_exact_tests"]]),
    "kernel_ms", [round(v, 2) for v in st["kernel_ms"]])
