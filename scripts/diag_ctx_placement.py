"""Does k_cprimary's process-to-process spread (6.8 - 8.6 ms) follow the context's buffers?  One process, several contexts in turn: each
allocates its own wavefront buffers, scene arena and grids; an optional dummy allocation in between shifts where they land."""
import os, sys
import torch
sys.path.insert(0, ".")
from lasgun_b200 import _native as N, scenes
sc, (w, h) = scenes.CONFIGS["mixed4k"]()
hs = N.HostScene(sc)
film = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
keep = []
for trial in range(int(os.environ.get("TRIALS", "8"))):
    ctx = N.Context(0)
    dev = N.DeviceScene(ctx, N.FlatScene(hs))
    for _ in range(2):
        dev.capture_device(w, h, film.data_ptr(), want_stats=True)
    best = None
    for _ in range(3):
        k, st = dev.capture_profile(w, h, film.data_ptr())
        if best is None or sum(x["ms"] for x in k) < sum(x["ms"] for x in best):
            best = k
    print(f"context {trial}:", " ".join(f"{x['name']} {x['ms']:.2f}" for x in best), flush=True)
    dev.destroy(); ctx.close()
    if os.environ.get("SHIFT"):
        keep.append(torch.empty(((trial + 1) * 37) << 20, dtype=torch.uint8, device="cuda"))       # moves the next context's allocations
