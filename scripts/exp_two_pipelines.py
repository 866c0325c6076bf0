"""Experiment: one frame as two half-frames (tile ranks 0/2 and 1/2) in flight at once on two contexts of the same GPU."""
import sys, time
import torch
sys.path.insert(0, ".")
from lasgun_b200 import _native as N, scenes
name = sys.argv[1] if len(sys.argv) > 1 else "mixed4k"
parts = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sc, (w, h) = scenes.CONFIGS[name]()
ctxs = [N.Context(0) for _ in range(parts)]
dev = N.DeviceScene(ctxs[0], N.FlatScene(sc))
layout, ptr, nbytes = dev.export()
devs = [dev] + [N.DeviceScene.adopt(c, layout, ptr, dev.spp, keep=dev) for c in ctxs[1:]]      # the same arena, one handle per context
film = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
ref = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
L = N.lib()
import ctypes as C
streams = [torch.cuda.Stream() for _ in range(parts)]
def frame():
    for r, (c, s, d) in enumerate(zip(ctxs, streams, devs)):
        c.check(L.lgb_capture_device(c.h, d.h, w, h, r, parts, C.c_void_p(film.data_ptr()), C.c_void_p(s.cuda_stream), None))
for _ in range(3):
    frame(); torch.cuda.synchronize()
t = []
for _ in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter(); frame(); torch.cuda.synchronize(); t.append((time.perf_counter() - t0) * 1e3)
print(name, parts, "pipelines in flight: ms/frame", [round(x, 2) for x in t])
dev.capture_device(w, h, ref.data_ptr(), want_stats=True)
t1 = []
for _ in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter(); dev.capture_device(w, h, ref.data_ptr(), stream=streams[0].cuda_stream); torch.cuda.synchronize(); t1.append((time.perf_counter() - t0) * 1e3)
print("one pipeline: ms/frame", [round(x, 2) for x in t1], "films equal", bool(torch.equal(film, ref)))
