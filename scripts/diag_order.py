"""k_cprimary takes 6.75 ms when the scene is created before torch allocates the film and 8.8 ms the other way round (same box, same
kernels: profiles/r2_v32_cprimary_regs.txt).  Which ingredient of "torch first" does it?  MODE selects what happens before the scene."""
import ctypes as C, os, sys
mode = os.environ.get("MODE", "scene_first")
import torch
sys.path.insert(0, ".")
from lasgun_b200 import _native as N, scenes
sc, (w, h) = scenes.CONFIGS["mixed4k"]()
hs = N.HostScene(sc)
film = None
rt = C.CDLL("libcudart.so.12")
def raw_malloc(nbytes):
    p = C.c_void_p(); assert rt.cudaMalloc(C.byref(p), C.c_size_t(nbytes)) == 0; return p.value
if mode == "torch_init_only":
    torch.cuda.init(); torch.cuda.synchronize()
elif mode == "film_first":
    film = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda").data_ptr()
elif mode == "raw_film_first":
    film = raw_malloc(w * h * 4)
elif mode == "small_torch_first":
    keep = torch.zeros(16, device="cuda")
elif mode == "big_raw_first":                   # something of the wave buffers' size allocated and freed first
    p = raw_malloc(8 << 30); rt.cudaFree(C.c_void_p(p))
elif mode == "raw2mb_before_init":
    raw_malloc(2 << 20)
elif mode == "small_torch_then_big_kept":
    keep = torch.zeros(16, device="cuda"); raw_malloc(8 << 30)
ctx = N.Context(0)
if mode == "raw2mb_after_init":
    raw_malloc(2 << 20)
if mode == "two_scenes":
    dev0 = N.DeviceScene(ctx, N.FlatScene(hs))
dev = N.DeviceScene(ctx, N.FlatScene(hs))
if mode == "two_scenes":
    dev0.destroy()
if mode == "raw2mb_after_scene":
    raw_malloc(2 << 20)
if mode == "raw64kb_after_scene":
    raw_malloc(64 << 10)
if film is None:
    if mode == "scene_first_raw_film": film = raw_malloc(w * h * 4)
    else:
        t = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda"); film = t.data_ptr()
for _ in range(2):
    dev.capture_device(w, h, film, want_stats=True)
best = None
for _ in range(3):
    k, st = dev.capture_profile(w, h, film)
    if best is None or sum(x["ms"] for x in k) < sum(x["ms"] for x in best):
        best = k
print(f"{mode:22s}", " ".join(f"{x['name']} {x['ms']:.2f}" for x in best), flush=True)
