"""Why does k_cprimary(+setup) take 10.4 ms inside bench.py and 7.8 ms in scripts/profile_kernels.py?  Same scene, same kernels.
Each ingredient of the bench's set-up is added in turn (DIAG_STEPS, comma separated) and the per-kernel times printed after it."""
import os, sys, time
import torch
sys.path.insert(0, ".")
from lasgun_b200 import _native as N, scenes
steps = os.environ.get("DIAG_STEPS", "plain,tstream,sampler,hostscene").split(",")
sc, (w, h) = scenes.CONFIGS["mixed4k"]()
ctx = N.Context(0)


def show(tag, dev, film):
    best = None
    for _ in range(3):
        k, st = dev.capture_profile(w, h, film.data_ptr())
        if best is None or sum(x["ms"] for x in k) < sum(x["ms"] for x in best):
            best = k
    print(f"{tag:34s}", " ".join(f"{x['name']} {x['ms']:.2f}" for x in best), flush=True)


if "scene_first" in steps:                 # as profile_kernels.py: the scene is created before torch allocates anything
    dev = N.DeviceScene(ctx, N.FlatScene(sc))
    film = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        dev.capture_device(w, h, film.data_ptr(), want_stats=True)
    show("scene first, then the film", dev, film)
    steps = [x for x in steps if x not in ("plain", "first_on_tstream")] + ["skip"]
else:
    film = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
if "skip" in steps:
    pass
elif "first_on_tstream" in steps:            # the bench's first capture (camera grid build, wave buffers) runs on a torch stream
    ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
    hs = N.HostScene(sc); flat = N.FlatScene(hs)
    dev = N.DeviceScene(ctx, flat)
    dev.capture_device(w, h, film.data_ptr(), rank=0, ranks=1, stream=ts.cuda_stream, want_stats=True)
    show("first capture on a torch stream", dev, film)
else:
    dev = N.DeviceScene(ctx, N.FlatScene(sc))
    for _ in range(2):
        dev.capture_device(w, h, film.data_ptr(), want_stats=True)
    show("plain (profile_kernels)", dev, film)
if "tstream" in steps:
    ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
    for _ in range(3):
        dev.capture_device(w, h, film.data_ptr(), rank=0, ranks=1, stream=ts.cuda_stream)
    torch.cuda.synchronize()
    show("after captures on a torch stream", dev, film)
if "sampler" in steps:
    import bench
    s = bench.ClockSampler(0); s.start(); time.sleep(0.5)
    for _ in range(5):
        dev.capture_device(w, h, film.data_ptr())
    torch.cuda.synchronize()
    t0 = s.mark(); time.sleep(0.2); t1 = s.mark(); s.stop(t0, t1)
    show("after the clock sampler", dev, film)
if "events" in steps:
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(5):
        dev.capture_device(w, h, film.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    ev[1].record(); torch.cuda.synchronize()
    print("5 frames under torch events: %.2f ms/frame" % (ev[0].elapsed_time(ev[1]) / 5))
    show("after timed frames", dev, film)
if "nodecount" in steps:
    print("node_count", dev.node_count)
    show("after node_count (BVH built)", dev, film)
