import os, sys, time
sys.path.insert(0, ".")
from lasgun_b200 import _native as N, scenes
name = sys.argv[1] if len(sys.argv) > 1 else "mixed4k"
sc, (w, h) = scenes.CONFIGS[name]()
ctx = N.Context(0)
flat = N.FlatScene(sc)
for mode in ("device", "host", "device"):
    if mode == "host": os.environ["LGB_HOST_BUILD"] = "1"
    else: os.environ.pop("LGB_HOST_BUILD", None)
    t = time.perf_counter(); dev = N.DeviceScene(ctx, flat); dt = time.perf_counter() - t
    v = dev.verify()
    st = None
    for _ in range(3):
        rgba, st = dev.capture(w, h)
    print(name, mode, "create ms", round(dt * 1e3, 1), "verify", v, "render_ms", round(st["render_ms"], 2), flush=True)
    if mode == "host": ref = rgba.copy()
    if mode == "device" and "ref" in globals(): print("   film identical to host-built:", bool((rgba == ref).all()))
    dev.destroy()
