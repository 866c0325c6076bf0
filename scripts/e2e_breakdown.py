"""Where does capture(scene, film) with host buffers spend its time?  (run with LGB_TIMING=1 for the phases)"""
import ctypes as C, sys, time
import numpy as np
sys.path.insert(0, ".")
from lasgun_b200 import _native as N, scenes
name = sys.argv[1] if len(sys.argv) > 1 else "mixed4k"
sc, (w, h) = scenes.CONFIGS[name]()
ctx = N.Context(0); L = N.lib()
t = time.perf_counter(); hs = N.HostScene(sc); t_replay = time.perf_counter() - t
film = np.zeros((h, w, 4), np.uint8)
for it in range(4):
    t = time.perf_counter(); flat = N.FlatScene(hs, lazy=True); t_flat = time.perf_counter() - t
    hsc = C.c_void_p()
    flat.desc.expected_film_pixels = w * h          # capture(scene, film) knows its film
    t0 = time.perf_counter(); ctx.check(L.lgb_scene_create(ctx.h, C.byref(flat.desc), C.byref(hsc))); t1 = time.perf_counter()
    st = N.Stats()
    ctx.check(L.lgb_capture(ctx.h, hsc, w, h, film.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(st))); t2 = time.perf_counter()
    L.lgb_scene_destroy(hsc); t3 = time.perf_counter()
    print(f"{name} it{it}: replay {t_replay*1e3:.0f} | flatten {t_flat*1e3:.1f} scene_create {1e3*(t1-t0):.1f} capture {1e3*(t2-t1):.1f} (render {st.render_ms:.1f}, total-dev {st.total_ms:.1f}) destroy {1e3*(t3-t2):.1f} | e2e {1e3*(t3-t0)+t_flat*1e3:.1f} ms", flush=True)
