"""`capture(scene, film)` and friends (src/lib.rs:42-162) on the B200 path.

capture = replay the description into the C++ host (BVH build + flatten, Accel::from) ->
lgb_scene_create (H2D) -> lgb_capture (kernels + D2H).  No CPU fallback exists.
"""
from __future__ import annotations

from . import _native as N
from .api import Film, Scene


class Accel:
    """`Accel::from(&scene)` (lib.rs:42, bvh.rs:135): the scene's BVH, flattened and resident on the GPU."""

    def __init__(self, scene: Scene, ctx: N.Context | None = None, lazy: bool = False):
        self.scene = scene
        self.ctx = ctx or N.default_context()
        self.flat = N.FlatScene(scene, lazy=lazy)
        self.dev = N.DeviceScene(self.ctx, self.flat)

    @staticmethod
    def from_scene(scene: Scene, **kw):
        return Accel(scene, **kw)

    def close(self):
        self.dev.destroy()


def capture(scene: Scene, film: Film, ctx: N.Context | None = None):
    """lib.rs:55.  Blocking; fills film.pixels() (row-major RGBA8).  Returns the device stats."""
    root = Accel(scene, ctx, lazy=True)          # the reference BVH is built only if a ray meets an exact-t tie
    try:
        _, stats = root.dev.capture(film.w, film.h, out=film.output)
    finally:
        root.close()
    return stats


def capture_subset(k: int, n: int, root: Accel, film: Film):
    """lib.rs:110: pixels k, k+n, ... of the row-major film only."""
    return root.dev.capture_subset(k, n, film.w, film.h, film.output)


def render(scene: Scene, resolution):
    """lib.rs:46."""
    film = Film(resolution[0], resolution[1])
    capture(scene, film)
    return film
