"""Multi-GPU frame sharding: scene + BVH replicated on every GPU, the film split into
32x32 macro tiles dealt round-robin along anti-diagonals (owner = (mx + my) % ranks, lgb_api.cu
build_tile_list), and one exchange step — the film to rank 0.

On GPUs the exchange is fused into the render: rank 0 owns the film (SharedFilm), the other ranks map it through a CUDA IPC
handle and their resolve stage stores its pixels straight into rank 0's HBM over NVLink; a barrier ends the frame.  The
NCCL form (a SUM reduction of zero-initialised per-rank films: tiles are disjoint, so SUM is the gather) remains for the
gloo CPU tests and as a fallback."""
from __future__ import annotations

import numpy as np

MACRO_TILE = 32


def tile_owner_map(w: int, h: int, ranks: int) -> np.ndarray:
    """(h, w) array of the rank that renders each pixel (mirrors build_tile_list in csrc/lgb_api.cu)."""
    my, mx = np.meshgrid(np.arange(h) // MACRO_TILE, np.arange(w) // MACRO_TILE, indexing="ij")
    return ((mx + my) % ranks).astype(np.int32)


def gather_film(film, dst: int = 0, group=None):
    """film: torch uint8 tensor (h, w, 4), zero outside this rank's tiles.  After the call rank `dst` holds the frame."""
    import torch.distributed as dist
    dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return film


class _DevicePointer:
    """Zero-copy torch view of library-owned device memory (`__cuda_array_interface__`)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def replicate_scene(ctx, flat_builder, rank: int, ranks: int, native):
    """Scene + BVH on every GPU, built ONCE: rank 0 runs `flat_builder()` (the reference BVH build + flatten) and
    lgb_scene_create, then broadcasts the relocatable device arena over NCCL / NVLink; the other ranks import it.
    Returns the rank's DeviceScene."""
    import torch
    import torch.distributed as dist
    if ranks == 1:
        return native.DeviceScene(ctx, flat_builder())
    nl = int(native.lib().lgb_scene_layout_bytes())
    if rank == 0:
        dev = native.DeviceScene(ctx, flat_builder())
        layout, ptr, nbytes = dev.export()
        head = torch.frombuffer(bytearray(layout) + int(dev.spp).to_bytes(8, "little"), dtype=torch.uint8).cuda()
        dist.broadcast(head, src=0)
        dist.broadcast(torch.as_tensor(_DevicePointer(ptr, nbytes), device="cuda"), src=0)
        return dev
    head = torch.empty(nl + 8, dtype=torch.uint8, device="cuda")
    dist.broadcast(head, src=0)
    raw = bytes(head.cpu().numpy())
    layout, spp = raw[:nl], int.from_bytes(raw[nl:], "little")
    nbytes = int.from_bytes(layout[8:16], "little")          # SceneLayout.arena_bytes
    arena = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dist.broadcast(arena, src=0)
    torch.cuda.current_stream().synchronize()
    return native.DeviceScene.adopt(ctx, layout, arena.data_ptr(), spp, keep=arena)


class SharedFilm:
    """The frame's film in rank 0's HBM, mapped by every other rank (lgb_film_alloc_shared / lgb_film_open_shared).
    Collective: every rank constructs it.  `tensor` (rank 0 only) is a zero-copy (h, w, 4) uint8 view."""

    def __init__(self, ctx, w: int, h: int, rank: int, ranks: int, native):
        import ctypes as C

        import torch
        import torch.distributed as dist
        self.ctx, self.rank, self.native, self.w, self.h = ctx, rank, native, w, h
        L = native.lib()
        handle = (C.c_uint8 * 64)()
        ptr = C.c_void_p()
        self.ranks = ranks
        self.film_bytes = w * h * 4
        self.flag_off = (self.film_bytes + 255) & ~255          # flags behind the film: word 0 = rank 0's `go`, word 16 r = rank r's `done`
        self.frame = 0
        if rank == 0:
            ctx.check(L.lgb_film_alloc_shared(ctx.h, self.flag_off + 64 * (ranks + 1), C.byref(ptr), handle))
        t = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device="cuda")
        if ranks > 1:
            dist.broadcast(t, src=0)
        if rank != 0:
            hb = (C.c_uint8 * 64)(*t.cpu().tolist())
            ctx.check(L.lgb_film_open_shared(ctx.h, hb, C.byref(ptr)))
        self.ptr = ptr.value
        self.tensor = torch.as_tensor(_DevicePointer(self.ptr, w * h * 4), device="cuda").view(h, w, 4) if rank == 0 else None
        self._sync = torch.zeros(1, dtype=torch.int32, device="cuda")
        if rank == 0:
            torch.as_tensor(_DevicePointer(self.ptr + self.flag_off, 64 * (ranks + 1)), device="cuda").zero_()
            torch.cuda.synchronize()
        if ranks > 1:
            dist.barrier()                                   # the flags are zero before any rank looks at them

    def close(self):
        if self.ptr:
            self.ctx.check(self.native.lib().lgb_film_release_shared(self.ctx.h, self.native.C.c_void_p(self.ptr), 1 if self.rank == 0 else 0))
            self.ptr = None


def capture_distributed(dev_scene, w: int, h: int, film, rank: int, ranks: int, stream: int | None = None, shared: SharedFilm | None = None):
    """One frame across `ranks` GPUs.  With `shared`: every rank's kernels store their tiles into rank 0's film over NVLink
    and a stream-ordered barrier completes the frame (no collective moves pixels).  Otherwise: render this rank's tiles into
    `film` (CUDA uint8 tensor) and SUM-reduce to rank 0.

    `stream` is the cudaStream_t the kernels are queued on; it must be the stream torch's collectives and `film.zero_()` run on
    (default: torch's current stream), otherwise the zeroing, the reduce or the end-of-frame barrier are not ordered against
    the stores.  With `shared`, the barrier at the end of frame k also orders rank 0's readers of frame k before any rank's
    stores of frame k + 1 only if the reader runs on this same stream before the next call (the bench and the tests do)."""
    import torch
    cur = torch.cuda.current_stream().cuda_stream
    if stream is None:
        stream = cur
    # the legacy default stream (0) cannot be handed to the library (NULL means its private stream), and a stream other than torch's
    # current one is not ordered against the collectives: in both cases fence on the host instead (slower, never wrong)
    fenced = ranks > 1 and (stream == 0 or stream != cur)

    def render(ptr):
        if fenced:
            torch.cuda.current_stream().synchronize()           # film.zero_() / the previous frame's readers are done
            dev_scene.capture_device(w, h, ptr, rank=rank, ranks=ranks, stream=stream, want_stats=True)      # synchronises
        else:
            dev_scene.capture_device(w, h, ptr, rank=rank, ranks=ranks, stream=stream)
    if shared is not None:
        # The frame barrier is two flag words in rank 0's memory, not a collective: rank 0 opens frame f (`go` = f: it is done with
        # the film of frame f - 1), every other rank waits for that word, renders, and stores f into its own `done` word behind a
        # system fence; rank 0 renders and then waits for all `done` words.  All of it queued on `stream`: the frame is whole when
        # rank 0's stream reaches the end of this call.
        L, ctx = shared.native.lib(), shared.ctx
        vp = shared.native.C.c_void_p
        shared.frame += 1
        f, flags = shared.frame, shared.ptr + shared.flag_off
        st = vp(stream) if stream else None
        if ranks > 1 and not fenced:
            if rank == 0:
                ctx.check(L.lgb_film_signal(ctx.h, vp(flags), f, st))
                render(shared.ptr)
                ctx.check(L.lgb_film_wait(ctx.h, vp(flags), 1, ranks - 1, 64, f, st))
            else:
                ctx.check(L.lgb_film_wait(ctx.h, vp(flags), 0, 1, 64, f, st))
                render(shared.ptr)
                ctx.check(L.lgb_film_signal(ctx.h, vp(flags + 64 * rank), f, st))
        else:
            import torch.distributed as dist
            render(shared.ptr)
            if ranks > 1:
                dist.all_reduce(shared._sync)        # host-fenced variant: a collective barrier on the current stream
        return shared.tensor
    if ranks > 1:
        film.zero_()
    render(film.data_ptr())
    if ranks > 1:
        gather_film(film)
    return film
