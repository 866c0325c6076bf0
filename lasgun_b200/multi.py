"""Multi-GPU frame sharding: scene + BVH replicated on every GPU, the film split into
32x32 macro tiles dealt round-robin along anti-diagonals (owner = (mx + my) % ranks, lgb_api.cu
build_tile_list), and one exchange step — the film to rank 0.  Tiles are disjoint, so a SUM reduction of
the zero-initialised per-rank films IS the gather (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
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


def capture_distributed(dev_scene, w: int, h: int, film, rank: int, ranks: int, stream: int = 0):
    """One frame across `ranks` GPUs: render this rank's tiles into `film` (CUDA uint8 tensor), gather to rank 0."""
    if ranks > 1:
        film.zero_()
    dev_scene.capture_device(w, h, film.data_ptr(), rank=rank, ranks=ranks, stream=stream)
    if ranks > 1:
        gather_film(film)
    return film
