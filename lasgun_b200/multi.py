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


def capture_distributed(dev_scene, w: int, h: int, film, rank: int, ranks: int, stream: int = 0):
    """One frame across `ranks` GPUs: render this rank's tiles into `film` (CUDA uint8 tensor), gather to rank 0."""
    if ranks > 1:
        film.zero_()
    dev_scene.capture_device(w, h, film.data_ptr(), rank=rank, ranks=ranks, stream=stream)
    if ranks > 1:
        gather_film(film)
    return film
