"""N > 1 path on the CPU: world_size-2 gloo run of the tile partition + film gather logic
(lasgun_b200/multi.py).  Each rank 'renders' its tiles by masking an oracle frame, then the SUM
gather must reproduce the full frame on rank 0."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, w, h, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from lasgun_b200 import multi, scenes
    from oracle import pyoracle as po
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc, _ = scenes.simple("a", 0, 64)
    full = po.OracleScene(sc).capture(w, h, threads=1)["rgba"]
    owner = multi.tile_owner_map(w, h, world)
    mine = np.where((owner == rank)[..., None], full, 0).astype(np.uint8)
    film = torch.from_numpy(mine.copy())
    multi.gather_film(film)
    if rank == 0:
        q.put(bool(np.array_equal(film.numpy(), full)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_gloo_tile_gather(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 100, 70, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert ok
    assert all(p.exitcode == 0 for p in procs)


def test_tile_owner_map_partitions():
    from lasgun_b200 import multi
    for ranks in (1, 2, 4, 8):
        m = multi.tile_owner_map(3840, 2160, ranks)
        counts = np.bincount(m.reshape(-1), minlength=ranks)
        assert counts.sum() == 3840 * 2160 and (m >= 0).all() and (m < ranks).all()
        assert counts.max() / counts.mean() < 1.03        # interleaved macro tiles balance the pixel count
