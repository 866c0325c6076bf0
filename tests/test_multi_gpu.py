"""Two-GPU frame: tiles rendered on two B200s, rank 1's pixels stored straight into rank 0's film over NVLink
(multi.SharedFilm), compared byte for byte with the one-GPU film.  Skipped on boxes with a single GPU."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["LGB_ROOT"])
from lasgun_b200 import _native as N, multi, scenes
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
ctx = N.Context(rank)
sc, (w, h) = scenes.mixed4k(mesh_n=80, nspheres=8000, res=(480, 270), supersampling=1)
host = N.HostScene(sc)
dev = multi.replicate_scene(ctx, lambda: N.FlatScene(host), rank, world, N)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
shared = multi.SharedFilm(ctx, w, h, rank, world, N)
for _ in range(3):                                   # frames back to back: no stale pixels, no missing ones
    out = multi.capture_distributed(dev, w, h, None, rank, world, st.cuda_stream, shared=shared)
torch.cuda.synchronize()
film = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
multi.capture_distributed(dev, w, h, film, rank, world, st.cuda_stream)      # NCCL gather of the same frame
torch.cuda.synchronize()
film2 = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
multi.capture_distributed(dev, w, h, film2, rank, world)                      # no stream given: torch's current stream
torch.cuda.synchronize()
film3 = torch.ones((h, w, 4), dtype=torch.uint8, device="cuda")
with torch.cuda.stream(torch.cuda.default_stream()):                          # the legacy default stream: host-fenced, still whole
    multi.capture_distributed(dev, w, h, film3, rank, world)
torch.cuda.synchronize()
if rank == 0:
    assert torch.equal(film2, film) and torch.equal(film3, film), "capture_distributed without an explicit stream lost pixels"
    one = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    dev.capture_device(w, h, one.data_ptr(), rank=0, ranks=1, stream=st.cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(out, one), "peer-stored film differs from the one-GPU film"
    assert torch.equal(film, one), "NCCL-gathered film differs from the one-GPU film"
    assert int(one[..., 3].min()) == 255
    print("MULTI_GPU_OK")
dist.barrier()
if rank != 0: shared.close()
dist.barrier()
if rank == 0: shared.close()
dev.destroy()
dist.destroy_process_group()
"""


@pytest.mark.gpu
def test_two_gpu_shared_film(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, LGB_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29631", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI_GPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.gpu
def test_device_group_in_one_process():
    """lgb_init_devices: one process, one blocking capture over every GPU of the box (the reference's own shape, lib.rs:55-104);
    the film must equal the one-GPU film byte for byte, also for a lazy scene and through the host-pointer entry."""
    import numpy as np
    import torch
    from lasgun_b200 import _native as N, scenes
    n = torch.cuda.device_count()
    single = N.Context(0)
    group = N.Context(devices=list(range(n)))            # n == 1: a group of one must behave like lgb_init
    assert group.n_devices == n and single.n_devices == 1
    try:
        for mk, lazy in ((lambda: scenes.mixed4k(mesh_n=80, nspheres=8000, res=(480, 270), supersampling=1), False),
                         (lambda: scenes.mixed4k(mesh_n=200, nspheres=40000, res=(640, 360), supersampling=3), True),
                         (lambda: scenes.simple("b", 2, 200), False)):
            sc, (w, h) = mk()
            host = N.HostScene(sc)
            one = N.DeviceScene(single, N.FlatScene(host))
            ref, st1 = one.capture(w, h)
            one.destroy()
            dev = N.DeviceScene(group, N.FlatScene(host, lazy=lazy))
            for _ in range(2):
                out, stn = dev.capture(w, h)
                assert np.array_equal(out, ref)
                assert stn["primary_rays"] == st1["primary_rays"] and stn["primary_hits"] == st1["primary_hits"]
            film = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda:0")
            dev.capture_device(w, h, film.data_ptr(), want_stats=True)
            assert np.array_equal(film.cpu().numpy(), ref)
            dev.destroy()
    finally:
        group.close(); single.close()
