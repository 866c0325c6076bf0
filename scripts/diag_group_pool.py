"""Which ingredient of the torchrun bench makes the leader's pool refuse to grow once it is shared with the peers?"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from lasgun_b200 import _native as N, scenes
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
backend = os.environ.get("DIAG_BACKEND", "nccl")
tag = f"world={world} backend={backend if world > 1 else '-'} big={os.environ.get('DIAG_BIG', '1')} ipc={os.environ.get('DIAG_IPC', '0')}"
if world > 1:
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group(backend, device_id=torch.device("cuda", rank) if backend == "nccl" else None)
    x = torch.ones(4, device="cuda"); 
    if backend == "nccl": dist.all_reduce(x)
sc, (w, h) = scenes.CONFIGS["mixed4k"]()
L = N.lib()
if rank == 0:
    hs = N.HostScene(sc)
    try:
        if os.environ.get("DIAG_BIG", "1") == "1":
            ctx = N.Context(0)
            dev = N.DeviceScene(ctx, N.FlatScene(hs))
            film = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda:0")
            dev.capture_device(w, h, film.data_ptr())
            torch.cuda.synchronize()
        g = N.Context(devices=[0, 1])
        gd = N.DeviceScene(g, N.FlatScene(hs, lazy=True))
        gf = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda:0")
        gd.capture_device(w, h, gf.data_ptr())
        t = time.perf_counter(); gd2 = N.DeviceScene(g, N.FlatScene(hs, lazy=True)); dt = time.perf_counter() - t
        print(tag, "OK", f"second create {dt*1e3:.1f} ms", flush=True)
    except Exception as e:
        print(tag, "FAIL", repr(e)[:200], flush=True)
if world > 1:
    dist.barrier() if backend == "gloo" else dist.barrier(device_ids=[rank])
