"""Where does the N-GPU frame lose time?  torchrun --nproc-per-node N scripts/diag_multi.py"""
import os, sys, time, statistics
import torch, torch.distributed as dist
sys.path.insert(0, ".")
from lasgun_b200 import _native as N, multi, scenes
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = N.Context(local)
sc, (w, h) = scenes.CONFIGS["mixed4k"]()
host = N.HostScene(sc)
dev = N.DeviceScene(ctx, N.FlatScene(host))
st = torch.cuda.Stream(); torch.cuda.set_stream(st); stream = st.cuda_stream
shared = multi.SharedFilm(ctx, w, h, rank, world, N)
def loop(fn, steps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    host_t = []
    ev[0].record()
    for i in range(steps):
        t0 = time.perf_counter(); fn(); host_t.append((time.perf_counter() - t0) * 1e6); ev[i + 1].record()
    torch.cuda.synchronize(); dist.barrier()
    per = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    return ev[0].elapsed_time(ev[steps]) / steps, statistics.median(per), max(per), statistics.median(host_t)
res = {}
res["barrier(flags)"] = loop(lambda: multi.capture_distributed(dev, w, h, None, rank, world, stream, shared=shared))
res["free-running (no barrier)"] = loop(lambda: dev.capture_device(w, h, shared.ptr, rank=rank, ranks=world, stream=stream))
film = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
res["own film, no barrier"] = loop(lambda: dev.capture_device(w, h, film.data_ptr(), rank=rank, ranks=world, stream=stream))
one = [dev.capture_device(w, h, film.data_ptr(), rank=rank, ranks=world, stream=stream, want_stats=True)["render_ms"] for _ in range(5)]
for k, v in res.items():
    t = torch.tensor(list(v), device="cuda"); g = [torch.zeros_like(t) for _ in range(world)]; dist.all_gather(g, t)
    if rank == 0:
        print(f"{k:28s} ms/frame per rank:", [round(float(x[0]), 3) for x in g], "| median step:", [round(float(x[1]), 3) for x in g], "| host us/step:", [round(float(x[3])) for x in g])
t = torch.tensor([statistics.median(one)], device="cuda"); g = [torch.zeros_like(t) for _ in range(world)]; dist.all_gather(g, t)
if rank == 0: print("render_ms alone (synchronous):", [round(float(x[0]), 3) for x in g])
torch.cuda.synchronize(); dist.barrier()
if rank != 0: shared.close()
dist.barrier()
if rank == 0: shared.close()
dev.destroy(); dist.destroy_process_group()
