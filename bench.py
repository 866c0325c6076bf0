#!/usr/bin/env python
"""bench.py — Mrays/s (primary + shadow) and ms/frame of the lasgun render loop on 1/2/4/8 B200.

  python bench.py --gpus N --steps K --warmup W [--workload mixed4k] [--impl reference]

A step is one full frame of the workload (default: BASELINE.json's config 5, `mixed4k`, 3840x2160 at
16 spp — the config the 1/2/4/8-GPU metric is quoted on; it fits one GPU).  Rays are counted with the
reference's semantics: primary = w*h*spp, shadow = lights * primary hits (integrate.rs:47-50).
  value  scene + BVH resident in HBM, frame rendered into a device film (rank 0 after the NVLink gather)
  e2e    `capture(scene, film)` as the reference defines it (lib.rs:55-104), with HOST buffers, every step:
         flatten of the host scene (the reference's own HLBVH build is deferred: the device asks for it only when
         a ray meets two primitives at bit-identical t, which this workload never does), lgb_scene_create (device
         BVH build, H2D of the scene), lgb_capture (render + D2H of the film), lgb_scene_destroy
--impl reference times the CPU oracle (C++ restatement of the reference algorithm — the Rust build
cannot be compiled here) on all host threads over a bounded sample of the same frame.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from lasgun_b200 import scenes  # noqa: E402

METRIC = "Mrays/s (primary+shadow)"
# Constants of record, SURVEY.md §8(d): FP32 operations (FMA = 2) and bytes fetched per unit of work.
OPS = {"node": 26, "sphere": (26, 39), "tri": (47, 68), "cuboid": 31, "camera": 30, "hit": 110, "light": 25 + 120, "ambient": 120,
       "background": 15, "film": 12}
BYTES = {"node": 32, "sphere": 16, "tri": 48, "cuboid": 32, "ref": 4, "film": 4}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "sm_max_mhz": float(p.get("sm_max_mhz", 1965.0)), "source": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (nvidia_ml_py)."""
    HW = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.samples, self.reasons, self.stop_flag, self.max_mhz, self.thread = [], set(), False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.HW.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if self.nv:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join()
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def workload(name, args):
    if name == "mixed4k" and args.small:
        return scenes.mixed4k(mesh_n=80, nspheres=8000, res=(480, 270), supersampling=1)
    return scenes.CONFIGS[name]()


def reference_sample(osc, w, h, threads, target_s, spp, n_lights):
    """Bounded sample of the frame on the CPU oracle: capture_subset(k, n) for k < threads."""
    probe_n = max(threads, (w * h) // (threads * 64))
    t0 = time.time(); r = osc.capture(w, h, threads=threads, counters=True, subset=(probe_n, 0, threads)); dt = max(time.time() - t0, 1e-4)
    pixels = threads * ((w * h + probe_n - 1) // probe_n)
    rate = pixels / dt
    want_pixels = min(w * h, max(pixels, int(rate * target_s)))
    n = max(threads, (w * h * threads) // want_pixels)
    return n


def run_reference(args):
    from oracle import pyoracle as po
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sc, (w, h) = workload(args.workload, args)
    threads = po.hardware_threads()
    osc = po.OracleScene(sc)
    spp, nl = sc.camera.num_samples(), len(sc.lights)
    n = reference_sample(osc, w, h, threads, args.ref_seconds / max(1, args.steps + args.warmup), spp, nl)
    times, rays = [], 0
    for i in range(args.warmup + args.steps):
        r = osc.capture(w, h, threads=threads, counters=True, subset=(n, 0, threads))
        if i >= args.warmup:
            times.append(r["render_ms"]); rays = r["counters"]["primary"] + r["counters"]["shadow"]
    ms = sum(times) / len(times)
    value = rays / (ms * 1e-3) / 1e6
    frame_rays_est = rays * (n / threads)
    line = {
        "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "impl": "reference",
        "config": {"workload": args.workload, "film": [w, h], "spp": spp, "lights": nl},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": "port",
                         "sample": f"capture_subset(k, n={n}) for k in 0..{threads} of the {w}x{h} frame ({rays} rays/step); "
                                   "C++ restatement of the reference algorithm, not the Rust build",
                         "bvh_build_ms": osc.build_ms, "ms_per_frame_est": ms * n / threads},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "est_frame_rays": frame_rays_est,
    }
    print(json.dumps(line))


def traversal_work(node_fetches, f, e):
    """Algorithmic FP32 ops and bytes of BVH traversal: one node fetch = two slab tests of 26 ops / 32 B each;
    f / e = primitive tests started / carried to the hit path, per type (sphere, cuboid, triangle)."""
    ops = 2 * node_fetches * OPS["node"]
    ops += (f[0] - e[0]) * OPS["sphere"][0] + e[0] * OPS["sphere"][1]
    ops += f[1] * OPS["cuboid"]
    ops += (f[2] - e[2]) * OPS["tri"][0] + e[2] * OPS["tri"][1]
    byts = 2 * node_fetches * BYTES["node"] + f[0] * BYTES["sphere"] + f[1] * BYTES["cuboid"] + f[2] * BYTES["tri"]
    return ops, byts


def algorithmic_work(st, w, h, spp, n_lights):
    """Algorithmic FP32 ops and bytes of one frame from the device's own work counters."""
    ops, byts = traversal_work(st["node_tests"], st["filter_tests"], st["exact_tests"])
    hits, prim = st["primary_hits"], st["primary_rays"]
    ops += prim * OPS["camera"] + hits * OPS["hit"] + st["shadow_rays_traced"] * OPS["light"] + hits * OPS["ambient"]
    ops += (prim - hits) * OPS["background"] + w * h * OPS["film"]
    byts += w * h * BYTES["film"]
    return ops, byts


def primary_kernel_work(st):
    """The dominant kernel (k_primary: camera ray + closest-hit traversal of every sample)."""
    ops, byts = traversal_work(st["primary_node_tests"], st["primary_filter_tests"], st["primary_exact_tests"])
    return ops + st["primary_rays"] * OPS["camera"], byts


def ncu_traffic(workload, world):
    """dram read + write bytes of the primary-ray phase of one frame (k_beam + k_leafp + k_primary fallback) from the committed
    `ncu --set full` capture (or None)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        e = t.get(workload)
        return (e["primary_phase_dram_bytes"] / world) if e and world == 1 else None
    except Exception:
        return None


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from lasgun_b200 import _native as N
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = N.Context(local)
    if os.environ.get("LGB_SIDE") == "0":              # experiments: the shadow chains of the lights on one stream
        ctx.set_side_streams(False)
    sc, (w, h) = workload(args.workload, args)
    spp, nl = sc.camera.num_samples(), len(sc.lights)
    hscene_host = N.HostScene(sc)                      # the built `Scene` (scene construction is not part of capture)
    flat = N.FlatScene(hscene_host)
    dev = N.DeviceScene(ctx, flat)
    film = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    # everything (our kernels, NCCL, the timing events) runs on ONE non-default torch stream so that
    # torch.cuda.Event brackets exactly the launches it is supposed to
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    from lasgun_b200 import multi

    # N > 1: rank 0 owns the film, the other ranks' kernels store their tiles into it over NVLink (multi.SharedFilm);
    # --gather nccl keeps per-rank films and SUM-reduces them instead
    shared = multi.SharedFilm(ctx, w, h, rank, world, N) if (world > 1 and args.gather == "p2p") else None

    def step():
        multi.capture_distributed(dev, w, h, film, rank, world, stream, shared=shared)

    # one counted frame: ray counts + work counters (not timed).  The algorithmic work of SURVEY 8d is the PER-RAY walk of the
    # device BVH, so it is counted with the pixel beams off; the timed frames (beams automatic) share the interior-node tests
    # of a pixel's rays and execute fewer (roofline.executed_ops_per_launch).
    ctx.set_count_work(True)
    ctx.set_beams(0)
    st = dev.capture_device(w, h, film.data_ptr(), rank=rank, ranks=world, stream=stream, want_stats=True)
    ctx.set_beams(-1)
    st_exec = dev.capture_device(w, h, film.data_ptr(), rank=rank, ranks=world, stream=stream, want_stats=True)
    ctx.set_count_work(False)
    cnt_e = torch.tensor([st_exec["primary_rays"], st_exec["primary_node_tests"]] + st_exec["primary_filter_tests"] + st_exec["primary_exact_tests"], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(cnt_e)
    te = cnt_e.tolist()
    frame_exec = {"primary_rays": te[0], "primary_node_tests": te[1], "primary_filter_tests": te[2:5], "primary_exact_tests": te[5:8]}
    cnt = torch.tensor([st["primary_rays"], st["primary_hits"], st["shadow_rays_traced"], st["node_tests"]] + st["filter_tests"] + st["exact_tests"]
                       + [st["primary_node_tests"]] + st["primary_filter_tests"] + st["primary_exact_tests"] + [st["shadow_cache_hits"]],
                       dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(cnt)
    tot = cnt.tolist()
    frame = {"primary_rays": tot[0], "primary_hits": tot[1], "shadow_rays_traced": tot[2], "node_tests": tot[3],
             "filter_tests": tot[4:7], "exact_tests": tot[7:10], "primary_node_tests": tot[10], "primary_filter_tests": tot[11:14],
             "primary_exact_tests": tot[14:17], "shadow_cache_hits": tot[17]}
    rays_frame = frame["primary_rays"] + nl * frame["primary_hits"]

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local); sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kern_ms = []
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(args.steps):
        step()
    ev[1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    total_ms = torch.tensor([ev[0].elapsed_time(ev[1])], device="cuda")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(total_ms.item()) / args.steps
    value = rays_frame / (ms_per_step * 1e-3) / 1e6

    gather_same = None
    if shared is not None:                       # not timed: the peer-stored film against the NCCL-gathered one, every byte
        multi.capture_distributed(dev, w, h, film, rank, world, stream)
        torch.cuda.synchronize()
        if rank == 0:
            gather_same = bool(torch.equal(film, shared.tensor))

    # render-kernel time alone (CUDA events inside the library, on its launch stream), a few frames
    phase_ms = []
    for _ in range(max(3, args.steps)):
        s2 = dev.capture_device(w, h, film.data_ptr(), rank=rank, ranks=world, stream=stream, want_stats=True)
        kern_ms.append(s2["render_ms"]); phase_ms.append(s2["kernel_ms"])
    kms = torch.tensor([statistics.median(kern_ms)] + [sum(p[i] for p in phase_ms) / len(phase_ms) for i in range(6)], device="cuda")
    if world > 1:
        dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    kernel_ms = float(kms[0].item())
    phases = [float(v) for v in kms[1:].tolist()]      # average launch duration per phase, CUDA events on the launch stream

    # e2e through the C ABI with host buffers (rank-local frame share; film gathered on the host side of rank 0)
    host_film_t = torch.zeros((h, w, 4), dtype=torch.uint8).pin_memory()       # the caller's film: pinned host memory
    host_film = host_film_t.numpy()
    e2e_ms, e2e_parts = [], []
    L = N.lib()
    for i in range(args.e2e_steps + 1):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        if world == 1:
            flat_i = N.FlatScene(hscene_host, lazy=True)   # Accel::from minus the reference BVH build: that runs (callback) only if a ray meets an exact-t tie
            t1 = time.perf_counter()
            hscene = C.c_void_p()
            ctx.check(L.lgb_scene_create(ctx.h, C.byref(flat_i.desc), C.byref(hscene)))
            t2 = time.perf_counter()
            ctx.check(L.lgb_capture(ctx.h, hscene, w, h, host_film.ctypes.data_as(C.POINTER(C.c_uint8)), None))
            L.lgb_scene_destroy(hscene)
        else:
            # the BVHs are built once, on rank 0, and the device arena is broadcast over NVLink (multi.replicate_scene)
            tf = [t0]
            def build_flat():
                f = N.FlatScene(hscene_host); tf[0] = time.perf_counter(); return f
            dev_i = multi.replicate_scene(ctx, build_flat, rank, world, N)
            t1 = tf[0]; t2 = time.perf_counter()
            out = multi.capture_distributed(dev_i, w, h, film, rank, world, stream, shared=shared)
            if rank == 0:
                host_film_t.copy_(out, non_blocking=True)
            torch.cuda.synchronize()
            dev_i.destroy()
        t3 = time.perf_counter()
        e2e_parts.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
        dt = torch.tensor([(t3 - t0) * 1e3], device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if i > 0:
            e2e_ms.append(float(dt.item()))
    e2e = statistics.median(e2e_ms)
    scene_bytes = int(dev.device_bytes)
    d_ = flat.desc                                   # what lgb_scene_create copies to the device: the caller's arrays + rank tables
    h2d_bytes = int(d_.n_spheres * 40 + d_.n_cuboids * 56 + d_.n_triangles * (44 + (36 if d_.tri_normals else 0))
                    + 8 * 4 * (d_.n_spheres + d_.n_cuboids + d_.n_triangles + 1))

    if rank == 0:
        peaks = load_peaks()
        ceil = ctx.measure()
        ops, byts = algorithmic_work(frame, w, h, spp, nl)
        fp32_peak = 2.0 * ceil["fp32_ffma_glanes"]                 # Gop/s with FMA = 2, measured live on this GPU
        ach = ops / (kernel_ms * 1e-3) / 1e9 / world                # per GPU, whole frame
        p_ops, p_bytes = primary_kernel_work(frame)
        e_ops, e_bytes = primary_kernel_work(frame_exec)
        sh_ops, sh_bytes = traversal_work(frame["node_tests"] - frame["primary_node_tests"],
                                          [a - b for a, b in zip(frame["filter_tests"], frame["primary_filter_tests"])],
                                          [a - b for a, b in zip(frame["exact_tests"], frame["primary_exact_tests"])])
        sh_ms = phases[2] + phases[3]
        l1_peak = 128.0 * torch.cuda.get_device_properties(local).multi_processor_count * float(clocks["sm_mhz"] or peaks["sm_max_mhz"]) * 1e6 / 1e9
        p_ach = p_ops / (phases[0] * 1e-3) / 1e9 / world           # dominant kernel alone
        l2_ach = byts / (kernel_ms * 1e-3) / 1e9 / world
        hbm_bytes = scene_bytes + frame["primary_rays"] * 48 / world + w * h * 4
        line = {
            "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": args.workload, "film": [w, h], "spp": spp, "lights": nl, "triangles": int(flat.desc.n_triangles),
                       "spheres": int(flat.desc.n_spheres), "cuboids": int(flat.desc.n_cuboids), "bvh_nodes": int(flat.desc.n_nodes),
                       "device_bvh_nodes": int(dev.node_count) if hasattr(dev, "node_count") else None, "parallelism": f"tiles{world}",
                       "film_gather": None if world == 1 else ("peer stores into rank 0's film (CUDA IPC over NVLink), identical to the NCCL-reduced film: %s" % gather_same
                                                               if shared is not None else "NCCL reduce(SUM) of disjoint tiles"),
                       "l2_policy": "per-frame working set (radiance buffer %.0f MB + scene %.0f MB) exceeds the 126 MB L2" % (frame["primary_rays"] * 24 / 1e6 / world, scene_bytes / 1e6)},
            "ms_per_frame": ms_per_step, "kernel_ms_per_frame": kernel_ms, "rays_per_frame": rays_frame,
            "rays_traced_per_frame": frame["primary_rays"] + frame["shadow_rays_traced"],
            "e2e": {"value": rays_frame / (e2e * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_frame": e2e,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": w * h * 4,
                    "reference_tree_built": bool(world > 1 or flat_i.tree_built),
                    "parts_ms": dict(zip(("flatten", "scene_create_device_bvh_upload", "render_readback_destroy"),
                                         [statistics.median(p[i] for p in e2e_parts[1:]) for i in range(3)]))},
            "gpu_launches": int(s2["kernel_launches"]) * args.steps,
            "roofline": {"kernel": ("primary-ray phase (k_beam bundle traversal + k_leafp leaf walk + k_primary fallback)" if s2["beams"] else
                                    "k_primary (camera rays + closest-hit traversal)") + ", %.1f%% of the frame" % (100.0 * phases[0] / max(sum(phases), 1e-9)),
                         "bound": "fp32_issue", "achieved": p_ach, "peak": fp32_peak, "unit": "Gop/s (FMA=2)", "frac": p_ach / fp32_peak,
                         "traffic": ncu_traffic(args.workload, world), "peak_source": "lgb_measure_fp32_gops, live on this GPU",
                         "algorithmic_ops_per_launch": p_ops, "algorithmic_bytes_per_launch": p_bytes, "launch_ms": phases[0],
                         "algorithmic_definition": "per-ray walk of the device BVH (SURVEY 8d), counted on a frame with LGB_OPT_BEAMS=0",
                         "executed_ops_per_launch": e_ops, "executed_bytes_per_launch": e_bytes,
                         "shadow_phase": {"kernels": "k_shadow (anchor rays) + k_pretest + k_shadow (rest)", "launch_ms": sh_ms,
                                          "algorithmic_ops": sh_ops, "achieved": sh_ops / (sh_ms * 1e-3) / 1e9 / world,
                                          "frac": sh_ops / (sh_ms * 1e-3) / 1e9 / world / fp32_peak},
                         # node / primitive fetches are L1 hits (96 %): the ceiling that binds is the L1 data pipe, 128 B/clk/SM
                         "l1_fetch": {"achieved_gbs": p_bytes / (phases[0] * 1e-3) / 1e9 / world, "peak_gbs": l1_peak,
                                      "frac": p_bytes / (phases[0] * 1e-3) / 1e9 / world / l1_peak,
                                      "peak_source": "128 B/clk/SM x SM count x SM clock under load"},
                         "frame": {"achieved": ach, "frac": ach / fp32_peak, "algorithmic_ops_per_frame": ops, "algorithmic_bytes_per_frame": byts},
                         "phase_ms": dict(zip(("primary", "setup", "shadow_anchor", "pretest_shadow_rest", "shade", "resolve"), phases)),
                         "l2": {"achieved_gbs": l2_ach, "peak_gbs": ceil["l2_read_gbs"], "frac": l2_ach / ceil["l2_read_gbs"]},
                         "hbm": {"achieved_gbs": hbm_bytes / (kernel_ms * 1e-3) / 1e9, "peak_gbs": peaks["hbm_gbs"],
                                 "frac": hbm_bytes / (kernel_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "peak_source": peaks["source"]},
                         "fp64_dfma_glanes_peak": ceil["fp64_dfma_glanes"]},
            "work_per_frame": frame,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import pyoracle as po
            threads = po.hardware_threads()
            osc = po.OracleScene(sc)
            n = reference_sample(osc, w, h, threads, args.cpu_seconds, spp, nl)
            r = osc.capture(w, h, threads=threads, counters=True, subset=(n, 0, threads))
            rays = r["counters"]["primary"] + r["counters"]["shadow"]
            # spot-check the device film at the oracle's sampled pixels (the checker, not the product)
            idx = np.concatenate([np.arange(k, w * h, n) for k in range(threads)])
            same = (host_film.reshape(-1, 4)[idx] == r["rgba"].reshape(-1, 4)[idx]).all(axis=1).mean()
            line["cpu_baseline"] = {"value": rays / (r["render_ms"] * 1e-3) / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
                                    "sample": f"capture_subset(k, n={n}) for k in 0..{threads} of the {w}x{h} frame ({rays} rays, {r['render_ms']:.0f} ms); "
                                              "C++ restatement of the reference algorithm, not the Rust build",
                                    "bvh_build_ms": osc.build_ms, "device_film_identical_frac_at_sample": float(same)}
        print(json.dumps(line))
    dev.destroy()
    if shared is not None:
        torch.cuda.synchronize(); dist.barrier()
        if rank != 0:
            shared.close()
        dist.barrier()
        if rank == 0:
            shared.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"], help="N>1 film gather: peer stores into rank 0's film, or NCCL reduce")
    ap.add_argument("--workload", default="mixed4k", choices=sorted(scenes.CONFIGS))
    ap.add_argument("--small", action="store_true", help="tiny variant of mixed4k (CPU smoke of the bench logic)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--ref-seconds", type=float, default=60.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
