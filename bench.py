#!/usr/bin/env python
"""bench.py — headline benchmark of the path-tracing hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4|...]

Metric (BASELINE.json): Mpaths/s (and Grays/s) at 1920x1080, 8 bounces. A "step" is one pass of the
hot path over one batch: the full config-2 render (sample scene, 1024 spp) on each GPU. With N GPUs
every rank renders its own 1024 frame indices of an N*1024-spp image (spp split, weak scaling) and
the float4 accumulation buffers are summed across the ranks inside the timed step.

 value : whole-job Mpaths/s, scene resident in HBM, device-timed (CUDA events on the renderer's
         stream, max over ranks), L2 flushed between steps.
 e2e   : same metric through the public API with HOST buffers: scene upload (pinned host memory),
         camera, reset, render, RGBA8 read-back to pinned host memory — every step, wall-clock.
 roofline : FP32 FMA issue (no stage is a dense contraction — no tensor cores): algorithmic flops
         = 19 per ray-sphere test + 7 per ray, from exact device counters.
 verified : the timed renders are checked after the timed region: per-pixel sample counts on every
         rank, at N=1 the SHA-256 of the config-2 accumulation buffer against the digest of the
         unmodified reference CUDA renderer's 1024 frames (tests/golden/c2_full.json), at N>1 the
         reduced buffer against the same frames rendered sequentially on rank 0.
 cpu_baseline : the reference's per-pixel shading compiled for the host (oracle/_ref/libref_cpu.so,
         kind "reference") or the oracle port, all host threads, bounded sample.
 ref_cuda_baseline : the reference's own CUDA renderer (oracle/_ref/ref_headless) on this GPU, per config:
         end-to-end Render() ms per frame and kernel-only ms of kernelRender (BASELINE.md §4 Baseline A).
 --impl reference : times the CPU implementation as its own arm (rank 0 only); that process never
         loads the product library.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # SMs x FP32 lanes x 2 flop x max SM clock = 74.4
FLOP_PER_TEST, FLOP_PER_RAY = 19, 7
SAMPLE_SCENE = ROOT / "tests" / "golden" / "sample_scene.json"

WORKLOADS = {
    # name: (description, width, height, spp, bounces)
    "c2": ("sample-scene-data/scene.json 1920x1080 1024spp 8 bounces", 1920, 1080, 1024, 8),
    "c3": ("synthetic 256 spheres 16 lights 3840x2160 256spp 8 bounces", 3840, 2160, 256, 8),
    "c4": ("synthetic 4096 spheres 3840x2160 64spp 8 bounces", 3840, 2160, 64, 8),
    "c1": ("sample-scene-data/scene.json 1280x720 1spp 5 bounces", 1280, 720, 1, 5),
    # config 5: the TOTAL spp is fixed and split across the ranks (strong scaling), float4 buffers summed across ranks
    "c5": ("sample-scene-data/scene.json 7680x4320 16384spp split across the GPUs, 8 bounces", 7680, 4320, 16384, 8),
    "cs": ("synthetic 12 spheres 3 lights 1920x1080 256spp 8 bounces (small scene, several lights)", 1920, 1080, 256, 8),
    "c16k": ("synthetic 16384 spheres (chunked TMA staging) 1920x1080 16spp 8 bounces", 1920, 1080, 16, 8),
    # strong-scaling jobs (north_star: "near-linear 8-GPU scaling on 4K/1024-spp renders"): total spp fixed, split
    "s2": ("sample-scene-data/scene.json 1920x1080 1024spp IN TOTAL split across the GPUs, 8 bounces", 1920, 1080, 1024, 8),
    "s4k": ("sample-scene-data/scene.json 3840x2160 1024spp IN TOTAL split across the GPUs, 8 bounces", 3840, 2160, 1024, 8),
    "s3": ("synthetic 256 spheres 16 lights 3840x2160 1024spp IN TOTAL split across the GPUs, 8 bounces", 3840, 2160, 1024, 8),
}
STRONG = ("c5", "s2", "s4k", "s3")
SCENE_OF = {"c1": "sample", "c2": "sample", "c5": "sample", "s2": "sample", "s4k": "sample", "c3": "config3", "s3": "config3",
            "c4": "config4", "cs": "small", "c16k": "stress16k"}
# frames of the unmodified reference CUDA renderer per config (its brute-force kernel needs 0.1-1 s per 4K frame)
REF_CUDA_FRAMES = {"c1": 48, "c2": 24, "c3": 6, "c4": 3}


_SCENES = {}


def load_scene(atx, name):
    kind = SCENE_OF[name]
    if kind not in _SCENES:      # the synthetic generators place thousands of spheres by rejection sampling: once per process
        if kind == "sample":
            _SCENES[kind] = atx.Utils.importScene(str(SAMPLE_SCENE))
        elif kind == "small":
            _SCENES[kind] = atx.synthetic.small(12, 3, seed=9)
        else:
            _SCENES[kind] = getattr(atx.synthetic, kind)()
    return _SCENES[kind]


def scene_file(name, tmpdir, atx=None):
    """Path of the workload's scene in the reference's scene.json schema. Without `atx` (the reference arm, which
    must not load the product library) a synthetic scene is written by a child process."""
    if SCENE_OF[name] == "sample":
        return SAMPLE_SCENE
    path = Path(tmpdir) / f"{SCENE_OF[name]}.json"
    if not path.exists():
        if atx is not None:
            atx.Utils.exportScene(load_scene(atx, name), str(path))
        else:
            code = ("import sys; sys.path.insert(0, %r); import ataraxia_b200 as atx; import bench; "
                    "atx.Utils.exportScene(bench.load_scene(atx, %r), %r)" % (str(ROOT), name, str(path)))
            subprocess.run([sys.executable, "-c", code], check=True, timeout=600)
    return path


def config_dict(name, world, spp, n_spheres, n_lights, split="spp"):
    """The workload description both arms print (identical keys and values)."""
    desc, W, H, _, bounces = WORKLOADS[name]
    return {"workload": desc, "width": W, "height": H, "spp_per_gpu": spp, "max_bounces": bounces,
            "spheres": int(n_spheres), "lights": int(n_lights), "parallelism": f"{split}-split x{world}",
            "l2": "GPU arm: flushed between steps (256 MB write)"}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return json.loads(p.read_text())
        except Exception:
            pass
    return {}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        """Start of the timed region: nvidia-smi is already running (it takes ~0.1 s to produce its first
        row), only rows from here on count."""
        self.rows = []

    def summary(self):
        sm, smax, reasons = [], [], set()
        for r in list(self.rows):
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        return self.summary()


def cpu_reference_arm(workload, steps, warmup, threads=0, target_seconds=10.0):
    """The reference's per-pixel path on the host cores: oracle/_ref/libref_cpu.so (the reference's own
    sources compiled host-side: its importer, its scene flatten, its camera, its perPixel) when present,
    else the oracle port. Bounded sample of the workload per step. Loads nothing of the product."""
    from oracle import bindings as ob
    desc, W, H, spp, bounces = WORKLOADS[workload]
    with tempfile.TemporaryDirectory() as td:
        path = scene_file(workload, td)
        if ob.have_reference_cpu():
            impl, kind = ob.ReferenceCpu(), "reference"
            s, m, l, info = impl.load_scene(path)
            s["material"] = np.where((s["material"] < 0) | (s["material"] >= len(m)), 0, s["material"])  # Renderer.cu:30-37
            pos = info["position"]
            rays, _, _ = impl.camera(pos, info["direction"], info["fov"], 0.1, 100.0, W, H)
        else:
            impl, kind = ob.OraclePort(), "port"
            j = json.loads(Path(path).read_text())
            s = impl.flatten_json(j)
            m = np.zeros(len(j["materials"]), ob.MATERIAL_DTYPE)
            for i, mm in enumerate(j["materials"]):
                m[i] = (mm["albedo"], mm["roughness"], mm["metallic"], mm["F0"], mm["emissionColor"], mm["emissionIntensity"], 0)
            l = np.zeros(len(j["lights"]), ob.LIGHT_DTYPE)
            for i, ll in enumerate(j["lights"]):
                l[i] = (ll["position"], ll["color"], ll["intensity"])
            pos = np.asarray(j["camera"]["position"], np.float32)
            rays, _, _ = impl.camera(pos, j["camera"]["direction"], j["camera"]["fov"], 0.1, 100.0, W, H)
    cores = impl.hardware_threads()
    # calibrate: one frame over a thin band, then size the sample to ~target_seconds per step
    band = max(cores, H // 8)
    t0 = time.perf_counter()
    impl.render(s, m, l, pos, rays, 1, 1, 1, bounces, False, rows=(H // 2 - band // 2, H // 2 - band // 2 + band), threads=threads)
    dt = max(time.perf_counter() - t0, 1e-4)
    rate = band * W / dt
    frames = int(max(1, min(spp, target_seconds * rate / (W * H))))
    rows = (0, H)
    if frames == 1 and W * H / rate > 2 * target_seconds:   # even one full frame is too slow: band of rows
        nrows = int(max(cores, min(H, target_seconds * rate / W)))
        rows = (H // 2 - nrows // 2, H // 2 - nrows // 2 + nrows)
    paths = (rows[1] - rows[0]) * W * frames
    times = []
    for i in range(warmup + steps):
        acc = np.zeros((H, W, 4), np.float32)
        t0 = time.perf_counter()
        impl.render(s, m, l, pos, rays, 1, frames, 1, bounces, False, accum=acc, rows=rows, threads=threads)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    sample = f"{W}x{rows[1] - rows[0]} rows of {W}x{H}, {frames} of {spp} spp, {bounces} bounces"
    return {"value": paths / sec / 1e6, "unit": "Mpaths/s", "cores": cores, "kind": kind, "sample": sample,
            "ms_per_step": sec * 1e3, "paths_per_step": paths, "spheres": len(s), "lights": len(l)}


def ref_cuda_baseline(workload, tmpdir, atx=None):
    """BASELINE.md §4 Baseline A: the reference's own CUDA renderer on this GPU (oracle/_ref/ref_headless, the
    unmodified sources built for sm_100a), one Render() per 1-spp frame, reduced frame count: (i) end-to-end
    ms per frame as the app's "Last Render Time" shows it and (ii) kernel-only ms of kernelRender (CUDA events
    around the launch, taken by the harness's cudaLaunchKernel interposer — the source is untouched)."""
    from oracle import bindings as ob
    if not ob.have_ref_headless() or workload not in REF_CUDA_FRAMES:
        return None
    desc, W, H, spp, bounces = WORKLOADS[workload]
    frames = REF_CUDA_FRAMES[workload]
    try:
        info, _ = ob.run_ref_headless(scene_file(workload, tmpdir, atx), W, H, bounces, False, frames, timeout=300)
    except Exception as e:  # reported baseline only: never fail the bench on it
        return {"error": str(e)[:200]}
    P = W * H
    out = {"kind": "reference-cuda (unmodified Renderer::Render, sm_100a build)", "config": desc, "frames": frames,
           "e2e_ms_per_frame": info["median_frame_ms"], "min_frame_ms": info["min_frame_ms"],
           "e2e_value": P / info["median_frame_ms"] / 1e3, "unit": "Mpaths/s",
           "note": "e2e = Render() wall time per 1-spp frame (host ray-table upload + sync + RGBA8 read-back included, as the app "
                   "runs it); kernel = CUDA events around the kernelRender launch; frames reduced, per-frame cost is constant"}
    if info.get("kernel_launches"):
        out["kernel_ms_per_frame"] = info["median_kernel_ms"]
        out["kernel_value"] = P / info["median_kernel_ms"] / 1e3
        out["kernel_launches"] = info["kernel_launches"]
    return out


def make_renderer(atx, name, local_rank, args=None):
    desc, W, H, spp, bounces = WORKLOADS[name]
    scene = load_scene(atx, name)
    cam = atx.Camera(scene.camera.getFov(), 0.1, 100.0, scene.camera.getPosition(), scene.camera.getDirection())
    r = atx.Renderer(local_rank)
    r.setSettings(atx.Settings(True, False, bounces))
    r.variant = atx.VARIANT_MEGAKERNEL
    if args is not None:
        r.variant = {"megakernel": atx.VARIANT_MEGAKERNEL, "wavefront": atx.VARIANT_WAVEFRONT, "auto": atx.VARIANT_AUTO}[args.variant]
        r.setTuning(atx.TUNE_MEGA_KIND, args.mega_kind)
        if args.park_threshold:
            r.setTuning(atx.TUNE_PARK_THRESHOLD, args.park_threshold)
        if args.chunk:
            r.setTuning(atx.TUNE_CHUNK_SPHERES, args.chunk)
        if args.claim_threshold:
            r.setTuning(atx.TUNE_CLAIM_THRESHOLD, args.claim_threshold)
        if args.reduce == "nccl":
            r.setTuning(atx.TUNE_REDUCE, 1)
    r.onResize(W, H)
    cam.Resize(W, H)
    spheres = atx.pack_spheres(atx.traverseSceneGraph(scene.rootNode))
    mats, lights = atx.pack_materials(scene.materials), atx.pack_lights(scene.lights)
    r.uploadArrays(spheres, mats, lights)
    r.setCamera(cam)
    return r, cam, scene, spheres, mats, lights


FORMS = {0: "wavefront", 1: "while-while", 2: "two-slot packed", 3: "warp-queue", 4: "two-slot packed, lockstep"}


def secondary_workload(atx, name, local_rank, flush, steps=2, warmup=3):
    """One of the other BASELINE configs on the same GPU, device-timed like the headline, with its own clock
    record: reported next to it (the headline scene has 3 spheres, so its FP32 fraction is small by
    construction; configs 3 and 4 are the ones the sphere loop dominates)."""
    import torch
    desc, W, H, spp, bounces = WORKLOADS[name]
    r, cam, scene, spheres, mats, lights = make_renderer(atx, name, local_rank)
    sampler = ClockSampler(local_rank).start()
    for _ in range(warmup):
        r.renderFrames(1, max(1, spp // 8), 1, zero_first=True)
    r.sync()
    r.resetCounters()
    sampler.mark()
    ms = []
    for _ in range(steps):
        flush.fill_(0)
        torch.cuda.synchronize()
        r.eventRecord(0)
        r.renderFrames(1, spp, 1, zero_first=True)
        r.eventRecord(1)
        ms.append(r.eventElapsedMs(0, 1))
    clocks = sampler.stop()
    c = r.counters()
    acc = r.getAccumulation()
    counts_ok = bool((acc[..., 3] == spp).all())
    # one end-to-end step: scene upload from host arrays, camera, render, RGBA8 read-back to the renderer's pinned image
    t0 = time.perf_counter()
    r.uploadArrays(spheres, mats, lights)
    r.setCamera(cam)
    r.renderFrames(1, spp, 1, zero_first=True)
    r.getRGBA8(divisor=spp, out=r.getImage().data)
    e2e_ms = (time.perf_counter() - t0) * 1e3
    total = sum(ms) * 1e-3
    flops = FLOP_PER_TEST * float(c.sphere_tests_executed) + FLOP_PER_RAY * float(c.rays_traced)
    out = {"workload": desc, "spheres": int(len(spheres)), "lights": int(len(scene.lights)), "steps": steps,
           "warmup": f"{warmup} x {max(1, spp // 8)} spp", "ms_per_step": sum(ms) / steps,
           "form": FORMS.get(r.lastMegaKind(), "?"), "clocks": clocks, "verified": {"sample_counts": counts_ok},
           "e2e_ms_per_step": e2e_ms,
           "value": float(c.paths) / total / 1e6, "unit": "Mpaths/s", "grays_per_s": float(c.rays) / total / 1e9,
           "roofline": {"bound": "fp32_fma", "achieved": flops / total / 1e12, "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s",
                        "frac": flops / total / 1e12 / FP32_PEAK_TFLOPS}}
    r.close()
    return out


def strong_workload(atx, name, rank, world, local_rank, flush, dist, steps=3, warmup=3, args=None, split="spp"):
    """A fixed job (total spp fixed) split across the ranks: frames rank+1, rank+1+world, ... on every GPU, buffers
    summed across the ranks inside the timed step; device-timed, max over ranks. The driver's own scaling curve is
    the weak-scaling headline; these lines say how much faster ONE job gets."""
    import torch
    from ataraxia_b200.distributed import frame_partition
    desc, W, H, total_spp, bounces = WORKLOADS[name]
    r, cam, scene, spheres, mats, lights = make_renderer(atx, name, local_rank)
    if args is not None and args.reduce == "nccl":
        r.setTuning(atx.TUNE_REDUCE, 1)
    if world > 1:
        uid = [atx.Renderer.commUniqueId() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        r.commInitRank(world, rank, uid[0])
    sh = frame_partition(total_spp, rank, world)

    tiles = split == "tiles"

    def step():
        r.eventRecord(0)
        if tiles:
            r.renderTiles(1, total_spp, zero_first=True)
        else:
            r.renderFrames(sh.first, sh.count, sh.stride, zero_first=True)
            if world > 1:
                r.allreduceAccum()
        r.eventRecord(1)

    for _ in range(warmup):
        step()
    r.sync()
    ms = []
    for _ in range(steps):
        flush.fill_(0)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        step()
        ms.append(r.eventElapsedMs(0, 1))
    total_ms = sum(ms)
    acc = r.getAccumulation()
    counts_ok = bool((acc[..., 3] == total_spp).all())
    reduce_kind = r.lastReduceKind()
    if world > 1:
        t = torch.tensor([total_ms, 0.0 if counts_ok else 1.0], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, counts_ok = float(t[0].item()), float(t[1].item()) == 0.0
        r.commDestroy()
    r.close()
    paths = float(W) * H * total_spp * steps
    return {"workload": desc, "scaling": "strong", "n_gpus": world, "split": split, "spp_total": total_spp, "spp_this_rank": total_spp if tiles else sh.count,
            "steps": steps, "ms_per_step": total_ms / steps, "value": paths / (total_ms * 1e-3) / 1e6, "unit": "Mpaths/s",
            "reduce": {0: None, 1: "peer memory", 2: "ncclAllReduce"}.get(reduce_kind), "verified": {"sample_counts_all_ranks": counts_ok}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel per step (default: the workload's)")
    ap.add_argument("--no-baselines", action="store_true", help="skip cpu_baseline / reference-CUDA legs and the secondary workloads")
    ap.add_argument("--variant", default="megakernel", choices=["megakernel", "wavefront", "auto"],
                    help="kernel family (bit-identical results); auto = atx_calibrate's pick")
    ap.add_argument("--mega-kind", type=int, default=0, help="0 auto, 1 while-while, 2 two-slot packed, 3 warp-queue (same results)")
    ap.add_argument("--park-threshold", type=int, default=0, help="while-while form: parked hits per warp that trigger the bounce phase")
    ap.add_argument("--claim-threshold", type=int, default=0, help="idle lanes per warp that trigger a batched pixel claim")
    ap.add_argument("--chunk", type=int, default=0, help="force the shared-memory chunk size in spheres (0 = automatic)")
    ap.add_argument("--split", default="spp", choices=["spp", "tiles"],
                    help="multi-GPU partitioning: spp = every rank renders its own frame indices of the whole image, buffers summed; "
                         "tiles = every rank renders all frames of its interleaved 8x4 tiles and stores them into every rank's image")
    ap.add_argument("--reduce", default="auto", choices=["auto", "nccl"], help="cross-GPU sum: peer-memory kernel when possible, or ncclAllReduce")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    desc, W, H, spp, bounces = WORKLOADS[args.workload]
    if args.spp:
        spp = args.spp
    strong = args.workload in STRONG
    total_spp = spp if strong else spp * world
    if strong:
        spp = max(1, spp // world)   # per-GPU share of the fixed total

    if args.impl == "reference":
        if rank != 0:
            return 0
        res = cpu_reference_arm(args.workload, max(args.steps, 1), max(args.warmup, 0),
                                target_seconds=float(os.environ.get("ATX_BENCH_CPU_SECONDS", "10")))
        try:   # this arm is the reference's code only: the product library must not even be mapped
            product_loaded = any("libataraxia_b200" in ln for ln in open("/proc/self/maps"))
        except OSError:
            product_loaded = None
        line = {"impl": "reference", "metric": "Mpaths/s", "value": res["value"], "unit": "Mpaths/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
                "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_dict(args.workload, world, spp, res["spheres"], res["lights"], args.split),
                "cpu_baseline": {"value": res["value"], "unit": "Mpaths/s", "cores": res["cores"], "kind": res["kind"],
                                 "sample": res["sample"]},
                "e2e": {"value": res["value"], "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "product_library_loaded": product_loaded}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    import ataraxia_b200 as atx

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    r, cam, scene, spheres, mats, lights = make_renderer(atx, args.workload, local_rank, args)
    if world > 1:
        uid = [atx.Renderer.commUniqueId() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        r.commInitRank(world, rank, uid[0])

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def render_share():
        if args.split == "tiles":
            # all total_spp frames of this rank's tiles, stored into every rank's image over NVLink while rendering
            r.renderTiles(1, total_spp, zero_first=True)
        else:
            # rank's share of the total_spp-spp image: frame indices rank+1, rank+1+world, ...; buffers summed across ranks
            r.renderFrames(rank + 1, spp, world, zero_first=True)
            if world > 1:
                r.allreduceAccum()

    def step():
        r.eventRecord(0)
        render_share()
        r.eventRecord(1)

    calibration = None
    if args.variant == "auto":
        calibration = r.calibrate(2)
    sampler = ClockSampler(local_rank).start()
    for _ in range(args.warmup):
        step()
    r.sync()
    r.resetCounters()
    barrier()
    sampler.mark()
    wall0 = time.perf_counter()
    dev_ms = []
    for _ in range(args.steps):
        flush.fill_(0)            # L2 flush between timed iterations (not inside the event bracket)
        barrier()                 # ranks enter the step together: the reduce inside it must not time their skew
        step()
        dev_ms.append(r.eventElapsedMs(0, 1))
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    clocks = sampler.stop()
    c = r.counters()
    form = FORMS.get(r.lastMegaKind(), "?")
    reduce_kind = {0: None, 1: "one kernel over NVLink peer memory (two-shot, rank-ordered sums)", 2: "ncclAllReduce"}.get(r.lastReduceKind())

    # ---- verification of what the timed region rendered (untimed) ------------------------------------
    acc = r.getAccumulation()
    verified = {"sample_counts": bool((acc[..., 3] == total_spp).all())}
    if world == 1 and args.workload == "c2" and spp == WORKLOADS["c2"][3]:
        gpath = ROOT / "tests" / "golden" / "c2_full.json"
        if gpath.exists():
            want = json.loads(gpath.read_text())["acc1024_sha256"]
            verified["acc_sha256_vs_reference_cuda"] = hashlib.sha256(np.ascontiguousarray(acc).tobytes()).hexdigest() == want
    if world > 1:
        # every rank holds the same reduced buffer; rank 0 compares it with the same frames rendered sequentially
        t = torch.from_numpy(acc).to(f"cuda:{local_rank}")
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        verified["all_ranks_hold_the_same_buffer"] = bool((lo == hi).all().item())
        del t, lo, hi
        if rank == 0:
            r.renderFrames(1, total_spp, 1, zero_first=True)
            seq = r.getAccumulation()
            err = np.abs(acc[..., :3] - seq[..., :3]) / np.maximum(np.abs(seq[..., :3]), 1e-3)
            verified["reduced_vs_sequential_max_rel"] = float(err.max())
            if args.split == "tiles":   # every pixel summed on one GPU in frame order: the single-GPU bits
                verified["tiles_bit_identical_to_sequential"] = bool((acc.view(np.uint32) == seq.view(np.uint32)).all())
            else:
                # N partial sums added in rank order against one sequential sum: float reassociation only. Both are within
                # (n - 1) u of the exact sum of n non-negative samples (u = 2^-24), so they differ by at most 2 n u relative
                # (8 ranks x 1024 spp measured: 2.1e-4 against the bound 9.8e-4); a missing or doubled frame shows in sample_counts
                verified["reduced_vs_sequential_bound"] = 2.0 * total_spp * 2.0 ** -24 + 2e-6
                verified["reduced_vs_sequential"] = bool(err.max() <= verified["reduced_vs_sequential_bound"])
        ok = torch.tensor([0.0 if verified["sample_counts"] else 1.0], device=f"cuda:{local_rank}")
        dist.all_reduce(ok, op=dist.ReduceOp.MAX)
        verified["sample_counts"] = float(ok.item()) == 0.0

    total_ms = sum(dev_ms)
    if world > 1:
        t = torch.tensor([total_ms], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        cnt = torch.tensor([c.paths, c.rays, c.sphere_tests, c.rays_traced, c.sphere_tests_executed],
                           device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(cnt)
        paths, rays, tests, rays_x, tests_x = (float(v) for v in cnt.tolist())
    else:
        paths, rays, tests = float(c.paths), float(c.rays), float(c.sphere_tests)
        rays_x, tests_x = float(c.rays_traced), float(c.sphere_tests_executed)
    launches = int(c.launches)

    # ---- e2e: public API, host buffers, H2D + D2H inside the timed region -------------------------
    pin = lambda a: torch.from_numpy(a.view(np.uint8).copy()).pin_memory()  # noqa: E731
    hs, hm, hl = pin(spheres), pin(mats), pin(lights)
    host_rgba = torch.empty((H, W), dtype=torch.int32).pin_memory()
    rgba_np = host_rgba.numpy().view(np.uint32)
    e2e_steps = max(1, args.steps)

    def e2e_step():
        r.uploadArrays(hs.numpy().view(atx.SPHERE_DTYPE), hm.numpy().view(atx.MATERIAL_DTYPE), hl.numpy().view(atx.LIGHT_DTYPE))
        r.setCamera(cam)
        render_share()
        r.getRGBA8(divisor=total_spp, out=rgba_np)   # resolve + D2H; waits for the stream

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_paths = float(W) * H * total_spp * e2e_steps
    verified["e2e_alpha_255"] = bool(((rgba_np >> 24) == 255).all())

    # ---- strong-scaling jobs next to the weak headline (all ranks take part) ----------------------
    strong_lines = None
    if args.workload == "c2" and not args.no_baselines and not args.spp:
        strong_lines = {f"{k}_{sp}": strong_workload(atx, k, rank, world, local_rank, flush, dist, args=args, split=sp)
                        for k in ("s2", "s4k", "s3") for sp in (("tiles", "spp") if world > 1 else ("spp",))}

    if rank == 0:
        peaks = measured_peaks()
        # roofline numerator: the sphere tests the device really executed (the per-pixel primary hit
        # is traced once per launch and reused by every frame); the reference-equivalent count,
        # which includes those reused primary rays, is reported next to it
        flops = FLOP_PER_TEST * tests_x + FLOP_PER_RAY * rays_x
        achieved = flops / (total_ms * 1e-3) / 1e12 / world      # per GPU
        achieved_ref = (FLOP_PER_TEST * tests + FLOP_PER_RAY * rays) / (total_ms * 1e-3) / 1e12 / world
        traffic = None
        tpath = ROOT / "profiles" / "traffic.json"
        if tpath.exists():
            try:
                traffic = json.loads(tpath.read_text()).get(args.workload)
            except Exception:
                traffic = None
        ncu_issue = None
        ipath = ROOT / "profiles" / "issue.json"
        if ipath.exists():
            try:
                ncu_issue = json.loads(ipath.read_text()).get(args.workload)
            except Exception:
                ncu_issue = None
        if ncu_issue and ncu_issue.get("issue_slots_busy_pct") and ncu_issue.get("active_lanes_per_instruction"):
            # the honest lens for a 3-sphere scene: the share of lane-issue slots doing work
            ncu_issue["lane_issue_utilisation"] = ncu_issue["issue_slots_busy_pct"] / 100.0 * ncu_issue["active_lanes_per_instruction"] / 32.0
        ms_per_step = total_ms / args.steps
        line = {
            "metric": "Mpaths/s", "value": paths / (total_ms * 1e-3) / 1e6, "unit": "Mpaths/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.workload, world, spp, len(spheres), len(lights), args.split),
            "kernel": {"variant": args.variant, "mega_kind": args.mega_kind, "form": form, "park_threshold": args.park_threshold,
                       "chunk": args.chunk, "reduce": reduce_kind},
            "verified": all(v for k, v in verified.items() if isinstance(v, bool)), "verification": verified,
            "calibration_ms": calibration,
            "grays_per_s": rays / (total_ms * 1e-3) / 1e9,
            "grays_traced_per_s": rays_x / (total_ms * 1e-3) / 1e9,
            "wall_ms_total": wall_ms,
            "gpu_launches": launches,
            "clocks": clocks,
            "e2e": {"value": e2e_paths / e2e_s / 1e6, "unit": "Mpaths/s",
                    "h2d_bytes_per_step": int(spheres.nbytes + mats.nbytes + lights.nbytes + 2 * 64 + 12),
                    "d2h_bytes_per_step": int(W * H * 4), "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3},
            "roofline": {"bound": "fp32_fma", "achieved": achieved, "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s",
                         "frac": achieved / FP32_PEAK_TFLOPS, "traffic": traffic, "ncu": ncu_issue,
                         "peak_source": "148 SMs x 128 FP32 lanes x 2 x clocks.max.sm 1965 MHz (MEASURED_PEAKS.json sm_max_mhz); "
                                        "no tensor-core or HBM bound applies to this kernel",
                         "algorithmic": "19 flop per executed ray-sphere test + 7 per traced ray, exact device counters",
                         "reference_equivalent": {"achieved": achieved_ref, "frac": achieved_ref / FP32_PEAK_TFLOPS,
                                                  "note": "counts every traceRay call of the reference, including the per-frame primary rays this kernel traces once per launch"},
                         "accum_hbm": {"bytes_per_pixel_per_launch": 16, "gbs": 16.0 * W * H / (ms_per_step * 1e-3) / 1e9,
                                       "peak_gbs": peaks.get("hbm_gbs")}},
        }
        if strong_lines:
            line["strong_scaling"] = strong_lines
        if world == 1 and not args.no_baselines:
            t_leg = time.perf_counter()

            def leg(name):
                nonlocal t_leg
                print(f"[bench] {name}: {time.perf_counter() - t_leg:.1f} s", file=sys.stderr, flush=True)
                t_leg = time.perf_counter()

            cb = cpu_reference_arm(args.workload, 1, 0)
            leg("cpu_baseline")
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            # BASELINE metric part 2: ms/frame at 1 spp (config 1), end to end through Renderer::Render
            s1 = load_scene(atx, "c1")
            cam1 = atx.Camera(s1.camera.getFov(), 0.1, 100.0, s1.camera.getPosition(), s1.camera.getDirection())
            r1 = atx.Renderer(local_rank)
            r1.setSettings(atx.Settings(True, False, 5))
            r1.onResize(1280, 720); cam1.Resize(1280, 720)
            for _ in range(5):
                r1.Render(cam1, s1)
            t0 = time.perf_counter()
            for _ in range(50):
                r1.Render(cam1, s1)
            line["ms_per_frame_1spp"] = {"value": (time.perf_counter() - t0) / 50 * 1e3, "unit": "ms",
                                         "config": WORKLOADS["c1"][0], "kernel_ms": r1.lastRenderMs(),
                                         "includes": "Render(): camera, launch, RGBA8 read-back to host"}
            r1.close()
            leg("ms_per_frame_1spp")
            # the other BASELINE configs on this GPU (the ones the sphere loop dominates)
            if args.workload == "c2":
                line["other_workloads"] = {k: secondary_workload(atx, k, local_rank, flush) for k in ("c3", "c4")}
                leg("other_workloads")
            # Baseline A: the reference's CUDA renderer per config, e2e and kernel-only, with our per-frame figures beside it
            with tempfile.TemporaryDirectory() as td:
                names = ("c1", "c2", "c3", "c4") if args.workload == "c2" else (args.workload,)
                ours = {args.workload: {"kernel_ms_per_frame": ms_per_step / spp, "e2e_ms_per_frame": e2e_s / e2e_steps * 1e3 / spp}}
                ours["c1"] = {"kernel_ms_per_frame": line["ms_per_frame_1spp"]["kernel_ms"], "e2e_ms_per_frame": line["ms_per_frame_1spp"]["value"]}
                for k, v in line.get("other_workloads", {}).items():
                    ours[k] = {"kernel_ms_per_frame": v["ms_per_step"] / WORKLOADS[k][3], "e2e_ms_per_frame": v["e2e_ms_per_step"] / WORKLOADS[k][3]}
                rcs = {}
                for k in names:
                    rc = ref_cuda_baseline(k, td, atx)
                    if not rc:
                        continue
                    if "error" not in rc and k in ours:
                        rc["ours"] = ours[k]
                        if rc.get("kernel_ms_per_frame"):
                            rc["speedup_vs_ref_cuda_kernel"] = rc["kernel_ms_per_frame"] / ours[k]["kernel_ms_per_frame"]
                        if "e2e_ms_per_frame" in ours[k]:
                            rc["speedup_vs_ref_cuda_e2e"] = rc["e2e_ms_per_frame"] / ours[k]["e2e_ms_per_frame"]
                    rcs[k] = rc
                if rcs:
                    line["ref_cuda_baseline"] = rcs
            leg("ref_cuda_baseline")
        print(json.dumps(line))
    r.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
