"""Generates the CPU golden vectors from the UNMODIFIED reference compiled for the host
(oracle/_ref/libref_cpu.so; needs /root/reference to have been built by oracle/ref/Makefile).

    python tests/golden/make_golden_cpu.py

Writes, under tests/golden/:
  sample_scene.json      the reference's sample scene re-exported by the reference's own
                         Utils::exportScene (import -> export round trip)
  small_scene.json       the small synthetic scene, written by THIS repo's exporter (the reference
                         importer reads it back identically — checked here)
  cpu_golden.npz         PCG vectors, flattened spheres, camera matrices, ray tables, primary hit
                         maps and accumulation buffers produced by the reference's host-compiled
                         perPixel / traceRay / Camera / traverseSceneGraph
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.bindings import ReferenceCpu, REF_SCENE  # noqa: E402
import ataraxia_b200 as atx  # noqa: E402

G = Path(__file__).resolve().parent


def main():
    ref = ReferenceCpu()
    out = {}
    ref.reexport_scene(REF_SCENE, G / "sample_scene.json")
    atx.Utils.exportScene(atx.synthetic.small(), str(G / "small_scene.json"))

    seeds = np.array([0, 1, 2, 12345, 921599, 0xFFFFFFFF, 0x80000000, 747796405, 2891336453, 65536], np.uint64)
    out["pcg_seeds"] = seeds
    out["pcg_hash"] = np.array([ref.pcg_hash(int(s)) for s in seeds], np.uint64)
    chain_f, chain_s, s = [], [], 0
    for _ in range(16):
        f, s = ref.pcg_float(s)
        chain_f.append(f)
        chain_s.append(s)
    out["pcg_chain_float"] = np.array(chain_f, np.float32)
    out["pcg_chain_seed"] = np.array(chain_s, np.uint64)

    for name, path, W, H, bounces, sky, frames in [("sample", G / "sample_scene.json", 96, 54, 5, False, 3),
                                                   ("small", G / "small_scene.json", 64, 36, 8, True, 2)]:
        s_, m_, l_, info = ref.load_scene(path)
        rays, ip, iv = ref.camera(info["position"], info["direction"], info["fov"], 0.1, 100.0, W, H)
        out[f"{name}_spheres"] = s_.view(np.uint8)
        out[f"{name}_materials"] = m_.view(np.uint8)
        out[f"{name}_lights"] = l_.view(np.uint8)
        out[f"{name}_campos"] = info["position"]
        out[f"{name}_camdir"] = info["direction"]
        out[f"{name}_fov"] = np.float32(info["fov"])
        out[f"{name}_dims"] = np.array([W, H, bounces, int(sky), frames], np.int32)
        out[f"{name}_invproj"] = ip
        out[f"{name}_invview"] = iv
        out[f"{name}_rays"] = rays
        out[f"{name}_hits"] = ref.primary_hits(s_, info["position"], rays).astype(np.int16)
        acc = ref.render(s_, m_, l_, info["position"], rays, 1, 1, 1, bounces, sky)
        out[f"{name}_acc1"] = acc.copy()
        acc = ref.render(s_, m_, l_, info["position"], rays, 2, frames - 1, 1, bounces, sky, accum=acc)
        out[f"{name}_accK"] = acc
        out[f"{name}_rgbaK"] = ref.pack_rgba8(acc, frames)
    # C1 primary visibility histogram at full size (SURVEY.md §8c sanity numbers)
    s_, m_, l_, info = ref.load_scene(G / "sample_scene.json")
    rays, _, _ = ref.camera(info["position"], info["direction"], info["fov"], 0.1, 100.0, 1280, 720)
    hits = ref.primary_hits(s_, info["position"], rays)
    out["c1_hit_histogram"] = np.array([(hits == k).sum() for k in (-1, 0, 1, 2)], np.int64)
    out["c1_center_ray"] = rays[360, 640]
    # Camera::onUpdate (Camera.cpp:30-108) under a scripted input sequence: 160 steps of keys, mouse motion and
    # right-button state (fixed seed), starting from the sample scene's camera
    from oracle.bindings import CAMERA_STEP_DTYPE
    rng = np.random.default_rng(0xCA3E7A)
    n = 160
    steps = np.zeros(n, CAMERA_STEP_DTYPE)
    steps["dt"] = rng.uniform(0.001, 0.05, n).astype(np.float32)
    steps["mouse_x"] = np.cumsum(rng.normal(0, 25, n)).astype(np.float32)
    steps["mouse_y"] = np.cumsum(rng.normal(0, 15, n)).astype(np.float32)
    steps["keys"] = rng.integers(0, 64, n)
    steps["right"] = rng.random(n) < 0.8
    steps["mouse_x"][40:44] = steps["mouse_x"][39]          # a few steps without mouse motion
    steps["mouse_y"][40:44] = steps["mouse_y"][39]
    steps["keys"][40:42] = 0                                # ... and without keys: onUpdate returns false
    op, od, oiv, om, wrays = ref.camera_walk(info["position"], info["direction"], info["fov"], 0.1, 100.0, 48, 27, steps)
    out["walk_steps"] = steps.view(np.uint8)
    out["walk_pos"], out["walk_dir"], out["walk_invview"], out["walk_moved"], out["walk_rays"] = op, od, oiv, om, wrays
    np.savez_compressed(G / "cpu_golden.npz", **out)
    print("wrote", G / "cpu_golden.npz", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
