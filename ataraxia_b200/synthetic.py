"""Synthetic scenes of BASELINE.json configs 3 and 4, emitted in the reference's
own scene.json schema (a flat root node) so the reference reads the identical file.

C3: 256 spheres (255 random + one ground sphere), 32 mixed materials (2 emissive), 16 point lights.
C4: 4096 spheres stress case, 4 lights.
Generator: numpy default_rng(0xA7A2A41A); everything is rounded to float32 before use.
"""
from __future__ import annotations

import numpy as np

from .api import Camera, Light, Material, Scene, SceneNode, Settings, Sphere

SEED = 0xA7A2A41A


def _scatter(rng, n, lo, hi, rlo, rhi, max_tries=200):
    """Rejection-sampled non-overlapping spheres (grid-hashed so 4096 spheres stay fast)."""
    lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
    cell = 2.0 * rhi
    grid = {}
    centers, radii = [], []
    while len(centers) < n:
        placed = False
        for _ in range(max_tries):
            c = rng.uniform(lo, hi)
            r = rng.uniform(rlo, rhi)
            key = tuple(np.floor(c / cell).astype(int))
            ok = True
            for dx in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    for dz in (-1, 0, 1):
                        for j in grid.get((key[0] + dx, key[1] + dy, key[2] + dz), ()):
                            if np.linalg.norm(c - centers[j]) < r + radii[j]:
                                ok = False
                                break
            if ok:
                grid.setdefault(key, []).append(len(centers))
                centers.append(c)
                radii.append(r)
                placed = True
                break
        if not placed:  # give up on separation for this one (dense boxes)
            grid.setdefault(key, []).append(len(centers))
            centers.append(c)
            radii.append(r)
    return np.asarray(centers, np.float32), np.asarray(radii, np.float32)


def make_scene(n_spheres: int, n_lights: int, box_lo, box_hi, rlo, rhi, cam_pos, cam_dir, seed=SEED) -> Scene:
    rng = np.random.default_rng(seed)
    scene = Scene()
    root = scene.rootNode

    # 32 materials: albedo U[0.2,1]^3, roughness U[0.05,1], metallic 0 (50 %) or U(0,1] (50 %), F0 = 0.04, 2 emissive
    n_mat = 32
    for i in range(n_mat):
        albedo = rng.uniform(0.2, 1.0, 3)
        rough = rng.uniform(0.05, 1.0)
        metallic = 0.0 if rng.uniform() < 0.5 else 1.0 - rng.uniform(0.0, 1.0)  # (0, 1]
        m = Material(albedo=tuple(np.float32(albedo)), roughness=float(np.float32(rough)),
                     metallic=float(np.float32(metallic)), F0=(0.04, 0.04, 0.04))
        if i < 2:
            m.emissionColor = tuple(np.float32(rng.uniform(0.5, 1.0, 3)))
            m.emissionIntensity = 5.0
        scene.materials.append(m)

    centers, radii = _scatter(rng, n_spheres - 1, box_lo, box_hi, rlo, rhi)
    mats = rng.integers(0, n_mat, n_spheres)
    for i in range(n_spheres - 1):
        root.addSphere(Sphere(tuple(float(v) for v in centers[i]), float(radii[i]), int(mats[i])))
    # ground sphere, counted in n_spheres; a non-emissive material
    root.addSphere(Sphere((0.0, -1000.0, 0.0), 1000.0, int(2 + mats[-1] % (n_mat - 2))))

    lx = max(abs(box_lo[0]), abs(box_hi[0])) * 1.25
    for _ in range(n_lights):
        pos = rng.uniform((-lx, box_hi[1] + 2.0, -lx), (lx, box_hi[1] + 14.0, lx))
        col = rng.uniform(0.5, 1.0, 3)
        inten = rng.uniform(0.5, 2.0)
        scene.lights.append(Light(tuple(np.float32(pos)), tuple(np.float32(col)), float(np.float32(inten))))

    d = np.asarray(cam_dir, np.float64)
    d = d / np.linalg.norm(d)
    scene.camera = Camera()
    scene.camera.setPosition(np.float32(cam_pos))
    scene.camera.setDirection(np.float32(d))
    scene.camera.setFov(45.0)
    scene.settings = Settings(accumulation=True, skyLight=False, maxBounces=8)
    return scene


def config3() -> Scene:
    """256 spheres in [-12,12]x[0.3,6]x[-12,12], radii U[0.3,1.2], 16 lights, camera (0,6,28) -> (0,-0.15,-1)."""
    return make_scene(256, 16, (-12.0, 0.3, -12.0), (12.0, 6.0, 12.0), 0.3, 1.2, (0.0, 6.0, 28.0), (0.0, -0.15, -1.0))


def config4() -> Scene:
    """4096 spheres in [-60,60]x[0.3,20]x[-60,60], radii U[0.2,0.8], 4 lights, camera (0,25,110)."""
    return make_scene(4096, 4, (-60.0, 0.3, -60.0), (60.0, 20.0, 60.0), 0.2, 0.8, (0.0, 25.0, 110.0), (0.0, -0.15, -1.0))


def stress16k() -> Scene:
    """16384 spheres (beyond the resident shared-memory budget: the chunked, TMA-staged walk), same box as config 4."""
    return make_scene(16384, 4, (-60.0, 0.3, -60.0), (60.0, 20.0, 60.0), 0.15, 0.5, (0.0, 25.0, 110.0), (0.0, -0.15, -1.0))


def small(n_spheres=24, n_lights=3, seed=7) -> Scene:
    """Small mixed scene for fast parity tests (emissive + metallic + diffuse, several lights)."""
    return make_scene(n_spheres, n_lights, (-4.0, 0.3, -4.0), (4.0, 3.0, 4.0), 0.3, 0.9, (0.0, 3.0, 12.0),
                      (0.0, -0.15, -1.0), seed=seed)
