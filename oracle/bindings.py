"""TEST INFRASTRUCTURE — ctypes bindings of the two CPU oracles.

  OraclePort      oracle/liboracle.so        restatement written for this repo (oracle/oracle.cpp)
  ReferenceCpu    oracle/_ref/libref_cpu.so  the UNMODIFIED reference sources compiled for the host
                                             (exists only where oracle/ref/Makefile has been run)
  run_ref_headless                            oracle/_ref/ref_headless: the reference's CUDA renderer

Both CPU classes expose the same methods, so tests can pin the port against the reference.
"""
from __future__ import annotations

import ctypes as C
import json
import subprocess
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
PORT_LIB = HERE / "liboracle.so"
REFCPU_LIB = HERE / "_ref" / "libref_cpu.so"
REF_HEADLESS = HERE / "_ref" / "ref_headless"
REF_SCENE = HERE / "_ref" / "scene.json"

SPHERE_DTYPE = np.dtype([("center", "<f4", 3), ("radius", "<f4"), ("material", "<i4")])
MATERIAL_DTYPE = np.dtype([("albedo", "<f4", 3), ("roughness", "<f4"), ("metallic", "<f4"), ("F0", "<f4", 3),
                           ("emissionColor", "<f4", 3), ("emissionIntensity", "<f4"), ("id", "<i4")])
LIGHT_DTYPE = np.dtype([("position", "<f4", 3), ("color", "<f4", 3), ("intensity", "<f4")])

_fp = C.POINTER(C.c_float)


def _p(a):
    return C.c_void_p(a.ctypes.data)


def _f(a):
    return a.ctypes.data_as(_fp)


def _f32(v, n):
    a = np.ascontiguousarray(np.asarray(v, np.float32).reshape(-1))
    assert a.size == n
    return a


CAMERA_STEP_DTYPE = np.dtype([("dt", np.float32), ("mouse_x", np.float32), ("mouse_y", np.float32), ("keys", np.uint32),
                              ("right", np.uint32)])


class _CpuOracle:
    prefix = ""
    path: Path

    def __init__(self):
        if not self.path.exists():
            raise FileNotFoundError(f"{self.path} has not been built")
        self.lib = C.CDLL(str(self.path))
        g = lambda n: getattr(self.lib, self.prefix + n)  # noqa: E731
        g("pcg_hash").restype = C.c_uint32
        g("pcg_hash").argtypes = [C.c_uint32]
        g("pcg_float").restype = C.c_float
        g("pcg_float").argtypes = [C.POINTER(C.c_uint32)]
        g("hardware_threads").restype = C.c_int
        self._g = g

    def pcg_hash(self, seed: int) -> int:
        return int(self._g("pcg_hash")(seed & 0xFFFFFFFF))

    def pcg_float(self, seed: int):
        s = C.c_uint32(seed & 0xFFFFFFFF)
        v = self._g("pcg_float")(C.byref(s))
        return float(v), int(s.value)

    def hardware_threads(self) -> int:
        return int(self._g("hardware_threads")())

    def primary_hits(self, spheres, origin, dirs, threads=0):
        H, W = dirs.shape[:2]
        out = np.empty((H, W), np.int32)
        s = np.ascontiguousarray(spheres)
        d = np.ascontiguousarray(dirs, np.float32)
        o = _f32(origin, 3)
        fn = self._g("primary_hits")
        fn.restype = None
        fn.argtypes = [C.c_void_p, C.c_uint32, _fp, _fp, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int]
        fn(_p(s), len(s), _f(o), _f(d), W, H, _p(out), threads)
        return out

    def render(self, spheres, materials, lights, origin, dirs, first_frame=1, n_frames=1, stride=1, max_bounces=5,
               sky=False, accum=None, rows=None, threads=0):
        H, W = dirs.shape[:2]
        if accum is None:
            accum = np.zeros((H, W, 4), np.float32)
        y0, y1 = rows if rows is not None else (0, H)
        s, m, l = (np.ascontiguousarray(a) for a in (spheres, materials, lights))
        d = np.ascontiguousarray(dirs, np.float32)
        o = _f32(origin, 3)
        fn = self._g("render")
        fn.restype = C.c_uint64 if self.prefix == "orc_" else None
        fn.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, _fp, _fp, C.c_uint32,
                       C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                       C.c_void_p, C.c_int]
        r = fn(_p(s), len(s), _p(m), len(m), _p(l), len(l), _f(o), _f(d), W, H, y0, y1, first_frame, n_frames, stride,
               max_bounces, int(sky), _p(accum), threads)
        self.last_rays = int(r) if r is not None else None
        return accum

    def pack_rgba8(self, accum, divisor):
        a = np.ascontiguousarray(accum, np.float32)
        n = a.size // 4
        out = np.empty(a.shape[:-1], np.uint32)
        fn = self._g("pack_rgba8")
        fn.restype = None
        fn.argtypes = [_fp, C.c_uint32, C.c_float, C.c_void_p]
        fn(_f(a), n, float(divisor), _p(out))
        return out


class OraclePort(_CpuOracle):
    prefix = "orc_"
    path = PORT_LIB

    def camera(self, pos, direction, fov, near, far, W, H, threads=0):
        ip, iv = np.empty(16, np.float32), np.empty(16, np.float32)
        fn = self.lib.orc_camera_matrices
        fn.restype = None
        fn.argtypes = [_fp, _fp, C.c_float, C.c_float, C.c_float, C.c_uint32, C.c_uint32, _fp, _fp]
        fn(_f(_f32(pos, 3)), _f(_f32(direction, 3)), fov, near, far, W, H, _f(ip), _f(iv))
        rays = np.empty((H, W, 3), np.float32)
        fr = self.lib.orc_ray_directions
        fr.restype = None
        fr.argtypes = [_fp, _fp, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int]
        fr(_f(ip), _f(iv), W, H, _p(rays), threads)
        return rays, ip, iv

    def camera_walk(self, pos, direction, steps):
        """Camera::onUpdate (Camera.cpp:30-108) over a scripted input sequence (CAMERA_STEP_DTYPE): per-step
        positions, directions and moved flags."""
        steps = np.ascontiguousarray(steps, CAMERA_STEP_DTYPE)
        p, d, last = _f32(pos, 3).copy(), _f32(direction, 3).copy(), np.zeros(2, np.float32)
        fn = self.lib.orc_camera_update
        fn.restype = C.c_int
        fn.argtypes = [_fp, _fp, _fp, C.c_uint32, C.c_int, C.c_float, C.c_float, C.c_float]
        op, od, om = np.empty((len(steps), 3), np.float32), np.empty((len(steps), 3), np.float32), np.empty(len(steps), bool)
        for i, st in enumerate(steps):
            om[i] = bool(fn(_f(p), _f(d), _f(last), int(st["keys"]), int(st["right"]), float(st["mouse_x"]), float(st["mouse_y"]), float(st["dt"])))
            op[i], od[i] = p, d
        return op, od, om

    def node_transform(self, parent, position, rot_xyzw, scale):
        out = np.empty(16, np.float32)
        fn = self.lib.orc_node_transform
        fn.restype = None
        fn.argtypes = [_fp, _fp, _fp, _fp, _fp]
        fn(_f(_f32(parent, 16)), _f(_f32(position, 3)), _f(_f32(rot_xyzw, 4)), _f(_f32(scale, 3)), _f(out))
        return out

    def transform_sphere(self, glob, sphere5):
        out = np.empty(5, np.float32)
        fn = self.lib.orc_transform_sphere
        fn.restype = None
        fn.argtypes = [_fp, _fp, _fp]
        fn(_f(_f32(glob, 16)), _f(_f32(sphere5, 5)), _f(out))
        return out

    def flatten_json(self, j):
        """Flatten a parsed scene.json with the port's own transform code (pre-order, Renderer.cu:67-96)."""
        spheres = []

        def rec(node, parent):
            t = node["transformation"]
            g = self.node_transform(parent, t["position"], t["rotation"], t["scale"])
            for s in node.get("spheres", []) or []:
                src = np.array([*s["center"], s["radius"], 0.0], np.float32)
                w = self.transform_sphere(g, src)
                spheres.append((w[0], w[1], w[2], w[3], int(s["materialIndex"])))
            for c in node.get("children", []) or []:
                rec(c, g)

        rec(j["sceneGraph"], np.eye(4, dtype=np.float32).reshape(-1))
        n_mat = len(j.get("materials", []))
        out = np.zeros(len(spheres), SPHERE_DTYPE)
        for i, s in enumerate(spheres):
            out["center"][i] = s[:3]
            out["radius"][i] = s[3]
            out["material"][i] = s[4] if 0 <= s[4] < n_mat else 0  # Renderer.cu:30-37
        return out


class ReferenceCpu(_CpuOracle):
    prefix = "refcpu_"
    path = REFCPU_LIB

    def camera(self, pos, direction, fov, near, far, W, H, threads=0):
        rays = np.empty((H, W, 3), np.float32)
        ip, iv = np.empty(16, np.float32), np.empty(16, np.float32)
        fn = self.lib.refcpu_camera
        fn.restype = C.c_int
        fn.argtypes = [_fp, _fp, C.c_float, C.c_float, C.c_float, C.c_uint32, C.c_uint32, C.c_void_p, _fp, _fp]
        rc = fn(_f(_f32(pos, 3)), _f(_f32(direction, 3)), fov, near, far, W, H, _p(rays), _f(ip), _f(iv))
        if rc != 0:
            raise RuntimeError("reference Camera::Resize built no ray table (1600x900 quirk)")
        return rays, ip, iv

    def camera_walk(self, pos, direction, fov, near, far, W, H, steps):
        """The reference's Camera::onUpdate (Camera.cpp:30-108) driven by scripted input. steps: array of
        CAMERA_STEP_DTYPE (dt, mouse_x, mouse_y, keys [W=1 S=2 A=4 D=8 Q=16 E=32], right). Returns per-step
        (positions, directions, inverse view matrices, moved flags) and the final ray table."""
        steps = np.ascontiguousarray(steps, CAMERA_STEP_DTYPE)
        n = len(steps)
        op, od = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
        oiv, om = np.empty((n, 16), np.float32), np.empty(n, np.int32)
        rays = np.empty((H, W, 3), np.float32)
        fn = self.lib.refcpu_camera_walk
        fn.restype = C.c_int
        fn.argtypes = [_fp, _fp, C.c_float, C.c_float, C.c_float, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int, C.c_void_p,
                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        fn(_f(_f32(pos, 3)), _f(_f32(direction, 3)), fov, near, far, W, H, _p(steps), n, _p(op), _p(od), _p(oiv), _p(om), _p(rays))
        return op, od, oiv, om.astype(bool), rays

    def load_scene(self, path):
        """Utils::importScene + traverseSceneGraph: (spheres, materials, lights, info dict)."""
        self.lib.refcpu_scene_load.restype = C.c_void_p
        self.lib.refcpu_scene_load.argtypes = [C.c_char_p]
        h = C.c_void_p(self.lib.refcpu_scene_load(str(path).encode()))
        nS, nM, nL = C.c_uint32(), C.c_uint32(), C.c_uint32()
        pos, dr = np.empty(3, np.float32), np.empty(3, np.float32)
        fov = C.c_float()
        mb, sky, acc = C.c_int(), C.c_int(), C.c_int()
        self.lib.refcpu_scene_info.restype = None
        self.lib.refcpu_scene_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint32)] * 3 + [_fp, _fp, _fp] + [C.POINTER(C.c_int)] * 3
        self.lib.refcpu_scene_info(h, C.byref(nS), C.byref(nM), C.byref(nL), _f(pos), _f(dr), C.byref(fov), C.byref(mb),
                                   C.byref(sky), C.byref(acc))
        s = np.zeros(nS.value, SPHERE_DTYPE)
        m = np.zeros(nM.value, MATERIAL_DTYPE)
        l = np.zeros(nL.value, LIGHT_DTYPE)
        self.lib.refcpu_scene_arrays.restype = None
        self.lib.refcpu_scene_arrays.argtypes = [C.c_void_p] * 4
        self.lib.refcpu_scene_arrays(h, _p(s), _p(m), _p(l))
        self.lib.refcpu_scene_free.argtypes = [C.c_void_p]
        self.lib.refcpu_scene_free(h)
        info = dict(position=pos, direction=dr, fov=float(fov.value), maxBounces=mb.value, skyLight=bool(sky.value),
                    accumulation=bool(acc.value))
        return s, m, l, info

    def reexport_scene(self, src, dst):
        """Utils::importScene(src) -> Utils::exportScene(dst)."""
        self.lib.refcpu_scene_load.restype = C.c_void_p
        self.lib.refcpu_scene_load.argtypes = [C.c_char_p]
        h = C.c_void_p(self.lib.refcpu_scene_load(str(src).encode()))
        self.lib.refcpu_scene_export.argtypes = [C.c_void_p, C.c_char_p]
        self.lib.refcpu_scene_export(h, str(dst).encode())
        self.lib.refcpu_scene_free.argtypes = [C.c_void_p]
        self.lib.refcpu_scene_free(h)


def have_reference_cpu() -> bool:
    return REFCPU_LIB.exists()


def have_ref_headless() -> bool:
    return REF_HEADLESS.exists()


REF_HEADLESS_SHIM = HERE / "_ref" / "ref_headless_shim"


def have_dropin_shim() -> bool:
    return REF_HEADLESS_SHIM.exists()


def run_dropin_shim(scene_json, W, H, bounces, sky, frames, dump_at=(), timeout=600):
    """Run oracle/_ref/ref_headless_shim: the reference application's protocol and its own Camera/SceneNode/Utils
    objects, with examples/dropin/Renderer.cpp (the product behind the reference's unmodified Renderer.h) in place
    of Renderer.cu. Same arguments and dump files as run_ref_headless (no rays/hit dumps)."""
    return run_ref_headless(scene_json, W, H, bounces, sky, frames, dump_at, timeout, exe=REF_HEADLESS_SHIM, extras=False)


def run_ref_headless(scene_json, W, H, bounces, sky, frames, dump_at=(), timeout=600, exe=None, extras=True):
    """Run the reference's CUDA renderer (needs a GPU). Returns (info, {name: ndarray})."""
    with tempfile.TemporaryDirectory() as td:
        prefix = str(Path(td) / "ref") if dump_at else "-"
        cmd = [str(exe or REF_HEADLESS), str(scene_json), str(W), str(H), str(bounces), str(int(sky)), str(frames), prefix]
        if dump_at:
            cmd.append(",".join(str(k) for k in dump_at))
        proc = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        if proc.returncode != 0:
            raise RuntimeError(f"ref_headless failed ({proc.returncode}): {proc.stdout}\n{proc.stderr}")
        info = json.loads(proc.stdout.strip().splitlines()[-1])
        out = {}
        if dump_at:
            if extras:
                out["rays"] = np.fromfile(prefix + ".rays.f32", np.float32).reshape(H, W, 3)
                out["hit"] = np.fromfile(prefix + ".hit.i32", np.int32).reshape(H, W)
                out["spheres"] = np.fromfile(prefix + ".spheres.f32", np.float32).reshape(-1, 5)
            for k in dump_at:
                out[f"acc{k}"] = np.fromfile(prefix + f".acc{k}.f32", np.float32).reshape(H, W, 4)
                out[f"rgba{k}"] = np.fromfile(prefix + f".rgba{k}.u32", np.uint32).reshape(H, W)
        return info, out
