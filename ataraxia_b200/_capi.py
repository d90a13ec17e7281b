"""ctypes binding of the C-ABI (include/ataraxia_b200.h) — the same symbols a C++ host links.

The shared library is built in-tree by ataraxia_b200/build.py (nvcc, sm_100a). If it is
missing this module raises: there is no Python or CPU fallback for the render path.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libataraxia_b200.so"
if os.environ.get("ATX_LIB"):  # development: an experimental build of the same library (tools/build_variant.sh)
    LIB_PATH = Path(os.environ["ATX_LIB"]).resolve()

# every symbol include/ataraxia_b200.h declares (tests check the .so exports exactly these)
SYMBOLS = [
    "atx_create", "atx_destroy", "atx_last_error", "atx_version",
    "atx_resize", "atx_upload_scene", "atx_set_camera", "atx_set_camera_matrices", "atx_set_settings",
    "atx_set_tuning", "atx_reset", "atx_frame_index",
    "atx_render", "atx_render_frames", "atx_calibrate", "atx_sync", "atx_last_render_ms", "atx_last_mega_kind", "atx_event_record", "atx_event_elapsed_ms",
    "atx_read_accum", "atx_write_accum", "atx_read_rgba8", "atx_read_hit_ids", "atx_read_ray_directions",
    "atx_get_counters", "atx_reset_counters", "atx_accum_device_ptr", "atx_stream", "atx_host_alloc", "atx_host_free",
    "atx_comm_unique_id", "atx_comm_init_rank", "atx_comm_destroy", "atx_allreduce_accum", "atx_allreduce_preview", "atx_read_preview",
    "atx_host_camera_matrices", "atx_host_ray_directions", "atx_host_node_transform", "atx_host_transform_sphere", "atx_host_mat4_mul",
    "atx_host_camera_update",
    "atx_save_checkpoint", "atx_load_checkpoint", "atx_scene_sha256", "atx_host_sha256", "atx_last_reduce_kind", "atx_render_tiles", "atx_render_tile_share",
]

ATX_OK = 0
ATX_ERR_INVALID, ATX_ERR_CUDA, ATX_ERR_NCCL, ATX_ERR_NO_DEVICE, ATX_ERR_ALLOC = -1, -2, -3, -4, -5
VARIANT_AUTO, VARIANT_MEGAKERNEL, VARIANT_WAVEFRONT = 0, 1, 2
TUNE_CHUNK_SPHERES, TUNE_MEGA_KIND, TUNE_PARK_THRESHOLD, TUNE_CLAIM_THRESHOLD, TUNE_REDUCE = 1, 2, 3, 4, 5
REDUCE_NONE, REDUCE_PEER_MEMORY, REDUCE_NCCL = 0, 1, 2
MEGA_AUTO, MEGA_WHILE_WHILE, MEGA_PAIR, MEGA_WARP_QUEUE, MEGA_PAIR_LOCKSTEP = 0, 1, 2, 3, 4

# numpy views of the reference PODs (SceneNode.h:11-21, Scene.h:17-47)
SPHERE_DTYPE = np.dtype([("center", "<f4", 3), ("radius", "<f4"), ("material", "<i4")])
MATERIAL_DTYPE = np.dtype([("albedo", "<f4", 3), ("roughness", "<f4"), ("metallic", "<f4"), ("F0", "<f4", 3),
                           ("emissionColor", "<f4", 3), ("emissionIntensity", "<f4"), ("id", "<i4")])
LIGHT_DTYPE = np.dtype([("position", "<f4", 3), ("color", "<f4", 3), ("intensity", "<f4")])
assert SPHERE_DTYPE.itemsize == 20 and MATERIAL_DTYPE.itemsize == 52 and LIGHT_DTYPE.itemsize == 28


class Counters(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("rays", C.c_uint64), ("sphere_tests", C.c_uint64), ("launches", C.c_uint64),
                ("rays_traced", C.c_uint64), ("sphere_tests_executed", C.c_uint64)]


class AtxError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"ataraxia_b200 error {status}: {message}")
        self.status = status


KEY_W, KEY_S, KEY_A, KEY_D, KEY_Q, KEY_E = 1, 2, 4, 8, 16, 32


class CameraInput(C.Structure):
    """atx_camera_input (include/ataraxia_b200.h): what Camera::onUpdate reads from the window."""
    _fields_ = [("keys", C.c_uint32), ("right_button", C.c_uint32), ("mouse_x", C.c_float), ("mouse_y", C.c_float)]


_lib = None


def lib() -> C.CDLL:
    """Load the CUDA extension; fail loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m ataraxia_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback for the render path.")
    l = C.CDLL(str(LIB_PATH))
    fp = C.POINTER(C.c_float)
    vp = C.c_void_p
    l.atx_last_error.restype = C.c_char_p
    l.atx_version.restype = C.c_char_p
    sig = {
        "atx_create": [C.c_int, C.POINTER(vp)],
        "atx_destroy": [vp],
        "atx_resize": [vp, C.c_uint32, C.c_uint32],
        "atx_upload_scene": [vp, vp, C.c_size_t, vp, C.c_size_t, vp, C.c_size_t],
        "atx_set_camera": [vp, fp, fp, C.c_float, C.c_float, C.c_float],
        "atx_set_camera_matrices": [vp, fp, fp, fp],
        "atx_set_settings": [vp, C.c_int, C.c_int, C.c_int],
        "atx_set_tuning": [vp, C.c_int, C.c_int64],
        "atx_reset": [vp],
        "atx_frame_index": [vp, C.POINTER(C.c_uint32)],
        "atx_render": [vp, C.c_uint32, C.c_int],
        "atx_render_frames": [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int],
        "atx_calibrate": [vp, C.c_uint32, fp, fp],
        "atx_sync": [vp],
        "atx_last_render_ms": [vp, fp],
        "atx_event_record": [vp, C.c_int],
        "atx_event_elapsed_ms": [vp, C.c_int, C.c_int, fp],
        "atx_read_accum": [vp, vp],
        "atx_write_accum": [vp, vp, C.c_uint32],
        "atx_read_rgba8": [vp, vp, C.c_uint32],
        "atx_read_hit_ids": [vp, vp],
        "atx_read_ray_directions": [vp, vp],
        "atx_get_counters": [vp, C.POINTER(Counters)],
        "atx_reset_counters": [vp],
        "atx_accum_device_ptr": [vp, C.POINTER(vp)],
        "atx_stream": [vp, C.POINTER(vp)],
        "atx_comm_unique_id": [C.POINTER(C.c_uint8)],
        "atx_comm_init_rank": [vp, C.c_int, C.c_int, C.POINTER(C.c_uint8)],
        "atx_comm_destroy": [vp],
        "atx_allreduce_accum": [vp],
        "atx_allreduce_preview": [vp],
        "atx_read_preview": [vp, vp, vp, C.c_uint32],
        "atx_last_mega_kind": [vp, C.POINTER(C.c_int)],
        "atx_host_alloc": [C.c_size_t, C.POINTER(vp)],
        "atx_host_free": [vp],
        "atx_host_camera_matrices": [fp, fp, C.c_float, C.c_float, C.c_float, C.c_uint32, C.c_uint32, fp, fp, fp, fp],
        "atx_host_ray_directions": [fp, fp, C.c_uint32, C.c_uint32, vp],
        "atx_host_node_transform": [fp, fp, fp, fp, fp, fp],
        "atx_host_transform_sphere": [fp, vp, vp],
        "atx_host_mat4_mul": [fp, fp, fp],
        "atx_host_camera_update": [fp, fp, fp, C.POINTER(CameraInput), C.c_float, C.POINTER(C.c_int)],
        "atx_save_checkpoint": [vp, C.c_char_p, C.c_uint32, C.c_uint32],
        "atx_load_checkpoint": [vp, C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)],
        "atx_scene_sha256": [vp, C.POINTER(C.c_uint8)],
        "atx_host_sha256": [vp, C.c_size_t, C.POINTER(C.c_uint8)],
        "atx_last_reduce_kind": [vp, C.POINTER(C.c_int)],
        "atx_render_tiles": [vp, C.c_uint32, C.c_uint32, C.c_int, C.c_int],
        "atx_render_tile_share": [vp, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_uint32, C.c_uint32],
    }
    for name, argtypes in sig.items():
        fn = getattr(l, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    _lib = l
    return l


def check(status: int) -> None:
    if status != ATX_OK:
        raise AtxError(status, lib().atx_last_error().decode("utf-8", "replace"))


def f32(values, n=None) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(values, dtype=np.float32).reshape(-1))
    if n is not None and a.size != n:
        raise ValueError(f"expected {n} floats, got {a.size}")
    return a


def fptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def vptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)
