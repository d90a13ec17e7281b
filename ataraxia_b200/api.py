"""Host-side mirror of the reference's Engine API for the hot path, over the C-ABI.

Same class names, method names, argument meaning and quirks as
/root/reference/Engine/include/{Scene,SceneNode,Camera,Renderer}.h, so the parity tests read
like reference client code (Engine/src/main.cpp:211-220):

    renderer.onResize(w, h); camera.Resize(w, h); renderer.Render(camera, scene)

All arithmetic that parity depends on (matrices, flattening) runs inside the native library
(atx_host_* helpers); all rendering runs in the CUDA kernels. Nothing here computes pixels.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _capi
from ._capi import LIGHT_DTYPE, MATERIAL_DTYPE, SPHERE_DTYPE, check, f32, fptr, vptr

IDENTITY = np.eye(4, dtype=np.float32).reshape(-1)  # column-major == row-major for I


def _vec3(v) -> np.ndarray:
    return np.asarray(v, dtype=np.float32).reshape(3).copy()


@dataclass
class Sphere:
    """SceneNode.h:11-21."""
    center: Sequence[float] = (0.0, 0.0, 0.0)
    radius: float = 0.0
    id: int = 0  # material index


@dataclass
class Material:
    """Scene.h:28-47 (defaults included)."""
    albedo: Sequence[float] = (1.0, 1.0, 1.0)
    roughness: float = 0.0
    metallic: float = 0.0
    F0: Sequence[float] = (0.04, 0.04, 0.04)
    emissionColor: Sequence[float] = (0.0, 0.0, 0.0)
    emissionIntensity: float = 0.0
    id: int = 0

    def getEmission(self):
        return _vec3(self.emissionColor) * np.float32(self.emissionIntensity)


@dataclass
class Light:
    """Scene.h:17-26."""
    position: Sequence[float] = (0.0, 0.0, 0.0)
    color: Sequence[float] = (0.0, 0.0, 0.0)
    intensity: float = 0.0


@dataclass
class Settings:
    """Scene.h:49-56."""
    accumulation: bool = True
    skyLight: bool = False
    maxBounces: int = 15


class SceneNode:
    """SceneNode.h:23-66 / SceneNode.cpp."""

    def __init__(self, name: str = "Untitled"):
        self.m_name = name
        self.m_position = np.zeros(3, np.float32)
        self.m_rotation = np.zeros(4, np.float32)  # (x, y, z, w); glm::quat() value-initialises to zeros
        self.m_scale = np.ones(3, np.float32)
        self.m_localTransform = IDENTITY.copy()
        self.m_globalTransform = IDENTITY.copy()
        self.m_children: List["SceneNode"] = []
        self.m_spheres: List[Sphere] = []
        self.m_transformDirty = True

    def setPosition(self, p): self.m_position = _vec3(p); self.m_transformDirty = True
    def setRotation(self, q_xyzw): self.m_rotation = np.asarray(q_xyzw, np.float32).reshape(4).copy(); self.m_transformDirty = True
    def setScale(self, s): self.m_scale = _vec3(s); self.m_transformDirty = True
    def getPosition(self): return self.m_position
    def getRotation(self): return self.m_rotation
    def getScale(self): return self.m_scale
    def addChild(self, child: "SceneNode"): self.m_children.append(child)

    def removeChild(self, child: "SceneNode"):
        self.m_children = [c for c in self.m_children if c is not child]

    def getChildren(self): return self.m_children
    def addSphere(self, sphere: Sphere): self.m_spheres.append(sphere)

    def removeSphere(self, index: int):
        if 0 <= index < len(self.m_spheres):
            del self.m_spheres[index]

    def getSpheres(self): return self.m_spheres
    def getName(self): return self.m_name
    def setName(self, name): self.m_name = name
    def getGlobalTransform(self): return self.m_globalTransform

    def updateGlobalTransform(self, parentTransform=None):
        """SceneNode.cpp:42-59."""
        parent = IDENTITY if parentTransform is None else f32(parentTransform, 16)
        local = np.empty(16, np.float32)
        glob = np.empty(16, np.float32)
        if self.m_transformDirty:
            check(_capi.lib().atx_host_node_transform(fptr(parent), fptr(self.m_position), fptr(self.m_rotation),
                                                      fptr(self.m_scale), fptr(local), fptr(glob)))
            self.m_localTransform = local
            self.m_transformDirty = False
        else:
            glob = _mat_mul(parent, self.m_localTransform)
        self.m_globalTransform = glob
        for child in self.m_children:
            child.updateGlobalTransform(self.m_globalTransform)


def _mat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """a * b in glm's mul4x4 order, through the native helper."""
    out = np.empty(16, np.float32)
    check(_capi.lib().atx_host_mat4_mul(fptr(f32(a, 16)), fptr(f32(b, 16)), fptr(out)))
    return out


@dataclass
class InputState:
    """What Camera::onUpdate asks the window for (Core/include/input/Input.h:11-16), as plain data so that
    camera motion can be scripted: keys is a string of the pressed keys among "WASDQE", mouse the cursor
    position, right_button whether MouseButton::Right is held (the reference only moves the camera then)."""
    keys: str = ""
    mouse: Sequence[float] = (0.0, 0.0)
    right_button: bool = False

    def pack(self) -> _capi.CameraInput:
        bits = {"W": _capi.KEY_W, "S": _capi.KEY_S, "A": _capi.KEY_A, "D": _capi.KEY_D, "Q": _capi.KEY_Q, "E": _capi.KEY_E}
        mask = 0
        for k in self.keys.upper():
            mask |= bits[k]
        return _capi.CameraInput(mask, int(bool(self.right_button)), float(self.mouse[0]), float(self.mouse[1]))


class Camera:
    """Camera.h:14-88 / Camera.cpp. onUpdate takes the input as data (InputState) instead of polling GLFW."""

    def __init__(self, fov: Optional[float] = None, nearClip: float = 0.1, farClip: float = 100.0,
                 position=None, direction=None):
        self.m_projectionMatrix = IDENTITY.copy()
        self.m_viewMatrix = IDENTITY.copy()
        self.m_inverseProjectionMatrix = IDENTITY.copy()
        self.m_inverseViewMatrix = IDENTITY.copy()
        self.m_position = np.zeros(3, np.float32)
        self.m_direction = np.zeros(3, np.float32)
        self.m_rayDirection: Optional[np.ndarray] = None
        self.m_fov, self.m_nearClip, self.m_farClip = 45.0, 0.1, 100.0
        self.m_width = self.m_height = 0
        self.m_viewDirty = self.m_projectionDirty = True
        self.m_lastMousePos = np.zeros(2, np.float32)  # Camera.h:83
        if fov is not None:
            # Camera.cpp:14-29: both value constructors preset 1600x900 and build both matrices
            self.m_fov, self.m_nearClip, self.m_farClip = float(fov), float(nearClip), float(farClip)
            self.m_width, self.m_height = 1600, 900
            if position is None:
                self.m_direction = _vec3((0.0, 0.0, -1.0))
                self.m_position = _vec3((0.0, 0.0, 3.0))
            else:
                self.m_position = _vec3(position)
                self.m_direction = _vec3(direction)
            self._updateViewMatrix()
            self._updateProjectionMatrix()

    # setters only mark dirty (Camera.h:56-58)
    def setPosition(self, p): self.m_position = _vec3(p); self.m_viewDirty = True
    def setDirection(self, d): self.m_direction = _vec3(d); self.m_viewDirty = True
    def setFov(self, fov): self.m_fov = float(fov); self.m_projectionDirty = True
    def getPosition(self): return self.m_position
    def getDirection(self): return self.m_direction
    def getFov(self): return self.m_fov
    def getViewMatrix(self): return self.m_viewMatrix
    def getProjectionMatrix(self): return self.m_projectionMatrix
    def getInverseViewMatrix(self): return self.m_inverseViewMatrix
    def getInverseProjectionMatrix(self): return self.m_inverseProjectionMatrix
    @staticmethod
    def getRotationSpeed(): return 0.3

    def onUpdate(self, ts: float, inp: Optional[InputState] = None) -> bool:
        """Camera::onUpdate (Camera.cpp:30-108): W/S, A/D, Q/E at speed 5 and mouse look at rotation speed 0.3
        while the right button is held; returns whether the camera moved. The arithmetic runs in the library
        (atx_host_camera_update) in glm's evaluation order."""
        inp = inp or InputState()
        moved = C.c_int(0)
        packed = inp.pack()
        check(_capi.lib().atx_host_camera_update(fptr(self.m_position), fptr(self.m_direction), fptr(self.m_lastMousePos),
                                                 C.byref(packed), float(ts), C.byref(moved)))
        if moved.value:
            self.m_viewDirty = True
        if self.m_viewDirty and inp.right_button:   # :101-105 (not reached when the button is up, :36-40)
            self._updateViewMatrix()
            self.m_rayDirection = None
        return bool(moved.value)

    def copy(self) -> "Camera":
        """Value copy, like `m_scene.camera = m_camera` (Camera.h:20-38)."""
        c = Camera()
        for k, v in self.__dict__.items():
            setattr(c, k, v.copy() if isinstance(v, np.ndarray) else v)
        c.m_rayDirection = None
        return c

    def Resize(self, width: int, height: int):
        """Camera.cpp:110-127 (including the early return that leaves a 1600x900 camera without rays)."""
        if width == 0 or height == 0:
            print("Error: Width or height cannot be zero.")
            return
        if width == self.m_width and height == self.m_height:
            return
        self.m_width, self.m_height = int(width), int(height)
        self.m_projectionDirty = True
        self._updateProjectionMatrix()
        self.m_rayDirection = None  # rebuilt lazily; the renderer generates rays in-kernel

    def getRayDirection(self) -> np.ndarray:
        """Camera.h:60 — the host ray table (Camera.cpp:161-195), computed on demand."""
        if self.m_rayDirection is None:
            out = np.empty((self.m_height, self.m_width, 3), np.float32)
            check(_capi.lib().atx_host_ray_directions(fptr(self.m_inverseProjectionMatrix), fptr(self.m_inverseViewMatrix),
                                                      self.m_width, self.m_height, vptr(out)))
            self.m_rayDirection = out
        return self.m_rayDirection

    def _matrices(self):
        proj, view, iproj, iview = (np.empty(16, np.float32) for _ in range(4))
        w, h = max(self.m_width, 1), max(self.m_height, 1)
        check(_capi.lib().atx_host_camera_matrices(fptr(self.m_position), fptr(self.m_direction), self.m_fov,
                                                   self.m_nearClip, self.m_farClip, w, h,
                                                   fptr(proj), fptr(view), fptr(iproj), fptr(iview)))
        return proj, view, iproj, iview

    def _updateProjectionMatrix(self):
        if self.m_projectionDirty:
            proj, _, iproj, _ = self._matrices()
            self.m_projectionMatrix, self.m_inverseProjectionMatrix = proj, iproj
            self.m_projectionDirty = False

    def _updateViewMatrix(self):
        if self.m_viewDirty:
            _, view, _, iview = self._matrices()
            self.m_viewMatrix, self.m_inverseViewMatrix = view, iview
            self.m_viewDirty = False


@dataclass
class Scene:
    """Scene.h:58-80."""
    rootNode: SceneNode = field(default_factory=lambda: SceneNode("Scene"))
    materials: List[Material] = field(default_factory=list)
    lights: List[Light] = field(default_factory=list)
    settings: Settings = field(default_factory=Settings)
    camera: Camera = field(default_factory=Camera)


class _PinnedPixels:
    """Owner of one atx_host_alloc block, exposed to numpy through the array interface: arrays made from it hold
    a reference, and the block is released when the last of them is gone."""

    def __init__(self, ptr: C.c_void_p, shape):
        self._ptr = ptr
        self.__array_interface__ = {"data": (ptr.value, False), "shape": tuple(shape), "typestr": "<u4", "version": 3}

    def __del__(self):
        try:
            _capi.lib().atx_host_free(self._ptr)
        except Exception:
            pass


class Image:
    """Headless stand-in for Core/include/Image.h: what Renderer touches (ctor, setData, getWidth/getHeight)."""

    def __init__(self, width: int, height: int, pinned: bool = False):
        self.m_width, self.m_height = width, height
        self._pinned = False
        if pinned and width * height > 0:
            # the renderer's image: page-locked, so the per-frame read-back is a single DMA
            ptr = C.c_void_p()
            if _capi.lib().atx_host_alloc(width * height * 4, C.byref(ptr)) == _capi.ATX_OK:
                self.data = np.asarray(_PinnedPixels(ptr, (height, width)))   # the array (and its views) keep the block alive
                self.data[...] = 0
                self._pinned = True
                return
        self.data = np.zeros((height, width), np.uint32)

    def getWidth(self): return self.m_width
    def getHeight(self): return self.m_height
    def setData(self, data):
        if self._pinned:
            self.data[...] = np.asarray(data, np.uint32).reshape(self.m_height, self.m_width)
        else:
            self.data = data

    def savePPM(self, path: str) -> None:
        """Output sink in place of the Vulkan texture upload (Core/src/Image.cpp:183-271): binary PPM, rows top to
        bottom, i.e. with the V flip the UI applies when it shows the texture (main.cpp:185-187)."""
        px = np.ascontiguousarray(self.data[::-1])
        rgb = np.stack([px & 0xFF, (px >> 8) & 0xFF, (px >> 16) & 0xFF], axis=-1).astype(np.uint8)
        with open(path, "wb") as f:
            f.write(f"P6\n{self.m_width} {self.m_height}\n255\n".encode())
            f.write(rgb.tobytes())

    def savePNG(self, path: str) -> None:
        """The same image (RGBA8 as packed by Renderer.h:70-78, V flipped like the UI) as a PNG: 8-bit RGBA, filter
        0 on every row, zlib stream from the standard library."""
        import struct
        import zlib
        px = np.ascontiguousarray(self.data[::-1]).astype("<u4")
        rows = px.view(np.uint8).reshape(self.m_height, self.m_width * 4)   # bytes r, g, b, a
        raw = np.concatenate([np.zeros((self.m_height, 1), np.uint8), rows], axis=1).tobytes()

        def chunk(kind: bytes, body: bytes) -> bytes:
            return struct.pack(">I", len(body)) + kind + body + struct.pack(">I", zlib.crc32(kind + body) & 0xFFFFFFFF)

        with open(path, "wb") as f:
            f.write(b"\x89PNG\r\n\x1a\n")
            f.write(chunk(b"IHDR", struct.pack(">IIBBBBB", self.m_width, self.m_height, 8, 6, 0, 0, 0)))
            f.write(chunk(b"IDAT", zlib.compress(raw, 6)))
            f.write(chunk(b"IEND", b""))


def traverseSceneGraph(node: Optional[SceneNode], parentTransform=None) -> List[Sphere]:
    """Renderer::traverseSceneGraph (Renderer.cu:67-96): pre-order, world-space centres, mean-scale radii."""
    out: List[Sphere] = []

    def rec(n: Optional[SceneNode], parent):
        if n is None:
            return
        n.updateGlobalTransform(parent)
        g = n.getGlobalTransform()
        for s in n.getSpheres():
            src = np.zeros(1, SPHERE_DTYPE)
            src["center"][0] = _vec3(s.center)
            src["radius"][0] = s.radius
            src["material"][0] = s.id
            dst = np.zeros(1, SPHERE_DTYPE)
            check(_capi.lib().atx_host_transform_sphere(fptr(g), vptr(src), vptr(dst)))
            out.append(Sphere(tuple(float(v) for v in dst["center"][0]), float(dst["radius"][0]), int(dst["material"][0])))
        for c in n.getChildren():
            rec(c, g)

    rec(node, IDENTITY if parentTransform is None else parentTransform)
    return out


def pack_spheres(spheres: Sequence[Sphere]) -> np.ndarray:
    a = np.zeros(len(spheres), SPHERE_DTYPE)
    for i, s in enumerate(spheres):
        a["center"][i] = _vec3(s.center); a["radius"][i] = s.radius; a["material"][i] = s.id
    return a


def pack_materials(materials: Sequence[Material]) -> np.ndarray:
    a = np.zeros(len(materials), MATERIAL_DTYPE)
    for i, m in enumerate(materials):
        a["albedo"][i] = _vec3(m.albedo); a["roughness"][i] = m.roughness; a["metallic"][i] = m.metallic
        a["F0"][i] = _vec3(m.F0); a["emissionColor"][i] = _vec3(m.emissionColor)
        a["emissionIntensity"][i] = m.emissionIntensity; a["id"][i] = m.id
    return a


def pack_lights(lights: Sequence[Light]) -> np.ndarray:
    a = np.zeros(len(lights), LIGHT_DTYPE)
    for i, l in enumerate(lights):
        a["position"][i] = _vec3(l.position); a["color"][i] = _vec3(l.color); a["intensity"][i] = l.intensity
    return a


class Renderer:
    """Renderer.h:19-31 over the C-ABI, plus the headless additions (accumulation read-back,
    multi-frame launches, counters, spp-split reduce)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        check(_capi.lib().atx_create(device, C.byref(self._h)))
        self.m_settings = Settings()
        self.m_scene = None
        self.m_image: Optional[Image] = None
        self.m_width = self.m_height = 0
        self.variant = _capi.VARIANT_AUTO
        check(_capi.lib().atx_set_settings(self._h, 1, 0, self.m_settings.maxBounces))

    def close(self):
        if self._h:
            _capi.lib().atx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- reference API ----------------------------------------------------
    def onResize(self, width: int, height: int):
        if self.m_image is not None and self.m_image.getWidth() == width and self.m_image.getHeight() == height:
            return
        check(_capi.lib().atx_resize(self._h, width, height))
        self.m_image = Image(width, height, pinned=True)
        self.m_width, self.m_height = width, height

    def getImage(self): return self.m_image
    def getSettings(self): return self.m_settings

    def setSettings(self, settings: Settings):
        self.m_settings = settings
        check(_capi.lib().atx_set_settings(self._h, int(settings.accumulation), int(settings.skyLight), int(settings.maxBounces)))

    def resetFrameIndex(self):
        check(_capi.lib().atx_reset(self._h))

    def frameIndex(self) -> int:
        v = C.c_uint32()
        check(_capi.lib().atx_frame_index(self._h, C.byref(v)))
        return v.value

    def Render(self, camera: Camera, scene: Scene, frames: int = 1, readback: bool = True):
        """Renderer::Render (Renderer.cu:173-249); `frames` > 1 renders that many frames in one launch."""
        if self.m_scene is not scene or self.frameIndex() == 1:
            self.m_scene = scene
            self.uploadScene(scene)  # allocateDeviceMemory, Renderer.cu:175-179
        if self.m_image is None:
            return
        self.setCamera(camera)
        check(_capi.lib().atx_set_settings(self._h, int(self.m_settings.accumulation), int(self.m_settings.skyLight),
                                           int(self.m_settings.maxBounces)))
        check(_capi.lib().atx_render(self._h, frames, self.variant))
        if readback:
            check(_capi.lib().atx_read_rgba8(self._h, vptr(self.m_image.data), 0))

    # ---- headless additions -------------------------------------------------
    def uploadScene(self, scene: Scene):
        self.uploadArrays(pack_spheres(traverseSceneGraph(scene.rootNode)), pack_materials(scene.materials),
                          pack_lights(scene.lights))

    def uploadArrays(self, spheres: np.ndarray, materials: np.ndarray, lights: np.ndarray):
        assert spheres.dtype == SPHERE_DTYPE and materials.dtype == MATERIAL_DTYPE and lights.dtype == LIGHT_DTYPE
        self._keep = (np.ascontiguousarray(spheres), np.ascontiguousarray(materials), np.ascontiguousarray(lights))
        s, m, l = self._keep
        check(_capi.lib().atx_upload_scene(self._h, vptr(s), len(s), vptr(m), len(m), vptr(l), len(l)))

    def setCamera(self, camera: Camera):
        check(_capi.lib().atx_set_camera_matrices(self._h, fptr(camera.m_position), fptr(camera.m_inverseProjectionMatrix),
                                                  fptr(camera.m_inverseViewMatrix)))

    def renderFrames(self, first: int, count: int, stride: int = 1, zero_first: bool = False):
        check(_capi.lib().atx_render_frames(self._h, first, count, stride, int(zero_first), self.variant))

    def calibrate(self, frames: int = 2):
        """Time the megakernel and the wavefront variant on the current scene (scratch buffer) and make
        VARIANT_AUTO the faster one. Returns (megakernel_ms, wavefront_ms)."""
        a, b = C.c_float(), C.c_float()
        check(_capi.lib().atx_calibrate(self._h, frames, C.byref(a), C.byref(b)))
        return a.value, b.value

    def sync(self): check(_capi.lib().atx_sync(self._h))

    def lastRenderMs(self) -> float:
        v = C.c_float()
        check(_capi.lib().atx_last_render_ms(self._h, C.byref(v)))
        return v.value

    def lastMegaKind(self) -> int:
        """Megakernel form of the last launch (MEGA_WHILE_WHILE / MEGA_PAIR / MEGA_WARP_QUEUE)."""
        v = C.c_int()
        check(_capi.lib().atx_last_mega_kind(self._h, C.byref(v)))
        return v.value

    def eventRecord(self, slot: int): check(_capi.lib().atx_event_record(self._h, slot))

    def eventElapsedMs(self, begin: int, end: int) -> float:
        v = C.c_float()
        check(_capi.lib().atx_event_elapsed_ms(self._h, begin, end, C.byref(v)))
        return v.value

    def setTuning(self, key: int, value: int):
        check(_capi.lib().atx_set_tuning(self._h, key, value))

    def getAccumulation(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.m_height, self.m_width, 4), np.float32)
        check(_capi.lib().atx_read_accum(self._h, vptr(out)))
        return out

    def saveAccumulationPFM(self, path: str) -> None:
        """Float radiance (accumulation / samples, unclamped) as a little-endian PFM, bottom row first as PFM defines
        it — which is the buffer's own row order (row 0 is the bottom of the image, main.cpp:186-187)."""
        acc = self.getAccumulation()
        n = np.where(acc[..., 3:4] > 0, acc[..., 3:4], np.float32(1.0))
        rad = (acc[..., :3] / n).astype("<f4")
        with open(path, "wb") as f:
            f.write(f"PF\n{self.m_width} {self.m_height}\n-1.0\n".encode())
            f.write(np.ascontiguousarray(rad).tobytes())

    def setAccumulation(self, acc: np.ndarray, next_frame_index: int):
        a = np.ascontiguousarray(acc, np.float32)
        check(_capi.lib().atx_write_accum(self._h, vptr(a), next_frame_index))

    def saveCheckpoint(self, path: str, next_frame_index: int = 0, frame_stride: int = 0) -> None:
        """Accumulation buffer + next frame index on disk, bound to the scene and camera by a hash (atx_save_checkpoint)."""
        check(_capi.lib().atx_save_checkpoint(self._h, str(path).encode(), int(next_frame_index), int(frame_stride)))

    def loadCheckpoint(self, path: str):
        """Resume: (next_frame_index, frame_stride). Refuses a file rendered with another size, scene, camera or settings."""
        a, b = C.c_uint32(), C.c_uint32()
        check(_capi.lib().atx_load_checkpoint(self._h, str(path).encode(), C.byref(a), C.byref(b)))
        return a.value, b.value

    def sceneSha256(self) -> bytes:
        buf = (C.c_uint8 * 32)()
        check(_capi.lib().atx_scene_sha256(self._h, buf))
        return bytes(buf)

    def getRGBA8(self, divisor: int = 0, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.m_height, self.m_width), np.uint32)
        check(_capi.lib().atx_read_rgba8(self._h, vptr(out), divisor))
        return out

    def getHitIds(self) -> np.ndarray:
        out = np.empty((self.m_height, self.m_width), np.int32)
        check(_capi.lib().atx_read_hit_ids(self._h, vptr(out)))
        return out

    def getRayDirections(self) -> np.ndarray:
        out = np.empty((self.m_height, self.m_width, 3), np.float32)
        check(_capi.lib().atx_read_ray_directions(self._h, vptr(out)))
        return out

    def counters(self) -> _capi.Counters:
        c = _capi.Counters()
        check(_capi.lib().atx_get_counters(self._h, C.byref(c)))
        return c

    def resetCounters(self): check(_capi.lib().atx_reset_counters(self._h))

    def accumDevicePtr(self) -> int:
        p = C.c_void_p()
        check(_capi.lib().atx_accum_device_ptr(self._h, C.byref(p)))
        return p.value or 0

    # multi-GPU
    @staticmethod
    def commUniqueId() -> bytes:
        buf = (C.c_uint8 * 128)()
        check(_capi.lib().atx_comm_unique_id(buf))
        return bytes(buf)

    def commInitRank(self, n_ranks: int, rank: int, uid: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        check(_capi.lib().atx_comm_init_rank(self._h, n_ranks, rank, buf))

    def commDestroy(self): check(_capi.lib().atx_comm_destroy(self._h))

    def allreducePreview(self):
        """Sum of all ranks' accumulation buffers into the preview buffer; the ranks' own sums stay as they are."""
        check(_capi.lib().atx_allreduce_preview(self._h))

    def getPreview(self, divisor: int):
        """(float4 sums, RGBA8 image resolved with `divisor` = total samples per pixel) of the last preview."""
        acc = np.empty((self.m_height, self.m_width, 4), np.float32)
        rgba = np.empty((self.m_height, self.m_width), np.uint32)
        check(_capi.lib().atx_read_preview(self._h, vptr(acc), vptr(rgba), int(divisor)))
        return acc, rgba
    def allreduceAccum(self): check(_capi.lib().atx_allreduce_accum(self._h))

    def renderTiles(self, first: int, count: int, zero_first: bool = False):
        """Image-tile split across the communicator's ranks (atx_render_tiles): this rank renders frames first..first+count-1
        of its interleaved 8x4 tiles and stores them into every rank's image over NVLink. Collective."""
        check(_capi.lib().atx_render_tiles(self._h, first, count, int(zero_first), self.variant))

    def renderTileShare(self, first: int, count: int, n_shares: int, share: int, zero_first: bool = False):
        """One share of the image-tile split, locally (atx_render_tile_share)."""
        check(_capi.lib().atx_render_tile_share(self._h, first, count, int(zero_first), self.variant, n_shares, share))

    def lastReduceKind(self) -> int:
        """Transport of the last allreduceAccum: REDUCE_PEER_MEMORY (one kernel over NVLink peer memory) or REDUCE_NCCL."""
        v = C.c_int()
        check(_capi.lib().atx_last_reduce_kind(self._h, C.byref(v)))
        return v.value


class Ataraxia:
    """Headless mirror of the `Ataraxia` layer (Engine/src/main.cpp:8-283): what the application does AROUND
    Renderer::Render, i.e. the caller-side semantics a user of the path sees without the window.

      onUpdate(ts, input)   main.cpp:22-32: camera motion resets the accumulation; global transforms refreshed
      Render(frames)        main.cpp:211-220: onResize, camera.Resize, Renderer::Render, wall-clock "Last Render Time"
      ImportScene/ExportScene  main.cpp:196-209
      the UI widgets of onGuiRender (main.cpp:34-176) as methods, each with the widget's own reset behaviour:
      node / sphere / camera edits call resetFrameIndex(); material, light, "Sky Light" and "Ray Depth" edits do
      NOT (main.cpp:44-45, 146-176) — and since the scene is only re-uploaded when frameIndex == 1
      (Renderer.cu:175-179), material and light edits stay invisible until the next reset. That is the
      reference's behaviour and the default here; `eagerEdits=True` resets on those edits too.
    """

    def __init__(self, device: int = 0, eagerEdits: bool = False):
        import time as _time
        self._clock = _time.perf_counter
        self.m_camera = Camera(45.0, 0.1, 100.0)
        self.m_renderer = Renderer(device)
        self.m_scene = Scene()
        self.m_scene.camera = self.m_camera.copy()
        self.m_scene.settings = self.m_renderer.getSettings()
        self.m_viewportWidth = self.m_viewportHeight = 0
        self.m_lastRenderTime = 0.0
        self.eagerEdits = eagerEdits
        self._initializeScene()

    def close(self):
        self.m_renderer.close()

    # ---- Layer interface ------------------------------------------------------
    def onUpdate(self, ts: float, inp: Optional[InputState] = None):
        if self.m_camera.onUpdate(ts, inp):
            self.m_renderer.resetFrameIndex()
            self.m_scene.camera = self.m_camera.copy()
            self.m_scene.settings = self.m_renderer.getSettings()
        self.m_scene.rootNode.updateGlobalTransform()

    def setViewport(self, width: int, height: int):
        """ImGui::GetContentRegionAvail of the "Viewport" window (main.cpp:181-182)."""
        self.m_viewportWidth, self.m_viewportHeight = int(width), int(height)

    def Render(self, frames: int = 1, readback: bool = True):
        t0 = self._clock()
        self.m_renderer.onResize(self.m_viewportWidth, self.m_viewportHeight)
        self.m_camera.Resize(self.m_viewportWidth, self.m_viewportHeight)
        self.m_renderer.Render(self.m_camera, self.m_scene, frames=frames, readback=readback)
        self.m_lastRenderTime = (self._clock() - t0) * 1e3

    def ImportScene(self, path: str = "scene.json"):
        from . import utils
        self.m_scene = utils.importScene(path)
        self.m_camera = self.m_scene.camera.copy()
        self.m_renderer.setSettings(self.m_scene.settings)
        self.m_renderer.resetFrameIndex()

    def ExportScene(self, path: str = "scene.json"):
        from . import utils
        self.m_scene.camera = self.m_camera.copy()
        self.m_scene.settings = self.m_renderer.getSettings()
        utils.exportScene(self.m_scene, path)

    def GetRenderer(self): return self.m_renderer
    def GetScene(self): return self.m_scene
    def SetScene(self, scene: Scene): self.m_scene = scene
    def lastRenderTimeMs(self) -> float: return self.m_lastRenderTime

    # ---- "Settings" window (main.cpp:38-66) -------------------------------------
    def setAccumulation(self, on: bool): self.m_renderer.getSettings().accumulation = bool(on)
    def resetFrameIndex(self): self.m_renderer.resetFrameIndex()

    def setSkyLight(self, on: bool):
        self.m_renderer.getSettings().skyLight = bool(on)
        self._edited()

    def setMaxBounces(self, n: int):
        self.m_renderer.getSettings().maxBounces = max(1, min(500, int(n)))  # DragInt range, main.cpp:45
        self._edited()

    def setFov(self, fov: float):
        self.m_camera = Camera(float(fov), 0.1, 100.0, self.m_camera.getPosition(), self.m_camera.getDirection())
        self.m_scene.camera = self.m_camera.copy()
        self.m_renderer.resetFrameIndex()

    def resetCamera(self):
        self.m_camera = Camera(45.0, 0.1, 100.0)
        self.m_scene.camera = self.m_camera.copy()
        self.m_renderer.resetFrameIndex()

    # ---- "Hierarchy" window (main.cpp:75-143) and the "Add" menu (:292-300) ----------
    def setNodePosition(self, node: SceneNode, p): node.setPosition(p); self.m_renderer.resetFrameIndex()
    def setNodeRotation(self, node: SceneNode, q_xyzw): node.setRotation(q_xyzw); self.m_renderer.resetFrameIndex()
    def setNodeScale(self, node: SceneNode, s): node.setScale(s); self.m_renderer.resetFrameIndex()

    def removeNode(self, node: SceneNode):
        self.m_scene.rootNode.removeChild(node)   # only direct children of the root are found (main.cpp:105)
        self.m_renderer.resetFrameIndex()

    def setSphereCenter(self, node: SceneNode, index: int, c): node.getSpheres()[index].center = tuple(_vec3(c)); self.m_renderer.resetFrameIndex()
    def setSphereRadius(self, node: SceneNode, index: int, r: float): node.getSpheres()[index].radius = float(r); self.m_renderer.resetFrameIndex()
    def setSphereMaterial(self, node: SceneNode, index: int, m: int): node.getSpheres()[index].id = int(m); self.m_renderer.resetFrameIndex()

    def addSphere(self):
        """Menu "Add > Sphere" (main.cpp:294-298): no reset — it shows up at the next one."""
        self.m_scene.rootNode.addSphere(Sphere((0.0, 0.0, 0.0), 1.0, 0))
        self._edited()

    # ---- "Material settings" / "Light settings" (main.cpp:145-176): no reset in the reference ----
    def editMaterial(self, index: int, **fields):
        for k, v in fields.items():
            if not hasattr(self.m_scene.materials[index], k):
                raise AttributeError(k)
            setattr(self.m_scene.materials[index], k, v)
        self._edited()

    def editLight(self, index: int, **fields):
        for k, v in fields.items():
            if not hasattr(self.m_scene.lights[index], k):
                raise AttributeError(k)
            setattr(self.m_scene.lights[index], k, v)
        self._edited()

    def _edited(self):
        if self.eagerEdits:
            self.m_renderer.resetFrameIndex()

    # ---- scripted camera path ------------------------------------------------------
    def playCameraPath(self, steps, frames_per_step: int = 1, on_frame=None):
        """Run a scripted sequence of (ts, InputState) through onUpdate + Render, as the main loop would
        (Application::run: onUpdate, then onGuiRender -> Render, once per UI frame). `on_frame(i, app)` is called
        after every step; returns the frame index after each step."""
        out = []
        for i, (ts, inp) in enumerate(steps):
            self.onUpdate(ts, inp)
            self.Render(frames_per_step)
            out.append(self.m_renderer.frameIndex())
            if on_frame is not None:
                on_frame(i, self)
        return out

    def _initializeScene(self):
        """main.cpp:234-265."""
        root = self.m_scene.rootNode
        root.addSphere(Sphere((0.0, 0.0, 0.0), 1.0, 0))
        child = SceneNode("ChildNode1")
        child.setPosition((2.0, 0.0, 0.0))
        child.addSphere(Sphere((0.0, 0.0, 0.0), 1.0, 1))
        root.addChild(child)
        grand = SceneNode("GrandChildNode")
        grand.setPosition((0.0, 2.0, 0.0))
        grand.addSphere(Sphere((0.0, 0.0, 0.0), 1.0, 2))
        child.addChild(grand)
        # Material(albedo, roughness, metallic, emissionColor, emissionIntensity, id): F0 keeps its 0.04 default (Scene.h:43-46)
        self.m_scene.materials.append(Material(albedo=(1.022, 0.782, 0.344), roughness=1.0, metallic=0.0, id=0))
        self.m_scene.materials.append(Material(albedo=(1.0, 0.0, 0.0), roughness=0.3, metallic=0.0, id=1))
        self.m_scene.materials.append(Material(albedo=(0.972, 0.960, 0.915), roughness=0.25, metallic=1.0, id=2))
        self.m_scene.lights.append(Light((10.0, 10.0, 0.0), (1.0, 1.0, 1.0), 1.0))
