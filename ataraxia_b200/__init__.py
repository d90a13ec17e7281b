"""ataraxia_b200 — B200-native implementation of Ataraxia's per-pixel path-tracing hot path.

Host mirror of the reference API (api.py, utils.py) over the C-ABI (include/ataraxia_b200.h,
csrc/). Importing the package never touches the oracle; the render path needs the CUDA
extension (ataraxia_b200/lib/libataraxia_b200.so) and fails loudly without it.
"""
from .api import (Ataraxia, Camera, Image, InputState, Light, Material, Renderer, Scene, SceneNode, Settings, Sphere,  # noqa: F401
                  pack_lights, pack_materials, pack_spheres, traverseSceneGraph)
from . import utils as Utils  # noqa: F401  (namespace Utils, Engine/include/Utils.h)
from . import synthetic  # noqa: F401
from ._capi import (AtxError, LIGHT_DTYPE, MATERIAL_DTYPE, SPHERE_DTYPE, TUNE_CHUNK_SPHERES, TUNE_MEGA_KIND, TUNE_PARK_THRESHOLD, TUNE_CLAIM_THRESHOLD, TUNE_REDUCE, REDUCE_NONE, REDUCE_PEER_MEMORY, REDUCE_NCCL,  # noqa: F401
                    MEGA_AUTO, MEGA_PAIR, MEGA_PAIR_LOCKSTEP, MEGA_WHILE_WHILE, MEGA_WARP_QUEUE,
                    VARIANT_AUTO, VARIANT_MEGAKERNEL, VARIANT_WAVEFRONT)

__all__ = ["Ataraxia", "InputState", "Camera", "Image", "Light", "Material", "Renderer", "Scene", "SceneNode", "Settings", "Sphere", "Utils",
           "synthetic", "traverseSceneGraph", "pack_spheres", "pack_materials", "pack_lights", "AtxError"]
