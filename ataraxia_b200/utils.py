"""scene.json import/export — mirror of /root/reference/Engine/src/Utils.cpp (namespace Utils).

Schema (Utils.cpp:3-187): camera{position[3],direction[3],fov}; sceneGraph{name,
transformation{position[3],rotation[x,y,z,w],scale[3]}, spheres[{center[3],radius,materialIndex}],
children[...]}; materials[{albedo[3],roughness,metallic,F0[3],emissionIntensity,emissionColor[3]}]
(Material::id is not serialised); lights[{position[3],intensity,color[3]}];
settings{maxBounces,skyLight,accumulation}.

Numbers are parsed as doubles and narrowed to float32 on assignment, exactly as nlohmann::json ->
float does; they are written back as the float32 value widened to double, shortest round-trip
repr, keys sorted, 4-space indent (nlohmann's std::map ordering and dump(4), Utils.cpp:59).
"""
from __future__ import annotations

import json
from typing import Any, Dict

import numpy as np

from .api import Camera, Light, Material, Scene, SceneNode, Settings, Sphere


def _f(x) -> float:
    """float32 narrowed value widened back to a Python float (what nlohmann stores for a float)."""
    return float(np.float32(x))


def _v3(v):
    return [_f(v[0]), _f(v[1]), _f(v[2])]


def serializeSceneNode(node: SceneNode) -> Dict[str, Any]:
    """Utils.cpp:64-95."""
    j: Dict[str, Any] = {
        "name": node.getName(),
        "transformation": {
            "position": _v3(node.getPosition()),
            "rotation": [_f(c) for c in node.getRotation()],  # x, y, z, w
            "scale": _v3(node.getScale()),
        },
    }
    if node.getSpheres():
        j["spheres"] = [{"center": _v3(s.center), "radius": _f(s.radius), "materialIndex": int(s.id)}
                        for s in node.getSpheres()]
    if node.getChildren():
        j["children"] = [serializeSceneNode(c) for c in node.getChildren()]
    return j


def serializeScene(scene: Scene) -> Dict[str, Any]:
    """Utils.cpp:3-51."""
    scene.rootNode.updateGlobalTransform()
    j: Dict[str, Any] = {
        "camera": {"position": _v3(scene.camera.getPosition()), "direction": _v3(scene.camera.getDirection()),
                   "fov": _f(scene.camera.getFov())},
    }
    if scene.rootNode is not None:
        j["sceneGraph"] = serializeSceneNode(scene.rootNode)
    if scene.materials:
        j["materials"] = [{"albedo": _v3(m.albedo), "roughness": _f(m.roughness), "metallic": _f(m.metallic),
                           "F0": _v3(m.F0), "emissionIntensity": _f(m.emissionIntensity),
                           "emissionColor": _v3(m.emissionColor)} for m in scene.materials]
    if scene.lights:
        j["lights"] = [{"position": _v3(l.position), "intensity": _f(l.intensity), "color": _v3(l.color)}
                       for l in scene.lights]
    j["settings"] = {"maxBounces": int(scene.settings.maxBounces), "skyLight": bool(scene.settings.skyLight),
                     "accumulation": bool(scene.settings.accumulation)}
    return j


def exportScene(scene: Scene, filename: str) -> None:
    """Utils.cpp:53-62."""
    with open(filename, "w") as f:
        f.write(json.dumps(serializeScene(scene), indent=4, sort_keys=True))


def deserializeSceneNode(j: Dict[str, Any], node: SceneNode | None = None) -> SceneNode:
    """Utils.cpp:139-173."""
    if node is None:
        node = SceneNode(j["name"])
    t = j["transformation"]
    node.setPosition(t["position"])
    node.setRotation(t["rotation"])  # file order is x, y, z, w (Utils.cpp:145 builds quat(w=rot[3], x, y, z))
    node.setScale(t["scale"])
    for s in j.get("spheres", []) or []:
        node.addSphere(Sphere(tuple(_f(c) for c in s["center"]), _f(s["radius"]), int(s["materialIndex"])))
    for c in j.get("children", []) or []:
        node.addChild(deserializeSceneNode(c))
    return node


def deserializeScene(j: Dict[str, Any]) -> Scene:
    """Utils.cpp:97-137. Missing keys raise (nlohmann throws; KeyError here)."""
    scene = Scene()
    scene.camera = Camera()  # default-constructed, setters only (quirk Q-cam ii)
    scene.camera.setPosition(j["camera"]["position"])
    scene.camera.setDirection(j["camera"]["direction"])
    scene.camera.setFov(j["camera"]["fov"])
    if "sceneGraph" in j:
        deserializeSceneNode(j["sceneGraph"], scene.rootNode)
        scene.rootNode.updateGlobalTransform()
    for m in j["materials"] or []:
        scene.materials.append(Material(albedo=_v3(m["albedo"]), roughness=_f(m["roughness"]), metallic=_f(m["metallic"]),
                                        F0=_v3(m["F0"]), emissionIntensity=_f(m["emissionIntensity"]),
                                        emissionColor=_v3(m["emissionColor"])))
    for l in j["lights"] or []:
        scene.lights.append(Light(position=_v3(l["position"]), intensity=_f(l["intensity"]), color=_v3(l["color"])))
    s = j["settings"]
    scene.settings = Settings(accumulation=bool(s["accumulation"]), skyLight=bool(s["skyLight"]),
                              maxBounces=int(s["maxBounces"]))
    return scene


def importScene(filename: str) -> Scene:
    """Utils.cpp:175-187: a missing file yields an empty Scene()."""
    try:
        with open(filename, "r") as f:
            j = json.load(f)
    except FileNotFoundError:
        return Scene()
    return deserializeScene(j)
