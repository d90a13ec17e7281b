"""CPU tests (-m "not gpu"): the drop-in boundary and the host mirror of the reference API.

 * the C-ABI library loads and exports every symbol include/ataraxia_b200.h declares (no compute
   calls without a GPU; atx_create must FAIL loudly here — there is no CPU fallback);
 * POD layouts equal the reference's (SURVEY.md §8c KATs);
 * host math helpers (camera matrices, ray table, node transforms, flatten) are bit-identical to
   the golden vectors produced by the reference's own host code;
 * scene.json import/export round-trips and matches the reference's reader;
 * reference quirks of Camera/Renderer that callers depend on.
"""
import ctypes as C
import json
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
import ataraxia_b200 as atx
from ataraxia_b200 import _capi


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN / "cpu_golden.npz")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_library_exports_every_declared_symbol(built):
    header = (ROOT / "include" / "ataraxia_b200.h").read_text()
    declared = sorted(set(re.findall(r"ATX_API\s+[\w\s\*]+?\b(atx_\w+)\s*\(", header)))
    assert declared == sorted(_capi.SYMBOLS), "binding list and header disagree"
    lib = _capi.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
    assert b"sm_100a" in lib.atx_version()


def test_header_is_plain_c_and_links(built, tmp_path):
    """include/ataraxia_b200.h as a C11 translation unit (-pedantic -Werror), linked against the library and run
    (host helpers only: no GPU needed) — the view a cgo / JNI / ctypes binding has of the boundary."""
    import subprocess
    lib = ROOT / "ataraxia_b200" / "lib"
    exe = tmp_path / "c_abi"
    proc = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", f"-I{ROOT / 'include'}", str(ROOT / "tests" / "c" / "c_abi.c"),
                           "-o", str(exe), f"-L{lib}", "-lataraxia_b200", f"-Wl,-rpath,{lib}"], capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0, run.stdout + run.stderr
    # W + D for 0.1 s at speed 5 from (0,0,3) looking down -z, then the mouse look: forward 0.5, right 0.5
    x, y, z = (float(v) for v in run.stdout.split()[-3:])
    assert abs(x - 0.5) < 1e-6 and y == 0.0 and abs(z - 2.5) < 1e-6


def test_no_cpu_fallback(built):
    """Without a CUDA device the product refuses to create a renderer (and says why)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    status = _capi.lib().atx_create(0, C.byref(h))
    assert status == _capi.ATX_ERR_NO_DEVICE
    assert b"no CPU fallback" in _capi.lib().atx_last_error()
    with pytest.raises(atx.AtxError):
        atx.Renderer(0)


def test_product_never_touches_the_oracle():
    for path in list((ROOT / "ataraxia_b200").rglob("*.py")) + list((ROOT / "ataraxia_b200" / "csrc").glob("*")) + \
            list((ROOT / "include").rglob("*.h")):
        if path.name == "build.py" or path.is_dir():
            continue  # build.py compiles the checker (never loads it): building is not using
        text = path.read_text(errors="ignore")
        if path.suffix == ".py":
            assert "liboracle" not in text and "from oracle" not in text and "import oracle" not in text, path
            assert "libref_cpu" not in text and "ref_headless" not in text, path
        else:  # native sources: no include of, and no string literal naming, anything under oracle/
            for line in text.splitlines():
                code = line.split("//")[0]
                assert not ("#include" in code and "oracle" in code), (path, line)
                assert not re.search(r'"[^"]*(oracle|libref_cpu|ref_headless)[^"]*"', code), (path, line)


def test_pod_layouts():
    # sizeof(Sphere)=20, Material=52, Light=28
    assert _capi.SPHERE_DTYPE.itemsize == 20
    assert _capi.MATERIAL_DTYPE.itemsize == 52
    assert _capi.LIGHT_DTYPE.itemsize == 28
    assert _capi.MATERIAL_DTYPE.fields["F0"][1] == 20 and _capi.MATERIAL_DTYPE.fields["emissionIntensity"][1] == 44
    assert _capi.LIGHT_DTYPE.fields["intensity"][1] == 24
    header = (ROOT / "include" / "ataraxia_b200.h").read_text()
    for field in ("center[3]", "radius", "albedo[3]", "roughness", "metallic", "F0[3]", "emissionColor[3]",
                  "emissionIntensity", "position[3]", "color[3]", "intensity"):
        assert field in header


# ---- Camera ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["sample", "small"])
def test_camera_matches_reference(built, gold, name):
    W, H = int(gold[f"{name}_dims"][0]), int(gold[f"{name}_dims"][1])
    cam = atx.Camera(float(gold[f"{name}_fov"]), 0.1, 100.0, gold[f"{name}_campos"], gold[f"{name}_camdir"])
    cam.Resize(W, H)
    assert (bits(cam.getInverseProjectionMatrix()) == bits(gold[f"{name}_invproj"])).all()
    assert (bits(cam.getInverseViewMatrix()) == bits(gold[f"{name}_invview"])).all()
    assert (bits(cam.getRayDirection()) == bits(gold[f"{name}_rays"])).all()


def test_camera_quirks(built):
    cam = atx.Camera(45.0, 0.1, 100.0)
    assert (cam.m_width, cam.m_height) == (1600, 900)          # Camera.cpp:15
    assert cam.getPosition().tolist() == [0, 0, 3] and cam.getDirection().tolist() == [0, 0, -1]
    before = cam.getInverseProjectionMatrix().copy()
    cam.Resize(1600, 900)                                      # early return, Camera.cpp:118-119
    assert (cam.getInverseProjectionMatrix() == before).all()
    cam.Resize(0, 10)                                          # prints and returns, Camera.cpp:112-116
    assert (cam.m_width, cam.m_height) == (1600, 900)
    # a camera that only went through the setters (deserializeScene) keeps an identity view matrix
    c2 = atx.Camera()
    c2.setPosition((1, 2, 3)); c2.setDirection((0, 0, -1)); c2.setFov(60.0)
    c2.Resize(64, 32)
    assert (c2.getInverseViewMatrix() == np.eye(4, dtype=np.float32).reshape(-1)).all()
    assert not (c2.getInverseProjectionMatrix() == np.eye(4, dtype=np.float32).reshape(-1)).all()


def _walk(cam, steps):
    pos, dirs, iv, moved = [], [], [], []
    for st in steps:
        keys = "".join(c for b, c in enumerate("WSADQE") if (int(st["keys"]) >> b) & 1)
        moved.append(cam.onUpdate(float(st["dt"]), atx.InputState(keys, (float(st["mouse_x"]), float(st["mouse_y"])), bool(st["right"]))))
        pos.append(cam.getPosition().copy()); dirs.append(cam.getDirection().copy()); iv.append(cam.getInverseViewMatrix().copy())
    return np.array(pos), np.array(dirs), np.array(iv), np.array(moved)


def test_scripted_camera_walk_matches_reference_golden(built, gold):
    """Camera::onUpdate (Camera.cpp:30-108) under 160 scripted input steps: position, direction, inverse view
    matrix and the returned flag after EVERY step, and the final ray table, bit for bit against what the
    unmodified reference produced with the same inputs (tests/golden/make_golden_cpu.py)."""
    from oracle.bindings import CAMERA_STEP_DTYPE
    steps = gold["walk_steps"].view(CAMERA_STEP_DTYPE)
    cam = atx.Camera(float(gold["sample_fov"]), 0.1, 100.0, gold["sample_campos"], gold["sample_camdir"])
    cam.Resize(48, 27)
    pos, dirs, iv, moved = _walk(cam, steps)
    assert (bits(pos) == bits(gold["walk_pos"])).all()
    assert (bits(dirs) == bits(gold["walk_dir"])).all()
    assert (bits(iv) == bits(gold["walk_invview"])).all()
    assert (moved == gold["walk_moved"]).all() and 0 < moved.sum() < len(moved)
    assert (bits(cam.getRayDirection()) == bits(gold["walk_rays"])).all()


def test_scripted_camera_walk_matches_live_reference(built, refcpu):
    """The same against the reference library itself on a fresh random script (only where it was built)."""
    from oracle.bindings import CAMERA_STEP_DTYPE
    rng = np.random.default_rng(77)
    n = 300
    steps = np.zeros(n, CAMERA_STEP_DTYPE)
    steps["dt"] = rng.uniform(0.0, 0.1, n)
    steps["mouse_x"] = np.cumsum(rng.normal(0, 40, n)); steps["mouse_y"] = np.cumsum(rng.normal(0, 40, n))
    steps["keys"] = rng.integers(0, 64, n); steps["right"] = rng.random(n) < 0.7
    start_p, start_d = np.array([3.0, 1.0, -4.0], np.float32), np.array([-0.6, -0.1, 0.79], np.float32)
    op, od, oiv, om, rays = refcpu.camera_walk(start_p, start_d, 60.0, 0.1, 100.0, 40, 30, steps)
    cam = atx.Camera(60.0, 0.1, 100.0, start_p, start_d)
    cam.Resize(40, 30)
    pos, dirs, iv, moved = _walk(cam, steps)
    assert (bits(pos) == bits(op)).all() and (bits(dirs) == bits(od)).all() and (bits(iv) == bits(oiv)).all()
    assert (moved == om).all()
    assert (bits(cam.getRayDirection()) == bits(rays)).all()


def test_camera_update_rejects_null_arguments(built):
    import ctypes as C
    lib = _capi.lib()
    pos, d, last = (np.zeros(n, np.float32) for n in (3, 3, 2))
    inp = _capi.CameraInput(0, 0, 0.0, 0.0)
    null_f = C.POINTER(C.c_float)()
    assert lib.atx_host_camera_update(null_f, _capi.fptr(d), _capi.fptr(last), C.byref(inp), 0.1, None) == _capi.ATX_ERR_INVALID
    assert lib.atx_host_camera_update(_capi.fptr(pos), _capi.fptr(d), _capi.fptr(last), None, 0.1, None) == _capi.ATX_ERR_INVALID
    assert b"null" in lib.atx_last_error()
    assert lib.atx_host_camera_update(_capi.fptr(pos), _capi.fptr(d), _capi.fptr(last), C.byref(inp), 0.1, None) == _capi.ATX_OK
    assert lib.atx_last_mega_kind(None, None) == _capi.ATX_ERR_INVALID


def test_camera_update_semantics(built):
    cam = atx.Camera(45.0, 0.1, 100.0)
    p0, d0 = cam.getPosition().copy(), cam.getDirection().copy()
    # button up: nothing moves, but the mouse position is remembered (Camera.cpp:33-40)
    assert cam.onUpdate(0.1, atx.InputState("W", (50.0, 20.0), False)) is False
    assert (cam.getPosition() == p0).all() and cam.m_lastMousePos.tolist() == [50.0, 20.0]
    # button down, same mouse position, W: forward by speed * dt = 0.5 along the direction
    assert cam.onUpdate(0.1, atx.InputState("W", (50.0, 20.0), True)) is True
    assert np.allclose(cam.getPosition(), p0 + d0 * 0.5) and (cam.getDirection() == d0).all()
    # W wins over S, A over D, Q over E (else-if chains, Camera.cpp:51-85)
    c2 = atx.Camera(45.0, 0.1, 100.0)
    c2.onUpdate(0.1, atx.InputState("WSADQE", (0.0, 0.0), True))
    c3 = atx.Camera(45.0, 0.1, 100.0)
    c3.onUpdate(0.1, atx.InputState("WAQ", (0.0, 0.0), True))
    assert (bits(c2.getPosition()) == bits(c3.getPosition())).all()
    # mouse motion alone rotates and keeps the position
    c4 = atx.Camera(45.0, 0.1, 100.0)
    assert c4.onUpdate(0.016, atx.InputState("", (30.0, -10.0), True)) is True
    assert (c4.getPosition() == p0).all() and not (c4.getDirection() == d0).all()
    assert abs(np.linalg.norm(c4.getDirection()) - 1.0) < 1e-5


# ---- scene graph -------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,file", [("sample", "sample_scene.json"), ("small", "small_scene.json")])
def test_import_and_flatten_match_reference(built, gold, name, file):
    scene = atx.Utils.importScene(str(GOLDEN / file))
    flat = atx.pack_spheres(atx.traverseSceneGraph(scene.rootNode))
    n_mat = len(scene.materials)
    flat["material"] = np.where((flat["material"] < 0) | (flat["material"] >= n_mat), 0, flat["material"])
    assert flat.tobytes() == gold[f"{name}_spheres"].tobytes()
    assert atx.pack_materials(scene.materials).tobytes() == gold[f"{name}_materials"].tobytes()
    assert atx.pack_lights(scene.lights).tobytes() == gold[f"{name}_lights"].tobytes()
    assert np.array_equal(scene.camera.getPosition(), gold[f"{name}_campos"])
    assert scene.camera.getFov() == float(gold[f"{name}_fov"])


def test_sample_scene_contents(built):
    scene = atx.Utils.importScene(str(GOLDEN / "sample_scene.json"))
    assert scene.settings.maxBounces == 25 and scene.settings.skyLight is False and scene.settings.accumulation is True
    assert len(scene.materials) == 3 and len(scene.lights) == 1
    assert scene.materials[0].metallic == 1.0 and abs(scene.materials[0].roughness - 0.6) < 1e-6
    assert all(m.id == 0 for m in scene.materials)  # Material::id is not serialised
    # 3-level graph: root -> child -> grandchild
    assert len(scene.rootNode.getChildren()) == 1 and len(scene.rootNode.getChildren()[0].getChildren()) == 1


def test_export_import_round_trip(built, tmp_path):
    scene = atx.synthetic.small(n_spheres=9, n_lights=2, seed=3)
    child = atx.SceneNode("Child")
    child.setPosition((1.0, 2.0, -3.0))
    child.setRotation((0.0, 0.38268343, 0.0, 0.9238795))  # 45 degrees about y
    child.setScale((2.0, 2.0, 2.0))
    child.addSphere(atx.Sphere((0.5, 0.0, 0.0), 0.25, 1))
    scene.rootNode.addChild(child)
    p = tmp_path / "scene.json"
    atx.Utils.exportScene(scene, str(p))
    text = p.read_text()
    j = json.loads(text)
    assert list(j.keys()) == sorted(j.keys())             # nlohmann's std::map ordering
    assert text.startswith('{\n    "camera"')              # dump(4)
    again = atx.Utils.importScene(str(p))
    a = atx.pack_spheres(atx.traverseSceneGraph(scene.rootNode))
    b = atx.pack_spheres(atx.traverseSceneGraph(again.rootNode))
    assert a.tobytes() == b.tobytes()
    assert atx.pack_materials(scene.materials).tobytes() == atx.pack_materials(again.materials).tobytes()
    # the rotated+scaled child: radius scales by the mean column length (Renderer.cu:81-86)
    assert abs(b["radius"][-1] - 0.5) < 1e-6
    # export -> import -> export is a fixed point
    p2 = tmp_path / "scene2.json"
    atx.Utils.exportScene(again, str(p2))
    assert p2.read_text() == text


def test_import_missing_file_gives_empty_scene(built, tmp_path):
    scene = atx.Utils.importScene(str(tmp_path / "nope.json"))   # Utils.cpp:178-179
    assert scene.materials == [] and scene.lights == [] and scene.rootNode.getSpheres() == []


def test_import_missing_key_raises(built, tmp_path):
    p = tmp_path / "bad.json"
    p.write_text(json.dumps({"camera": {"position": [0, 0, 0], "direction": [0, 0, -1], "fov": 45}}))
    with pytest.raises(KeyError):                                # nlohmann throws, uncaught (Utils.cpp:97-137)
        atx.Utils.importScene(str(p))


def test_reference_reads_our_export(built, refcpu, tmp_path):
    scene = atx.synthetic.config3()
    p = tmp_path / "c3.json"
    atx.Utils.exportScene(scene, str(p))
    s, m, l, info = refcpu.load_scene(p)
    assert len(s) == 256 and len(l) == 16 and len(m) == 32
    assert s.tobytes() == atx.pack_spheres(atx.traverseSceneGraph(scene.rootNode)).tobytes()
    assert m.tobytes() == atx.pack_materials(scene.materials).tobytes()
    assert l.tobytes() == atx.pack_lights(scene.lights).tobytes()


def test_random_scene_graphs_flatten_like_the_reference(built, refcpu, tmp_path):
    """Renderer::traverseSceneGraph + SceneNode::updateGlobalTransform (Renderer.cu:67-96, SceneNode.cpp:42-59) on
    random three-level graphs: arbitrary (unnormalised) quaternions, non-uniform and negative scales, nested
    translations, several spheres per node — flattened centres and radii bit for bit against the reference's own
    importer + traversal reading the file this exporter wrote."""
    from conftest import random_graph_scene
    rng = np.random.default_rng(20260101)
    for trial in range(6):
        scene = random_graph_scene(atx, rng)
        p = tmp_path / f"graph{trial}.json"
        atx.Utils.exportScene(scene, str(p))
        s, m, l, info = refcpu.load_scene(p)
        back = atx.Utils.importScene(str(p))               # what the file holds (positions etc. rounded to the JSON text)
        ours = atx.pack_spheres(atx.traverseSceneGraph(back.rootNode))
        ours["material"] = np.where((ours["material"] < 0) | (ours["material"] >= 3), 0, ours["material"])  # Renderer.cu:30-37
        assert len(s) == len(ours) and len(s) > 0
        assert s.tobytes() == ours.tobytes(), trial


def test_scene_node_api(built):
    root = atx.SceneNode("Scene")
    a, b = atx.SceneNode("a"), atx.SceneNode("b")
    root.addChild(a); root.addChild(b); root.removeChild(a)
    assert [c.getName() for c in root.getChildren()] == ["b"]
    root.addSphere(atx.Sphere((0, 0, 0), 1.0, 0)); root.addSphere(atx.Sphere((1, 0, 0), 1.0, 0))
    root.removeSphere(5); root.removeSphere(0)
    assert len(root.getSpheres()) == 1 and root.getSpheres()[0].center[0] == 1
    b.setPosition((2.0, 0.0, 0.0))
    b.addSphere(atx.Sphere((0, 1, 0), 0.5, 7))
    root.setPosition((0.0, 10.0, 0.0))
    flat = atx.traverseSceneGraph(root)
    assert flat[1].center == (2.0, 11.0, 0.0) and flat[1].id == 7   # ids are clamped at upload, not here
    # second traversal takes the non-dirty branch (global = parent * stored local)
    flat2 = atx.traverseSceneGraph(root)
    assert [s.center for s in flat] == [s.center for s in flat2]


def test_frame_partition():
    from ataraxia_b200.distributed import frame_partition
    for total in (0, 1, 7, 8, 1024, 16384):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                sh = frame_partition(total, r, world)
                seen += [sh.first + j * sh.stride for j in range(sh.count)]
            assert sorted(seen) == list(range(1, total + 1))
            counts = [frame_partition(total, r, world).count for r in range(world)]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        frame_partition(4, 2, 2)


def test_sha256_matches_hashlib(built):
    """atx_host_sha256 (the digest checkpoints bind scene and payload with) against hashlib on block-boundary sizes."""
    import ctypes as C
    import hashlib
    from ataraxia_b200 import _capi
    lib = _capi.lib()
    rng = np.random.default_rng(3)
    for n in (0, 1, 55, 56, 57, 63, 64, 65, 119, 120, 1000, 65536 + 3):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        out = (C.c_uint8 * 32)()
        assert lib.atx_host_sha256(data, n, out) == _capi.ATX_OK
        assert bytes(out) == hashlib.sha256(data).digest(), n
    assert lib.atx_host_sha256(None, 4, (C.c_uint8 * 32)()) == _capi.ATX_ERR_INVALID
    assert lib.atx_save_checkpoint(None, b"x", 0, 0) == _capi.ATX_ERR_INVALID
    assert lib.atx_load_checkpoint(None, b"x", None, None) == _capi.ATX_ERR_INVALID
