"""The header-only C++ mirror of the reference API (include/ataraxia/Ataraxia.h) over the C-ABI:
built with plain g++ (no nvcc, no glm, no nlohmann) and driven like reference user code.

CPU part: scene.json import, scene-graph flatten, camera matrices and the host ray table are
bit-identical to the Python mirror (which the other tests pin to the reference), and the C++ export is
read back identically by the Python mirror and — where it was built — by the reference's own reader.
GPU part (-m gpu): Renderer::Render through the C++ classes against the reference-CUDA golden vectors."""
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def driver(built, tmp_path_factory):
    out = tmp_path_factory.mktemp("cpp") / "mirror_main"
    lib = ROOT / "ataraxia_b200" / "lib"
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "tests" / "cpp" / "mirror_main.cpp"),
           "-o", str(out), f"-L{lib}", "-lataraxia_b200", f"-Wl,-rpath,{lib}"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    return out


@pytest.fixture(scope="module")
def example(built, tmp_path_factory):
    """examples/render_scene.cpp: the headless command-line use of the C++ mirror."""
    out = tmp_path_factory.mktemp("example") / "render_scene"
    lib = ROOT / "ataraxia_b200" / "lib"
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "examples" / "render_scene.cpp"),
           "-o", str(out), f"-L{lib}", "-lataraxia_b200", f"-Wl,-rpath,{lib}"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    return out


def f32(x):
    return np.asarray(x, np.float64).astype(np.float32)


def test_example_builds_and_prints_usage(example):
    proc = subprocess.run([str(example)], capture_output=True, text=True)
    assert proc.returncode == 2 and "usage:" in proc.stderr


def test_json_reader_writer_is_locale_independent_and_strict(driver):
    """include/ataraxia/Json.h: '.' decimals whatever LC_NUMERIC says (a host application may have called
    setlocale), JSON number grammar only, control characters escaped, nesting capped."""
    import locale
    names = ["C"]
    for cand in ("de_DE.UTF-8", "de_DE.utf8", "fr_FR.UTF-8", "de_DE"):
        try:
            locale.setlocale(locale.LC_NUMERIC, cand)
            names.append(cand)
            break
        except locale.Error:
            continue
    locale.setlocale(locale.LC_NUMERIC, "C")
    for name in names:
        proc = subprocess.run([str(driver), "json", name], capture_output=True, text=True)
        assert proc.returncode == 0, (name, proc.stdout, proc.stderr)
        assert proc.stdout.split()[0] == "0"


@pytest.mark.parametrize("file", ["sample_scene.json", "small_scene.json"])
def test_cpp_host_side_matches_python_mirror(driver, built, tmp_path, file):
    import ataraxia_b200 as atx
    exported = tmp_path / "export.json"
    proc = subprocess.run([str(driver), "cpu", str(GOLDEN / file), str(exported)], capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    got = json.loads(proc.stdout)
    scene = atx.Utils.importScene(str(GOLDEN / file))
    spheres = atx.pack_spheres(atx.traverseSceneGraph(scene.rootNode))
    assert len(got["spheres"]) == len(spheres)
    g = np.array(got["spheres"], np.float64)
    assert (f32(g[:, :3]).view(np.uint32) == spheres["center"].view(np.uint32)).all()
    assert (f32(g[:, 3]).view(np.uint32) == spheres["radius"].view(np.uint32)).all()
    assert (g[:, 4].astype(np.int32) == spheres["material"]).all()
    assert got["materials"] == len(scene.materials) and got["lights"] == len(scene.lights)
    assert got["maxBounces"] == scene.settings.maxBounces and got["rootName"] == "Scene"
    cam = atx.Camera(scene.camera.getFov(), 0.1, 100.0, scene.camera.getPosition(), scene.camera.getDirection())
    cam.Resize(160, 90)
    for key, m in (("projection", cam.getProjectionMatrix()), ("view", cam.getViewMatrix()),
                   ("inverseProjection", cam.getInverseProjectionMatrix()), ("inverseView", cam.getInverseViewMatrix())):
        assert (f32(got[key]).view(np.uint32) == np.asarray(m, np.float32).reshape(16).view(np.uint32)).all(), key
    rays = cam.getRayDirection().reshape(-1, 3)
    assert got["nRays"] == len(rays)
    assert (f32(got["ray0"]).view(np.uint32) == rays[0].view(np.uint32)).all()
    assert (f32(got["rayLast"]).view(np.uint32) == rays[-1].view(np.uint32)).all()
    # the C++ export is the same scene for the Python reader
    back = atx.Utils.importScene(str(exported))
    s2 = atx.pack_spheres(atx.traverseSceneGraph(back.rootNode))
    assert (s2.view(np.uint8) == spheres.view(np.uint8)).all()
    assert (atx.pack_materials(back.materials).view(np.uint8) == atx.pack_materials(scene.materials).view(np.uint8)).all()
    assert (atx.pack_lights(back.lights).view(np.uint8) == atx.pack_lights(scene.lights).view(np.uint8)).all()
    assert json.loads(exported.read_text()) == atx.Utils.serializeScene(scene)


def test_cpp_flattens_random_scene_graphs_like_python(driver, built, tmp_path):
    """Nested nodes with arbitrary quaternions and negative scales through the C++ mirror's importer and traversal."""
    import ataraxia_b200 as atx
    from conftest import random_graph_scene
    rng = np.random.default_rng(4242)
    for trial in range(3):
        p = tmp_path / f"graph{trial}.json"
        atx.Utils.exportScene(random_graph_scene(atx, rng), str(p))
        proc = subprocess.run([str(driver), "cpu", str(p)], capture_output=True, text=True)
        assert proc.returncode == 0, proc.stderr
        g = np.array(json.loads(proc.stdout)["spheres"], np.float64)
        spheres = atx.pack_spheres(atx.traverseSceneGraph(atx.Utils.importScene(str(p)).rootNode))
        assert len(g) == len(spheres) and len(g) > 0
        assert (f32(g[:, :3]).view(np.uint32) == spheres["center"].view(np.uint32)).all()
        assert (f32(g[:, 3]).view(np.uint32) == spheres["radius"].view(np.uint32)).all()
        assert (g[:, 4].astype(np.int32) == spheres["material"]).all()


def test_reference_reads_the_cpp_export(driver, refcpu, tmp_path):
    """The reference's own Utils::importScene + traverseSceneGraph on a file written by the C++ mirror."""
    import ataraxia_b200 as atx
    exported = tmp_path / "export.json"
    assert subprocess.run([str(driver), "cpu", str(GOLDEN / "small_scene.json"), str(exported)], capture_output=True).returncode == 0
    scene = atx.Utils.importScene(str(GOLDEN / "small_scene.json"))
    s, m, l, info = refcpu.load_scene(exported)
    assert s.tobytes() == atx.pack_spheres(atx.traverseSceneGraph(scene.rootNode)).tobytes()
    assert m.tobytes() == atx.pack_materials(scene.materials).tobytes()
    assert l.tobytes() == atx.pack_lights(scene.lights).tobytes()


def test_cpp_scripted_camera_walk_matches_reference_golden(driver, tmp_path):
    """Camera::onUpdate through the C++ mirror, per-step camera state against the reference's (cpu_golden.npz)."""
    gold = np.load(GOLDEN / "cpu_golden.npz")
    steps_path, out = tmp_path / "steps.bin", tmp_path / "walk.bin"
    steps_path.write_bytes(gold["walk_steps"].tobytes())
    start = [repr(float(v)) for v in list(gold["sample_campos"]) + list(gold["sample_camdir"]) + [gold["sample_fov"]]]
    proc = subprocess.run([str(driver), "walk", str(steps_path), str(out)] + start, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    raw = np.fromfile(out, np.float32)
    n = len(gold["walk_moved"])
    per = raw[:n * 23].reshape(n, 23)
    assert (per[:, 0:3].view(np.uint32) == gold["walk_pos"].view(np.uint32)).all()
    assert (per[:, 3:6].view(np.uint32) == gold["walk_dir"].view(np.uint32)).all()
    assert (per[:, 6:22].view(np.uint32) == gold["walk_invview"].view(np.uint32)).all()
    assert ((per[:, 22] == 1.0) == gold["walk_moved"]).all()
    assert (raw[n * 23:].view(np.uint32) == gold["walk_rays"].reshape(-1).view(np.uint32)).all()


def test_image_sinks_cpp_and_python(driver, tmp_path):
    """Image::savePPM / savePNG (the headless stand-in for the Vulkan texture upload, Core/src/Image.cpp:183-271):
    both mirrors write the same pixels, top row first (the UI's V flip, main.cpp:185-187), r in the low byte
    (Renderer.h:70-78); the PNG is read back with PIL."""
    from PIL import Image as PILImage
    import ataraxia_b200 as atx
    W, H = 37, 11
    x, y = np.meshgrid(np.arange(W, dtype=np.uint32), np.arange(H, dtype=np.uint32))
    px = (0xFF000000 | ((x * 7) & 0xFF) | (((y * 23) & 0xFF) << 8) | ((((x + y) * 5) & 0xFF) << 16)).astype(np.uint32)
    expect = np.stack([px & 0xFF, (px >> 8) & 0xFF, (px >> 16) & 0xFF, px >> 24], -1).astype(np.uint8)[::-1]
    assert subprocess.run([str(driver), "image", str(tmp_path / "cpp")], capture_output=True).returncode == 0
    img = atx.Image(W, H)
    img.setData(px)
    img.savePPM(str(tmp_path / "py.ppm")); img.savePNG(str(tmp_path / "py.png"))
    for who in ("cpp", "py"):
        png = np.asarray(PILImage.open(tmp_path / f"{who}.png"))
        assert png.shape == (H, W, 4) and (png == expect).all(), who
        ppm = np.asarray(PILImage.open(tmp_path / f"{who}.ppm"))
        assert (ppm == expect[..., :3]).all(), who
    assert (tmp_path / "cpp.ppm").read_bytes() == (tmp_path / "py.ppm").read_bytes()


def test_cpp_missing_file_and_missing_key(driver, tmp_path):
    # a missing file is an empty Scene (Utils.cpp:178-179): no spheres, no materials
    proc = subprocess.run([str(driver), "cpu", str(tmp_path / "nope.json")], capture_output=True, text=True)
    assert proc.returncode == 0 and '"spheres": []' in proc.stdout and '"materials": 0, "lights": 0, "maxBounces": 15' in proc.stdout
    bad = tmp_path / "bad.json"
    bad.write_text('{"camera": {"position": [0,0,0], "direction": [0,0,-1]}}')   # no fov: nlohmann throws, so do we
    proc = subprocess.run([str(driver), "cpu", str(bad)], capture_output=True, text=True)
    assert proc.returncode != 0


@pytest.mark.gpu
def test_cpp_renderer_bit_exact_vs_reference_cuda_golden(driver, tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    gold = np.load(GOLDEN / "gpu_golden.npz")
    for name, file in (("sample", "sample_scene.json"), ("small", "small_scene.json")):
        W, H, bounces, sky, frames = (int(v) for v in gold[f"{name}_dims"])
        out = tmp_path / f"{name}.bin"
        proc = subprocess.run([str(driver), "gpu", str(GOLDEN / file), str(W), str(H), str(bounces), str(sky), str(frames), str(out),
                               str(tmp_path / name)], capture_output=True, text=True)
        assert proc.returncode == 0, proc.stdout + proc.stderr
        raw = np.fromfile(out, np.uint32)
        P = W * H
        hits = raw[:P].view(np.int32).reshape(H, W)
        acc1 = raw[P:5 * P].reshape(H, W, 4)
        accK = raw[5 * P:9 * P].reshape(H, W, 4)
        rgba = raw[9 * P:10 * P].reshape(H, W)
        assert (hits == gold[f"{name}_hits"]).all()
        assert (acc1 == gold[f"{name}_acc1"].view(np.uint32)).all()
        assert (accK == gold[f"{name}_accK"].view(np.uint32)).all()
        assert (rgba == gold[f"{name}_rgbaK"]).all()
        # output sinks: PPM is the RGBA8 image flipped top to bottom, PFM the float radiance bottom row first
        ppm = (tmp_path / f"{name}.ppm").read_bytes()
        header = f"P6\n{W} {H}\n255\n".encode()
        assert ppm.startswith(header)
        rgb = np.frombuffer(ppm[len(header):], np.uint8).reshape(H, W, 3)
        assert (rgb[::-1, :, 0] == (rgba & 0xFF)).all() and (rgb[::-1, :, 2] == ((rgba >> 16) & 0xFF)).all()
        pfm = (tmp_path / f"{name}.pfm").read_bytes()
        ph = f"PF\n{W} {H}\n-1.0\n".encode()
        assert pfm.startswith(ph)
        rad = np.frombuffer(pfm[len(ph):], np.float32).reshape(H, W, 3)
        assert np.allclose(rad, accK.view(np.float32)[..., :3] / frames, rtol=1e-6, atol=0)


@pytest.mark.gpu
def test_cpp_application_layer_matches_python(driver, tmp_path):
    """The Ataraxia layer (main.cpp:8-283) in C++: frame-index behaviour of camera motion, edits and import,
    and the image it ends with equals the Python mirror's after the same session."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ataraxia_b200 as atx
    out = tmp_path / "app.bin"
    proc = subprocess.run([str(driver), "app", str(GOLDEN / "small_scene.json"), str(out)], capture_output=True, text=True)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    assert proc.stdout.split()[:6] == ["2", "5", "1", "3", "1", "3"]
    app = atx.Ataraxia()
    app.setViewport(64, 36)
    app.ImportScene(str(GOLDEN / "small_scene.json"))
    app.Render(2)
    acc = app.GetRenderer().getAccumulation()
    assert (np.fromfile(out, np.uint32) == acc.view(np.uint32).reshape(-1)).all()
    app.close()


@pytest.mark.gpu
def test_example_renders_the_same_image_as_the_python_mirror(example, tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from PIL import Image as PILImage
    import ataraxia_b200 as atx
    png, pfm = tmp_path / "out.png", tmp_path / "out.pfm"
    proc = subprocess.run([str(example), str(GOLDEN / "sample_scene.json"), str(png), "--width", "200", "--height", "120", "--spp", "40",
                           "--bounces", "6", "--sky", "1", "--pfm", str(pfm)], capture_output=True, text=True)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    assert "Mpaths/s" in proc.stdout
    scene = atx.Utils.importScene(str(GOLDEN / "sample_scene.json"))
    cam = atx.Camera(scene.camera.getFov(), 0.1, 100.0, scene.camera.getPosition(), scene.camera.getDirection())
    r = atx.Renderer(0)
    r.setSettings(atx.Settings(True, True, 6))
    r.onResize(200, 120); cam.Resize(200, 120)
    r.Render(cam, scene, frames=40)
    px = r.getImage().data
    expect = np.stack([px & 0xFF, (px >> 8) & 0xFF, (px >> 16) & 0xFF, px >> 24], -1).astype(np.uint8)[::-1]
    assert (np.asarray(PILImage.open(png)) == expect).all()
    ref_pfm = tmp_path / "py.pfm"
    r.saveAccumulationPFM(str(ref_pfm))
    assert pfm.read_bytes() == ref_pfm.read_bytes()
    r.close()
