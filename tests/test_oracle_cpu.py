"""CPU tests (-m "not gpu"): the oracle port (oracle/oracle.cpp) against the golden vectors that the
UNMODIFIED reference, compiled for the host, produced (tests/golden/make_golden_cpu.py), and — where
oracle/_ref/libref_cpu.so exists — against that reference library live. All comparisons bit-exact:
both sides are IEEE float32 with the same evaluation order and the same libm.
"""
import json

import numpy as np
import pytest

from conftest import GOLDEN
from oracle.bindings import LIGHT_DTYPE, MATERIAL_DTYPE, SPHERE_DTYPE


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN / "cpu_golden.npz")


def _scene(gold, name):
    s = gold[f"{name}_spheres"].view(SPHERE_DTYPE)
    m = gold[f"{name}_materials"].view(MATERIAL_DTYPE)
    l = gold[f"{name}_lights"].view(LIGHT_DTYPE)
    W, H, bounces, sky, frames = (int(v) for v in gold[f"{name}_dims"])
    return s, m, l, W, H, bounces, bool(sky), frames


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


# ---- Random.h:59-70 -----------------------------------------------------------------------------
SURVEY_KATS = [(0, 129708002), (1, 2831084092), (2, 2055130248), (12345, 4099845390), (921599, 3618561272),
               (0xFFFFFFFF, 3861530882)]


def test_pcg_hash_known_answers(port, gold):
    for seed, expect in SURVEY_KATS:
        assert port.pcg_hash(seed) == expect
    for seed, expect in zip(gold["pcg_seeds"], gold["pcg_hash"]):
        assert port.pcg_hash(int(seed)) == int(expect)


def test_pcg_float_chain(port, gold):
    s = 0
    for f_expect, s_expect in zip(gold["pcg_chain_float"], gold["pcg_chain_seed"]):
        f, s = port.pcg_float(s)
        assert s == int(s_expect)
        assert np.float32(f).view(np.uint32) == np.float32(f_expect).view(np.uint32)
    # 0.030199997, 0.19039936, 0.4994767 from seed 0
    assert np.allclose(gold["pcg_chain_float"][:3], [0.030199997, 0.19039936, 0.4994767], rtol=1e-7)


def test_pcg_float_range_includes_one(port):
    # float(0xFFFFFFFF) rounds to 2^32, so PcgFloat can return exactly 1.0 (quirk Q-rr)
    # find a seed whose hash is >= 0xFFFFFF80 by brute force over a small range is not guaranteed;
    # check the arithmetic instead: the float conversion of the largest hash value divides to 1.0
    assert np.float32(np.uint32(0xFFFFFFFF)) / np.float32(4294967296.0) == np.float32(1.0)


# ---- SceneNode.cpp:42-59 + Renderer.cu:67-96 ------------------------------------------------------
@pytest.mark.parametrize("name,file", [("sample", "sample_scene.json"), ("small", "small_scene.json")])
def test_flatten_matches_reference(port, gold, name, file):
    j = json.load(open(GOLDEN / file))
    mine = port.flatten_json(j)
    expect = gold[f"{name}_spheres"].view(SPHERE_DTYPE)
    assert mine.tobytes() == expect.tobytes()


def test_sample_scene_flatten_values(gold):
    s = gold["sample_spheres"].view(SPHERE_DTYPE)
    # S0 (0,0,0) r1 m0; S1 (2,-101,0) r100 m1; S2 (3,0,0) r1 m2
    assert s["center"].tolist() == [[0, 0, 0], [2, -101, 0], [3, 0, 0]]
    assert s["radius"].tolist() == [1, 100, 1] and s["material"].tolist() == [0, 1, 2]


# ---- Camera.cpp:134-195 ---------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["sample", "small"])
def test_camera_matrices_and_rays(port, gold, name):
    _, _, _, W, H, _, _, _ = _scene(gold, name)
    rays, ip, iv = port.camera(gold[f"{name}_campos"], gold[f"{name}_camdir"], float(gold[f"{name}_fov"]), 0.1, 100.0, W, H)
    assert (bits(ip) == bits(gold[f"{name}_invproj"])).all()
    assert (bits(iv) == bits(gold[f"{name}_invview"])).all()
    assert (bits(rays) == bits(gold[f"{name}_rays"])).all()
    # rays are unit length to float precision and y = 0 is the image bottom (no V flip in the engine)
    assert np.allclose(np.linalg.norm(rays, axis=-1), 1.0, atol=1e-6)
    assert rays[0, W // 2, 1] < rays[H - 1, W // 2, 1]


def test_c1_primary_visibility(port, gold):
    s = gold["sample_spheres"].view(SPHERE_DTYPE)
    rays, _, _ = port.camera(gold["sample_campos"], gold["sample_camdir"], float(gold["sample_fov"]), 0.1, 100.0, 1280, 720)
    hits = port.primary_hits(s, gold["sample_campos"], rays)
    hist = [(hits == k).sum() for k in (-1, 0, 1, 2)]
    assert hist == gold["c1_hit_histogram"].tolist()
    # the survey's float64 re-derivation: 364231 / 35974 / 490333 / 31062 (+- a few silhouette pixels)
    assert np.abs(np.array(hist) - np.array([364231, 35974, 490333, 31062])).max() <= 16
    assert hits[360, 640] == 1  # centre pixel sees the ground sphere
    assert (bits(rays[360, 640]) == bits(gold["c1_center_ray"])).all()


# ---- Renderer.cu:251-409 + BRDF.cu ---------------------------------------------------------------
@pytest.mark.parametrize("name", ["sample", "small"])
def test_render_matches_reference(port, gold, name):
    s, m, l, W, H, bounces, sky, frames = _scene(gold, name)
    rays = gold[f"{name}_rays"]
    hits = port.primary_hits(s, gold[f"{name}_campos"], rays)
    assert (hits == gold[f"{name}_hits"]).all()
    acc = port.render(s, m, l, gold[f"{name}_campos"], rays, 1, 1, 1, bounces, sky)
    assert (bits(acc) == bits(gold[f"{name}_acc1"])).all()
    acc = port.render(s, m, l, gold[f"{name}_campos"], rays, 2, frames - 1, 1, bounces, sky, accum=acc)
    assert (bits(acc) == bits(gold[f"{name}_accK"])).all()
    assert (acc[..., 3] == frames).all()  # .w is the exact sample count
    rgba = port.pack_rgba8(acc, frames)
    assert (rgba == gold[f"{name}_rgbaK"]).all()
    assert ((rgba >> 24) == 255).all()  # alpha always 255 (quirk Q-pack)


def test_render_edge_cases(port, gold):
    s, m, l, W, H, bounces, sky, frames = _scene(gold, "small")
    rays = gold["small_rays"]
    pos = gold["small_campos"]
    # maxBounces 0: the path loop does not run, every sample is (0,0,0,1)
    acc = port.render(s, m, l, pos, rays, 1, 3, 1, 0, True)
    assert (acc[..., :3] == 0).all() and (acc[..., 3] == 3).all()
    # no spheres, sky on: every pixel gets the sky constant once
    empty = np.zeros(0, SPHERE_DTYPE)
    acc = port.render(empty, m, l, pos, rays, 1, 1, 1, 4, True)
    assert np.allclose(acc[0, 0], [0.6, 0.7, 0.9, 1.0])
    # no lights: only emission and sky contribute, still finite
    acc = port.render(s, m, np.zeros(0, LIGHT_DTYPE), pos, rays, 1, 2, 1, 4, False)
    assert np.isfinite(acc).all()
    # frame stride: frames {1,3} + frames {2,4} visit the same samples as frames 1..4
    a = port.render(s, m, l, pos, rays, 1, 2, 2, bounces, sky)
    b = port.render(s, m, l, pos, rays, 2, 2, 2, bounces, sky)
    c = port.render(s, m, l, pos, rays, 1, 4, 1, bounces, sky)
    assert ((a + b)[..., 3] == 4).all()
    assert np.allclose(a + b, c, rtol=1e-5, atol=1e-6)
    # ragged row range leaves other rows untouched
    acc = port.render(s, m, l, pos, rays, 1, 1, 1, bounces, sky, rows=(5, 9))
    assert (acc[:5] == 0).all() and (acc[9:] == 0).all() and (acc[5:9, :, 3] == 1).all()


# ---- live cross-check against the host-compiled reference, where it has been built ----------------
def test_port_equals_reference_live(port, refcpu, tmp_path):
    import ataraxia_b200 as atx
    scene = atx.synthetic.small(n_spheres=17, n_lights=2, seed=99)
    p = tmp_path / "s.json"
    atx.Utils.exportScene(scene, str(p))
    s, m, l, info = refcpu.load_scene(p)
    W, H = 80, 45
    r_ref, ip_ref, iv_ref = refcpu.camera(info["position"], info["direction"], info["fov"], 0.1, 100.0, W, H)
    r_port, ip, iv = port.camera(info["position"], info["direction"], info["fov"], 0.1, 100.0, W, H)
    assert (bits(r_ref) == bits(r_port)).all() and (bits(ip) == bits(ip_ref)).all() and (bits(iv) == bits(iv_ref)).all()
    assert (refcpu.primary_hits(s, info["position"], r_ref) == port.primary_hits(s, info["position"], r_ref)).all()
    for sky in (False, True):
        a = refcpu.render(s, m, l, info["position"], r_ref, 1, 3, 1, 6, sky)
        b = port.render(s, m, l, info["position"], r_ref, 1, 3, 1, 6, sky)
        assert (bits(a) == bits(b)).all()
        assert (refcpu.pack_rgba8(a, 3) == port.pack_rgba8(b, 3)).all()
    for seed in (0, 1, 77, 0xDEADBEEF):
        assert refcpu.pcg_hash(seed) == port.pcg_hash(seed)


def test_camera_update_port_matches_reference_golden_and_product(port, gold):
    """The port's Camera::onUpdate (Camera.cpp:30-108) against the reference's per-step camera states (golden, 160
    scripted steps) and, on a second random script, against the product's host helper (atx_host_camera_update)."""
    from oracle.bindings import CAMERA_STEP_DTYPE
    import ataraxia_b200 as atx
    steps = gold["walk_steps"].view(CAMERA_STEP_DTYPE)
    op, od, om = port.camera_walk(gold["sample_campos"], gold["sample_camdir"], steps)
    assert (op.view(np.uint32) == gold["walk_pos"].view(np.uint32)).all()
    assert (od.view(np.uint32) == gold["walk_dir"].view(np.uint32)).all()
    assert (om == gold["walk_moved"]).all()
    rng = np.random.default_rng(99)
    n = 400
    steps = np.zeros(n, CAMERA_STEP_DTYPE)
    steps["dt"] = rng.uniform(0.0, 0.2, n)
    steps["mouse_x"] = np.cumsum(rng.normal(0, 60, n)); steps["mouse_y"] = np.cumsum(rng.normal(0, 60, n))
    steps["keys"] = rng.integers(0, 64, n); steps["right"] = rng.random(n) < 0.75
    start_p, start_d = np.array([1.0, 2.0, 3.0], np.float32), np.array([0.3, -0.2, -0.93], np.float32)
    op, od, om = port.camera_walk(start_p, start_d, steps)
    cam = atx.Camera(50.0, 0.1, 100.0, start_p, start_d)
    for i, st in enumerate(steps):
        keys = "".join(c for b, c in enumerate("WSADQE") if (int(st["keys"]) >> b) & 1)
        moved = cam.onUpdate(float(st["dt"]), atx.InputState(keys, (float(st["mouse_x"]), float(st["mouse_y"])), bool(st["right"])))
        assert moved == om[i]
        assert (cam.getPosition().view(np.uint32) == op[i].view(np.uint32)).all(), i
        assert (cam.getDirection().view(np.uint32) == od[i].view(np.uint32)).all(), i
