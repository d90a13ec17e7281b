import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def built():
    """The native libraries, built in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    from ataraxia_b200 import build
    build.build_product()
    build.build_oracle()
    return build


@pytest.fixture(scope="session")
def port(built):
    from oracle.bindings import OraclePort
    return OraclePort()


@pytest.fixture(scope="session")
def refcpu(built):
    """The unmodified reference compiled for the host; absent where oracle/ref/Makefile never ran."""
    from oracle import bindings
    if not bindings.have_reference_cpu():
        pytest.skip("oracle/_ref/libref_cpu.so not built (needs /root/reference)")
    return bindings.ReferenceCpu()


@pytest.fixture(scope="session")
def sample_scene_path():
    return GOLDEN / "sample_scene.json"


def random_graph_scene(atx, rng):
    """A random three-level scene graph: arbitrary (unnormalised) quaternions, non-uniform and negative scales,
    nested translations, several spheres per node (some with out-of-range material ids)."""
    import numpy as np
    scene = atx.Scene()
    scene.materials = [atx.Material(albedo=tuple(rng.uniform(0, 1, 3)), roughness=float(rng.uniform(0, 1))) for _ in range(3)]
    scene.lights = [atx.Light((1.0, 5.0, 2.0), (1.0, 1.0, 1.0), 1.0)]
    scene.camera = atx.Camera(45.0, 0.1, 100.0, (0.0, 1.0, 8.0), (0.0, 0.0, -1.0))

    def fill(node, depth):
        node.setPosition(rng.normal(0, 3, 3))
        node.setRotation(rng.normal(0, 1, 4) if rng.random() < 0.8 else np.zeros(4))   # all-zero = identity quirk
        node.setScale(rng.uniform(0.2, 3.0, 3) * rng.choice([1.0, 1.0, -1.0], 3))
        for _ in range(int(rng.integers(0, 4))):
            node.addSphere(atx.Sphere(tuple(rng.normal(0, 2, 3)), float(rng.uniform(0.1, 2.0)), int(rng.integers(-1, 5))))
        if depth < 3:
            for k in range(int(rng.integers(1, 3))):
                child = atx.SceneNode(f"n{depth}_{k}")
                fill(child, depth + 1)
                node.addChild(child)

    fill(scene.rootNode, 0)
    return scene
