"""The drop-in as code: examples/dropin/Renderer.cpp is compiled by plain g++ against the reference's
UNMODIFIED headers (Engine/include/Renderer.h:19-61) and linked with the reference's own Camera.cpp / SceneNode.cpp /
Utils.cpp objects and -lataraxia_b200 into oracle/_ref/ref_headless_shim (oracle/ref/Makefile). Driven by the
application's 3-call protocol (Engine/src/main.cpp:211-220), it must produce what the unmodified reference CUDA
renderer (oracle/_ref/ref_headless) produces on the same scene file: accumulation buffer and the RGBA8 image the
Renderer hands to the application, bit for bit."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN

ROOT = Path(__file__).resolve().parents[1]


def test_dropin_sources_do_not_edit_the_reference_header():
    """The shim includes the reference's Renderer.h as it is: no member is added (the handle lives in a side table)."""
    src = (ROOT / "examples" / "dropin" / "Renderer.cpp").read_text()
    assert '#include "Renderer.h"' in src and "m_backend" not in src
    mk = (ROOT / "oracle" / "ref" / "Makefile").read_text()
    assert "-I$(REF)/Engine/include" in mk and "examples/dropin/Renderer.cpp" in mk


def test_dropin_binary_is_linked_against_the_product(built):
    from oracle import bindings as ob
    if not ob.have_dropin_shim():
        pytest.skip("oracle/_ref/ref_headless_shim not built (needs /root/reference)")
    out = subprocess.run(["ldd", str(ob.REF_HEADLESS_SHIM)], capture_output=True, text=True).stdout
    assert "libataraxia_b200.so" in out
    syms = subprocess.run(["nm", "-C", str(ob.REF_HEADLESS_SHIM)], capture_output=True, text=True).stdout
    assert "Renderer::Render(Camera&, Scene const&)" in syms and "Camera::Resize" in syms
    assert "kernelRender" not in syms                     # none of the reference's device code is in it


@pytest.mark.gpu
@pytest.mark.parametrize("file,W,H,bounces,sky,frames", [("sample_scene.json", 320, 180, 8, False, 12),
                                                         ("small_scene.json", 200, 120, 6, True, 5),
                                                         ("sample_scene.json", 1280, 720, 5, False, 2)])
def test_dropin_renders_what_the_reference_renders(built, file, W, H, bounces, sky, frames):
    import torch
    from oracle import bindings as ob
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not (ob.have_dropin_shim() and ob.have_ref_headless()):
        pytest.skip("oracle/_ref binaries not present")
    info_r, ref = ob.run_ref_headless(GOLDEN / file, W, H, bounces, sky, frames, dump_at=(1, frames))
    info_s, shim = ob.run_dropin_shim(GOLDEN / file, W, H, bounces, sky, frames, dump_at=(1, frames))
    for k in (1, frames):
        assert (shim[f"acc{k}"].view(np.uint32) == ref[f"acc{k}"].view(np.uint32)).all(), k
        assert (shim[f"rgba{k}"] == ref[f"rgba{k}"]).all(), k
    assert info_s["paths"] == W * H * frames and info_s["kernel_launches"] >= frames
    print(f"{file} {W}x{H}: Render() per frame, drop-in {info_s['median_frame_ms']:.3f} ms vs reference {info_r['median_frame_ms']:.3f} ms")
