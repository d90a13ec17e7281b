"""Dev script (GPU box): small renders through every megakernel form, to be run under compute-sanitizer
(memcheck / racecheck): `compute-sanitizer --tool racecheck python tools/sanitize_wq.py`."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import ataraxia_b200 as atx

GOLDEN = Path(__file__).resolve().parents[1] / "tests" / "golden"
cases = [(atx.Utils.importScene(str(GOLDEN / "sample_scene.json")), 96, 54, 8, False, 48),
         (atx.synthetic.small(12, 3, seed=9), 64, 36, 6, True, 40)]
for scene, W, H, bounces, sky, frames in cases:
    cam = atx.Camera(scene.camera.getFov(), 0.1, 100.0, scene.camera.getPosition(), scene.camera.getDirection())
    r = atx.Renderer(0)
    r.setSettings(atx.Settings(True, sky, bounces))
    r.onResize(W, H); cam.Resize(W, H)
    r.uploadScene(scene); r.setCamera(cam)
    ref = None
    for kind in (atx.MEGA_WHILE_WHILE, atx.MEGA_WARP_QUEUE, atx.MEGA_PAIR):
        r.setTuning(atx.TUNE_MEGA_KIND, kind)
        r.renderFrames(1, frames, 1, zero_first=True)
        acc = r.getAccumulation()
        if ref is None:
            ref = acc
        assert (acc.view(np.uint32) == ref.view(np.uint32)).all(), kind
    r.close()
print("sanitize_wq: all forms rendered, bit-identical")
