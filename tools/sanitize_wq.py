"""Dev script (GPU box): small renders through every megakernel form, to be run under compute-sanitizer
(memcheck / racecheck): `compute-sanitizer --tool racecheck python tools/sanitize_wq.py`."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import ataraxia_b200 as atx

GOLDEN = Path(__file__).resolve().parents[1] / "tests" / "golden"
cases = [(atx.Utils.importScene(str(GOLDEN / "sample_scene.json")), 96, 54, 8, False, 48),
         (atx.synthetic.small(12, 3, seed=9), 64, 36, 6, True, 40),
         (atx.synthetic.small(70, 2, seed=5), 64, 36, 6, True, 12)]      # three blocks of 32 filter steps, the last one partial
for scene, W, H, bounces, sky, frames in cases:
    cam = atx.Camera(scene.camera.getFov(), 0.1, 100.0, scene.camera.getPosition(), scene.camera.getDirection())
    r = atx.Renderer(0)
    r.setSettings(atx.Settings(True, sky, bounces))
    r.onResize(W, H); cam.Resize(W, H)
    r.uploadScene(scene); r.setCamera(cam)
    ref = None
    for kind, chunk in ((atx.MEGA_WHILE_WHILE, 0), (atx.MEGA_WARP_QUEUE, 0), (atx.MEGA_PAIR, 0), (atx.MEGA_PAIR_LOCKSTEP, 0),
                        (atx.MEGA_PAIR, 24), (atx.MEGA_PAIR_LOCKSTEP, 24)):          # chunk 24: double-buffered TMA staging
        r.setTuning(atx.TUNE_MEGA_KIND, kind)
        r.setTuning(atx.TUNE_CHUNK_SPHERES, chunk)
        r.renderFrames(1, frames, 1, zero_first=True)
        acc = r.getAccumulation()
        if ref is None:
            ref = acc
        assert (acc.view(np.uint32) == ref.view(np.uint32)).all(), (kind, chunk)
        r.renderFrames(1, 0, 1, zero_first=True)                                     # image-tile shares (strided pixel pool)
        for share in range(3):
            r.renderTileShare(1, frames, 3, share, zero_first=False)
        assert (r.getAccumulation().view(np.uint32) == ref.view(np.uint32)).all(), (kind, chunk, "tiles")
    r.close()
print("sanitize_wq: all forms rendered, bit-identical")
