import sys, os, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ataraxia_b200 as atx
from oracle import bindings as ob

def run(scene_path, W, H, bounces, sky, label):
    scene = atx.Utils.importScene(str(scene_path))
    cam = atx.Camera(scene.camera.getFov(), 0.1, 100.0, scene.camera.getPosition(), scene.camera.getDirection())
    r = atx.Renderer(0)
    r.setSettings(atx.Settings(True, sky, bounces))
    r.onResize(W, H); cam.Resize(W, H)
    info, ref = ob.run_ref_headless(scene_path, W, H, bounces, sky, 1, dump_at=(1,))
    r.Render(cam, scene, frames=1)
    a = r.getAccumulation(); hits = r.getHitIds()
    bad = (a.view(np.uint32) != ref['acc1'].view(np.uint32)).any(-1)
    print(f"[{label} b={bounces}] mismatching {int(bad.sum())} / {W*H}")
    sph = ref['spheres']
    for hid in np.unique(hits):
        sel = hits == hid
        mat = int(sph[hid,4]) if hid>=0 else -1
        print(f"   primary hit {hid} (mat {mat}): px {int(sel.sum())} bad {int((bad&sel).sum())}")
    ys, xs = np.nonzero(bad)
    for k in range(min(4, len(ys))):
        y, x = ys[k], xs[k]
        print("   ", (x, y), "hit", hits[y,x], "mine", a[y,x,:3], "ref", ref['acc1'][y,x,:3], "ulp", (a[y,x,:3].view(np.int32)-ref['acc1'][y,x,:3].view(np.int32)))
    r.close()

sc = ob.REF_SCENE
for b in (1, 2, 3):
    run(sc, 320, 180, b, False, "scene.json")
with tempfile.TemporaryDirectory() as td:
    small = atx.synthetic.small()
    p = os.path.join(td, "small.json"); atx.Utils.exportScene(small, p)
    for b in (1, 2):
        run(p, 256, 144, b, True, "small")
