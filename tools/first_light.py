"""Dev script (GPU box): first comparison of the CUDA path against the reference CUDA renderer."""
import json, sys, time, tempfile, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ataraxia_b200 as atx
from oracle import bindings as ob

def compare(scene_path, W, H, bounces, sky, frames, label):
    scene = atx.Utils.importScene(str(scene_path))
    cam = atx.Camera(scene.camera.getFov(), 0.1, 100.0, scene.camera.getPosition(), scene.camera.getDirection())
    r = atx.Renderer(0)
    r.setSettings(atx.Settings(True, sky, bounces))
    r.onResize(W, H); cam.Resize(W, H)
    info, ref = ob.run_ref_headless(scene_path, W, H, bounces, sky, frames, dump_at=(1, frames))
    r.uploadScene(scene); r.setCamera(cam)
    rays = r.getRayDirections(); hits = r.getHitIds()
    print(f"[{label}] rays bit-exact: {rays.tobytes()==ref['rays'].tobytes()}  mismatching px: {int((rays!=ref['rays']).any(-1).sum())}")
    print(f"[{label}] hit ids equal: {(hits==ref['hit']).all()}  mismatching px: {int((hits!=ref['hit']).sum())}")
    r.Render(cam, scene, frames=1)
    a1 = r.getAccumulation(); rg1 = r.getRGBA8()
    d = (a1.view(np.uint32) != ref['acc1'].view(np.uint32))
    print(f"[{label}] acc after frame 1 bit-exact: {not d.any()}  mismatching px: {int(d.any(-1).sum())} maxabs {np.abs(a1-ref['acc1']).max()}")
    print(f"[{label}] rgba1 equal: {(rg1==ref['rgba1']).all()}")
    if frames > 1:
        r.Render(cam, scene, frames=frames-1)
        aK = r.getAccumulation(); rgK = r.getRGBA8()
        d = (aK.view(np.uint32) != ref[f'acc{frames}'].view(np.uint32))
        print(f"[{label}] acc after frame {frames} bit-exact: {not d.any()}  mismatching px: {int(d.any(-1).sum())} maxabs {np.abs(aK-ref[f'acc{frames}']).max()}")
        print(f"[{label}] rgba{frames} equal: {(rgK==ref[f'rgba{frames}']).all()}  w=={frames}: {(aK[...,3]==frames).all()}")
    print(f"[{label}] ref timing: {info}")
    r.close()

def timing(scene, W, H, bounces, frames, label, reps=3):
    cam = atx.Camera(scene.camera.getFov(), 0.1, 100.0, scene.camera.getPosition(), scene.camera.getDirection())
    r = atx.Renderer(0)
    r.setSettings(atx.Settings(True, False, bounces))
    r.onResize(W, H); cam.Resize(W, H)
    r.uploadScene(scene); r.setCamera(cam)
    for i in range(reps):
        r.resetFrameIndex(); r.resetCounters()
        r.Render(cam, scene, frames=frames, readback=False); r.sync()
        ms = r.lastRenderMs(); c = r.counters()
        print(f"[{label}] {W}x{H} {frames} spp {bounces} bounces: {ms:.3f} ms  {c.paths/ms/1e3:.1f} Mpaths/s {c.rays/ms/1e6:.3f} Grays/s  tests {c.sphere_tests/ms/1e6:.2f} G/s  flop-frac {(19*c.sphere_tests+7*c.rays)/(ms*1e-3)/74.4e12:.3f}")
    r.close()

if __name__ == "__main__":
    sc = ob.REF_SCENE
    compare(sc, 320, 180, 5, False, 4, "scene.json 320x180")
    compare(sc, 1280, 720, 5, False, 2, "C1")
    compare(sc, 640, 360, 8, True, 8, "scene.json sky 8b")
    with tempfile.TemporaryDirectory() as td:
        small = atx.synthetic.small()
        p = os.path.join(td, "small.json"); atx.Utils.exportScene(small, p)
        compare(p, 256, 144, 8, True, 8, "small synthetic")
        c3 = atx.synthetic.config3()
        p3 = os.path.join(td, "c3.json"); atx.Utils.exportScene(c3, p3)
        compare(p3, 480, 270, 8, False, 4, "C3 scene @480x270")
    scene = atx.Utils.importScene(str(sc))
    timing(scene, 1280, 720, 5, 1, "C1")
    timing(scene, 1920, 1080, 8, 64, "C2/16")
    timing(c3, 3840, 2160, 8, 2, "C3 2spp")
    c4 = atx.synthetic.config4()
    timing(c4, 3840, 2160, 8, 1, "C4 1spp", reps=2)
