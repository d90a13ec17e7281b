"""Generates the GPU golden vectors by running the UNMODIFIED reference CUDA renderer
(oracle/_ref/ref_headless, built by oracle/ref/Makefile from /root/reference) on a B200:

    gpurun -- python tests/golden/make_golden_gpu.py gpurun_out/golden

then copy gpurun_out/golden/gpu_golden.npz to tests/golden/. These pin bit-exactness of the CUDA
path on machines where the reference binary is not available.
"""
import hashlib
import os
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import bindings as ob  # noqa: E402
import ataraxia_b200 as atx  # noqa: E402

G = Path(__file__).resolve().parent


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def main(outdir):
    out = {}
    os.makedirs(outdir, exist_ok=True)
    # full-resolution small cases
    for name, path, W, H, bounces, sky, frames in [("sample", G / "sample_scene.json", 160, 90, 5, False, 8),
                                                   ("small", G / "small_scene.json", 128, 72, 8, True, 4)]:
        info, ref = ob.run_ref_headless(path, W, H, bounces, sky, frames, dump_at=(1, frames))
        out[f"{name}_dims"] = np.array([W, H, bounces, int(sky), frames], np.int32)
        out[f"{name}_hits"] = ref["hit"].astype(np.int16)
        out[f"{name}_acc1"] = ref["acc1"]
        out[f"{name}_accK"] = ref[f"acc{frames}"]
        out[f"{name}_rgbaK"] = ref[f"rgba{frames}"]
    # C1 (BASELINE config 1) at full size: hit map + digest + one row
    info, ref = ob.run_ref_headless(G / "sample_scene.json", 1280, 720, 5, False, 1, dump_at=(1,))
    out["c1_hits"] = ref["hit"].astype(np.int8)
    out["c1_acc1_sha256"] = sha(ref["acc1"])
    out["c1_rgba1_sha256"] = sha(ref["rgba1"])
    out["c1_rays_sha256"] = sha(ref["rays"])
    out["c1_acc1_row360"] = ref["acc1"][360]
    # config-3 scene (256 spheres, 16 lights) at reduced size: digests + hit map
    with tempfile.TemporaryDirectory() as td:
        p3 = os.path.join(td, "c3.json")
        atx.Utils.exportScene(atx.synthetic.config3(), p3)
        info, ref = ob.run_ref_headless(p3, 240, 135, 8, False, 2, dump_at=(1, 2))
        out["c3_dims"] = np.array([240, 135, 8, 0, 2], np.int32)
        out["c3_hits"] = ref["hit"].astype(np.int16)
        out["c3_acc1_sha256"] = sha(ref["acc1"])
        out["c3_acc2_sha256"] = sha(ref["acc2"])
        out["c3_acc2_row67"] = ref["acc2"][67]
    np.savez_compressed(os.path.join(outdir, "gpu_golden.npz"), **out)
    print("wrote", os.path.join(outdir, "gpu_golden.npz"))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
