"""Generates tests/golden/c2_full.json: SHA-256 digests of what the UNMODIFIED reference CUDA renderer
(oracle/_ref/ref_headless, built by oracle/ref/Makefile from /root/reference) produces for BASELINE config 2
at its full size — sample scene, 1920x1080, frames 1..1024, 8 bounces — on a B200:

    gpurun -- python tests/golden/make_golden_c2.py gpurun_out/golden

then copy gpurun_out/golden/c2_full.json to tests/golden/. bench.py compares the accumulation buffer of its
timed config-2 step with `acc1024_sha256` (the `verified` key of the bench line), so the driver-run record
carries bit-exactness against the reference without the reference binary having to be there.
"""
import hashlib
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import bindings as ob  # noqa: E402

G = Path(__file__).resolve().parent


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    W, H, bounces, frames = 1920, 1080, 8, 1024
    info, ref = ob.run_ref_headless(G / "sample_scene.json", W, H, bounces, False, frames, dump_at=(1, frames), timeout=900)
    out = {"generator": "tests/golden/make_golden_c2.py (oracle/_ref/ref_headless: unmodified reference sources, sm_100a)",
           "scene": "tests/golden/sample_scene.json", "width": W, "height": H, "max_bounces": bounces, "sky": False, "frames": frames,
           "acc1_sha256": digest(ref["acc1"]), "acc1024_sha256": digest(ref[f"acc{frames}"]),
           "rgba1024_sha256": digest(ref[f"rgba{frames}"]), "rays_sha256": digest(ref["rays"]), "hit_sha256": digest(ref["hit"]),
           "reference_median_frame_ms": info["median_frame_ms"], "reference_median_kernel_ms": info.get("median_kernel_ms")}
    Path(outdir, "c2_full.json").write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
