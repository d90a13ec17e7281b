"""Build the native pieces in-tree (the built .so files travel to the GPU box with the snapshot).

  ataraxia_b200/lib/libataraxia_b200.so   product: sm_100a kernels + C-ABI          (nvcc)
  oracle/liboracle.so                     TEST INFRASTRUCTURE: CPU restatement       (g++)
  oracle/_ref/*                           TEST INFRASTRUCTURE: the unmodified reference,
                                          only when /root/reference is present       (make)

nvcc cross-compiles sm_100a without a GPU, so this runs on the CPU-only build host.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "ataraxia_b200" / "csrc"
LIB = ROOT / "ataraxia_b200" / "lib" / "libataraxia_b200.so"
ORACLE_LIB = ROOT / "oracle" / "liboracle.so"
REFERENCE = Path(os.environ.get("ATX_REFERENCE", "/root/reference"))

NVCC_FLAGS = [
    "-std=c++17", "-O3",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    # device: the reference is built -use_fast_math; its three effects are requested one by
    # one so that nothing else (e.g. host-side fast math) comes along. All float arithmetic
    # in the kernels is explicit PTX (atx_exact.cuh), -ftz only decides the setp flavour.
    "-ftz=true", "-prec-div=false", "-prec-sqrt=false", *os.environ.get("ATX_EXTRA_NVCC", "").split(),
    # host: glm-order math must not be contracted or reassociated
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-fvisibility=hidden",
]


def _newer(target: Path, sources) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(s).stat().st_mtime <= t for s in sources)


def _run(cmd, cwd=None):
    proc = subprocess.run([str(c) for c in cmd], cwd=cwd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout)
        sys.stderr.write(proc.stderr)
        raise RuntimeError("build step failed: " + " ".join(str(c) for c in cmd))
    return proc.stdout + proc.stderr


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built")
    return exe


def build_product(force: bool = False, verbose: bool = False) -> Path:
    sources = sorted(CSRC.glob("*.cu"))
    deps = sources + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + sorted((ROOT / "include").rglob("*.h"))
    if not force and _newer(LIB, deps):
        return LIB
    LIB.parent.mkdir(parents=True, exist_ok=True)
    objs = []
    for src in sources:
        obj = LIB.parent / (src.stem + ".o")
        out = _run([nvcc(), *NVCC_FLAGS, "-Xptxas", "-v", "-c", src, "-o", obj])
        if verbose:
            print(out)
        objs.append(obj)
    _run([nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *objs, "-o", LIB, "-ldl"])
    return LIB


def build_oracle(force: bool = False) -> Path:
    src = ROOT / "oracle" / "oracle.cpp"
    if not src.exists():
        return ORACLE_LIB
    if not force and _newer(ORACLE_LIB, [src]):
        return ORACLE_LIB
    cxx = shutil.which("g++") or "g++"
    _run([cxx, "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-pthread",
          src, "-o", ORACLE_LIB])
    return ORACLE_LIB


def build_reference(force: bool = False) -> bool:
    """Compile the unmodified reference into oracle/_ref (only where /root/reference exists)."""
    if not (REFERENCE / "Engine" / "src" / "Renderer.cu").exists():
        return False
    mk = ROOT / "oracle" / "ref"
    if force:
        _run(["make", "-C", mk, "clean"])
    _run(["make", "-C", mk, "-j8", f"REF={REFERENCE}"])
    return True


def build_all(force: bool = False, verbose: bool = False):
    lib = build_product(force, verbose)
    build_oracle(force)
    build_reference(False)
    return lib


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
    print("built", LIB)
