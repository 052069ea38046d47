"""Build recipe for the native pieces (in-tree; nothing is installed into site-packages).

  libmiplib_cuda.so   CUDA kernels + the C ABI of include/mipcu.h       (nvcc, sm_100a only)
  libmiplib.so        the C++17 front-end mirror (Solver/Var/Expr/Constr/capi) + the
                      Backend::Cuda adapter, linked against libmiplib_cuda.so        (g++)
  miplib_unit_test    the reference's test assertions re-hosted on a small harness  (g++)

`python -m miplib_b200.build` builds whatever is stale.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
CUDA_LIB = PKG / "libmiplib_cuda.so"
FRONT_LIB = PKG / "libmiplib.so"
UNIT_TEST = PKG / "miplib_unit_test"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CXX = os.environ.get("CXX", "g++")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def _run(cmd):
    print("+", " ".join(str(c) for c in cmd), flush=True)
    subprocess.run([str(c) for c in cmd], check=True, cwd=ROOT)


def cuda_sources():
    return sorted(CSRC.glob("*.cu"))


def build_cuda(force: bool = False) -> Path:
    srcs = cuda_sources()
    deps = srcs + sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh")) + [ROOT / "include" / "mipcu.h"]
    if force or _stale(CUDA_LIB, deps):
        _run([NVCC, *NVCC_FLAGS, "-o", CUDA_LIB, *srcs, "-ldl"])
    return CUDA_LIB


def front_sources():
    return sorted((PKG / "miplib").rglob("*.cpp"))


def build_front(force: bool = False) -> Path:
    srcs = front_sources()
    if not srcs:
        return FRONT_LIB
    deps = srcs + sorted((PKG / "miplib").rglob("*.hpp")) + [ROOT / "include" / "mipcu.h"]
    if force or _stale(FRONT_LIB, deps) or _stale(FRONT_LIB, [CUDA_LIB]):
        _run([CXX, "-std=c++17", "-O2", "-g", "-fPIC", "-shared", "-Wall", "-Wextra",
              "-Wno-sign-compare", "-DWITH_CUDA", f"-I{PKG}", f"-I{ROOT / 'include'}",
              "-o", FRONT_LIB, *srcs, f"-L{PKG}", "-lmiplib_cuda", "-Wl,-rpath,$ORIGIN"])
    return FRONT_LIB


def build_unit_test(force: bool = False) -> Path:
    srcs = sorted((ROOT / "tests" / "cpp").glob("*.cpp"))
    if not srcs:
        return UNIT_TEST
    deps = srcs + sorted((ROOT / "tests" / "cpp").glob("*.hpp")) + [FRONT_LIB]
    if force or _stale(UNIT_TEST, deps):
        _run([CXX, "-std=c++17", "-O1", "-g", "-Wall", "-Wextra", "-Wno-sign-compare", "-DWITH_CUDA",
              f"-I{PKG}", f"-I{ROOT / 'include'}", "-o", UNIT_TEST, *srcs,
              f"-L{PKG}", "-lmiplib", "-lmiplib_cuda", "-Wl,-rpath,$ORIGIN"])
    return UNIT_TEST


def build_all(force: bool = False) -> None:
    build_cuda(force)
    build_front(force)
    build_unit_test(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
