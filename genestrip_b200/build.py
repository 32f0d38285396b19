"""Builds the CUDA extension (sm_100a) in-tree: genestrip_b200/_lib/libgenestrip_b200.so.

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box with the
gpurun snapshot.  There is no CPU fallback: if this library is missing the package refuses to work.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libgenestrip_b200.so")
SOURCES = ["gs_kernels.cu", "gs_text.cu", "gs_inflate.cu", "gs_capi.cu", "gs_host.cpp"]
# host-only translation units for the host compiler (run-time dispatched AVX2 bodies; nvcc's front end is kept away from them)
HOST_SOURCES = ["gs_pack.cpp"]
HEADERS = ["gs_kernels.cuh", "gs_device.cuh", "gs_host.hpp", "gs_pack.hpp", os.path.join("..", "..", "include", "genestrip_b200.h")]
CXX_FLAGS = ["-O3", "-std=c++17", "-fPIC", "-pthread", "-c"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-pthread", "-shared"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HOST_SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_native(force=False, verbose=False, defines=(), out=None):
    """Compile every CUDA/C++ source of the package into one shared library (`defines`/`out`: tuning variants)."""
    out = out or LIB_PATH
    if not force and out == LIB_PATH and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs = []
    for hs in HOST_SOURCES:
        obj = out + "." + os.path.splitext(hs)[0] + ".o"
        res = subprocess.run([os.environ.get("CXX") or shutil.which("g++") or "g++"] + CXX_FLAGS + ["-o", obj, os.path.join(CSRC, hs)], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
        objs.append(obj)
    cmd = [_nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out + ".tmp"] + srcs + objs + ["-lz", "-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    for obj in objs:
        os.remove(obj)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(out + ".tmp", out)
    if verbose:
        print(res.stderr)
    return out


if __name__ == "__main__":
    print(build_native(force=True, verbose=True))
