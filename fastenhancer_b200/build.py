"""In-tree build of the CUDA engine: nvcc -> fastenhancer_b200/libfastenhancer_b200.so (sm_100a only).

The shared library exports the C ABI of include/fastenhancer_b200.h.  Object files go to
fastenhancer_b200/_build/; both are git-ignored but travel with the repo snapshot to the GPU box.
nvcc cross-compiles without a GPU, so this also runs on the build container.
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "_build")
LIB = os.path.join(PKG, "libfastenhancer_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA engine cannot be built (there is no CPU fallback)")


def _newest(paths) -> float:
    return max(os.path.getmtime(p) for p in paths)


def _compile(nvcc: str, src: str, obj: str, verbose: bool) -> str:
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return r.stderr


def build(force: bool = False, jobs: int | None = None, verbose: bool = False) -> str:
    """Compile every translation unit that is out of date and link the shared library."""
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(PKG, "..", "include", "fastenhancer_b200.h")]
    hdr_time = _newest(hdrs)
    os.makedirs(OBJ, exist_ok=True)
    todo, objs = [], []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_time):
            todo.append((s, o))
    if not todo and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest(objs):
        return LIB
    nvcc = _nvcc()
    if todo:
        with cf.ThreadPoolExecutor(max_workers=jobs or min(len(todo), os.cpu_count() or 4)) as ex:
            for log in ex.map(lambda so: _compile(nvcc, so[0], so[1], verbose), todo):
                if verbose:
                    sys.stderr.write(log)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
