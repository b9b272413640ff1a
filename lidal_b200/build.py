"""Build lidal_b200/liblidal_b200.so (the C-ABI library) in-tree with nvcc for sm_100a."""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liblidal_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def sources():
    return sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")))


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(HERE, "csrc", "*.cuh")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(out: str, extra_flags) -> str:
    """A/B build of the library with extra nvcc flags (e.g. -DLIDAL_CONV_DEBUG) into ``out``; select it with LIDAL_LIB=out."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = out + ".obj"
    os.makedirs(objdir, exist_ok=True)
    procs, objs = [], []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        procs.append(subprocess.Popen([nvcc, *[f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")], *extra_flags, "-c", src, "-o", obj]))
    for pr in procs:
        if pr.wait() != 0:
            raise RuntimeError("nvcc failed")
    subprocess.check_call([nvcc, "-Wno-deprecated-gpu-targets", "-shared", "-o", out, *objs, "-lcudart"])
    return out


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "csrc", "_obj"), exist_ok=True)
    for src in sources():
        obj = os.path.join(HERE, "csrc", "_obj", os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        procs.append((src, subprocess.Popen([nvcc, *NVCC_FLAGS, "-c", src, "-o", obj],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"== {os.path.basename(src)}\n{out}")
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {src}")
    subprocess.check_call([nvcc, "-Wno-deprecated-gpu-targets", "-shared", "-o", LIB, *objs, "-lcudart"])
    with open(os.path.join(HERE, "csrc", "_obj", "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
