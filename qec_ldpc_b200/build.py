"""Builds the product library qec_ldpc_b200/lib/libqldpc_b200.so (C ABI of include/qldpc_b200.h) with nvcc for
sm_100a, in-tree, plus the C++ command-line driver.  nvcc cross-compiles without a GPU."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libqldpc_b200.so")
CLI = os.path.join(LIBDIR, "qec_ldpc")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC"]
LIB_SOURCES = ["code.cpp", "host_pack.cpp", "kernels.cu", "decoder.cu", "bp_global.cu"] + sorted(
    f for f in os.listdir(CSRC) if f.startswith("bp_shape_") and f.endswith(".cu"))


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _env():
    env = dict(os.environ)
    # this image exports CC/CXX pointing at a gcc without OpenMP specs; nvcc and g++ come from PATH
    env.pop("CC", None)
    env.pop("CXX", None)
    return env


# what the last build_all() did: {"library": "compiled" | "up to date", "cli": ...} -- __graft_entry__.build() prints it,
# so a log shows whether the nvcc step was exercised or the prebuilt, mtime-fresh binaries were reused
LAST_BUILD = {}


def build_library(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in LIB_SOURCES]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    deps.append(os.path.join(ROOT, "include", "qldpc_b200.h"))
    LAST_BUILD["library"] = "up to date"
    if force or _stale(LIB, deps):
        LAST_BUILD["library"] = "compiled (%d translation units, nvcc sm_100a)" % len(srcs)
        objs = []
        procs = []
        for s in srcs:
            o = os.path.join(LIBDIR, os.path.basename(s) + ".o")
            objs.append(o)
            cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((cmd, subprocess.Popen(cmd, env=_env())))
        for cmd, p in procs:
            if p.wait() != 0:
                raise RuntimeError("nvcc failed: " + " ".join(cmd))
        cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        subprocess.run(cmd, check=True, env=_env())
    return LIB


def build_cli(force=False):
    src = os.path.join(HERE, "cpp", "main.cpp")
    if not os.path.exists(src):
        return None
    deps = [src] + [os.path.join(HERE, "cpp", f) for f in os.listdir(os.path.join(HERE, "cpp"))]
    LAST_BUILD["cli"] = "up to date"
    if force or _stale(CLI, deps + [LIB]):
        LAST_BUILD["cli"] = "compiled"
        cmd = ["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(HERE, "cpp"), src,
               "-o", CLI, "-L", LIBDIR, "-lqldpc_b200", "-Wl,-rpath,$ORIGIN", "-pthread"]
        subprocess.run(cmd, check=True, env=_env())
    return CLI


def build_all(force=False, verbose=False):
    build_library(force, verbose)
    build_cli(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
