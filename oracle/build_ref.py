"""Builds oracle/_ref/: the reference's OWN sources, compiled unmodified from where they lie under /root/reference.
TEST INFRASTRUCTURE ONLY (checker and reported baseline; never on the product path).

    python oracle/build_ref.py            # called by __graft_entry__.build() when /root/reference exists

Nothing of the reference is copied into this repository: the compiler reads server/src/server.cpp, utils.cpp and
kernels.cu in place; only the binaries land in oracle/_ref/ (git-ignored, but they travel to the GPU box).  The
capture/show threads of the reference (server/src/threads.cpp: OpenCV + V4L2 webcam + sockets, not buildable here) are
replaced by the file-fed tests/host/threads_stub.cpp.

  ref_server_cpu      server.cpp with -DCPU: the reference's CPU filter chain, server.cpp:96-135 (average gray in place
  ref_server_cpu_O0   -> histogram -> two-max threshold -> binarize).  common.h:15 hard-defines GPU, which would select
                      both branches at once, so its include guard is pre-defined (-DCOMM_H_) and the few macros
                      server.cpp needs (K, LR_THRESHOLDS, CHARS_STR) are given on the command line with common.h's
                      values.  _O0 = no optimisation flag, as server/Makefile:13 builds it.
  ref_server_dropin   server.cpp as shipped (GPU branch, server.cpp:53,139) compiled against the reference's own
                      kernels.cuh and LINKED AGAINST libcvs_b200.so: the drop-in boundary exercised by the unmodified
                      caller.
  ref_server_refgpu   server.cpp + the reference's own server/src/kernels.cu built for sm_100a: what the reference's
                      GPU code does on a B200 (one 1024-thread block; it reads and writes 5,120 bytes past its frame
                      buffers and emits the payload in atomicInc order -- a baseline and a set-wise cross-check, not a
                      bit-exact oracle).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")
SRC = os.path.join(REF, "server", "src")
STUB = os.path.join(ROOT, "tests", "host", "threads_stub.cpp")
LIBDIR = os.path.join(ROOT, "cudavideostream_b200")

# server/include/common.h:6,13,14
COMMON = ["-DCOMM_H_", "-DK=3", "-DLR_THRESHOLDS=20", '-DCHARS_STR="0123456789BFPSWbkps :/"']


def available() -> bool:
    return os.path.exists(os.path.join(SRC, "server.cpp"))


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> dict:
    """Returns {name: path} of what was built (empty when /root/reference is absent)."""
    if not available():
        return {}
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(SRC, "server.cpp"), os.path.join(SRC, "utils.cpp"), STUB]
    inc = ["-I", SRC]  # "../include/x.hpp" in the stub resolves to the reference's server/include/
    built = {}

    def run(cmd, target):
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        built[os.path.basename(target)] = target

    gxx = shutil.which("g++") or "g++"
    for name, opt in (("ref_server_cpu", ["-O2"]), ("ref_server_cpu_O0", [])):
        tgt = os.path.join(OUT, name)
        if force or _stale(tgt, srcs + [__file__]):
            run([gxx, "-std=c++11", "-w"] + opt + COMMON + ["-DCPU", "-DSTUB_CPU"] + inc + srcs + ["-o", tgt, "-lpthread"], tgt)
        else:
            built[name] = tgt

    lib = os.path.join(LIBDIR, "libcvs_b200.so")
    tgt = os.path.join(OUT, "ref_server_dropin")
    if os.path.exists(lib) and (force or _stale(tgt, srcs + [lib, __file__])):
        run([gxx, "-std=c++11", "-w", "-O2"] + inc + srcs + ["-o", tgt, "-L", LIBDIR, "-l:libcvs_b200.so",
                                                              "-Wl,-rpath,$ORIGIN/../../cudavideostream_b200", "-lpthread"], tgt)
    elif os.path.exists(tgt):
        built["ref_server_dropin"] = tgt

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    tgt = os.path.join(OUT, "ref_server_refgpu")
    kern = os.path.join(SRC, "kernels.cu")
    if os.path.exists(nvcc) and (force or _stale(tgt, srcs + [kern, __file__])):
        run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-w", "-std=c++14"] + inc + srcs + [kern, "-o", tgt], tgt)
    elif os.path.exists(tgt):
        built["ref_server_refgpu"] = tgt
    return built


if __name__ == "__main__":
    b = build(force="--force" in sys.argv, verbose=True)
    print(b if b else "reference sources not found under /root/reference: nothing built")
