"""Build libtopo_b200.so in-tree with nvcc for sm_100a.

    python topo_audio_autoencoder_b200/csrc/build.py [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(os.path.dirname(HERE), "libtopo_b200.so")
# the product sources plus the unit-test GEMMs, compiled with -DTOPO_DEBUG_KERNELS=1: ablation knobs, globaltimer stamps and
# the topo_debug_* entry points (tests/test_gpu_tc.py, scripts/ablate_*.py).  Never loaded by the package itself.
OUT_DEBUG = os.path.join(os.path.dirname(HERE), "libtopo_b200_debug.so")
SOURCES = ["tables.cu", "gate.cu", "rectifier.cu", "operators.cu", "aggregate.cu", "combine.cu", "combine_tc.cu", "weight_images.cu", "combine_fwd16.cu", "combine_bwd_tc.cu", "distance.cu", "attention.cu"]
DEBUG_ONLY_SOURCES = ["gemm16_debug.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
         "-I", os.path.join(ROOT, "include"), "-I", HERE]


def _digest(sources, flags=()):
    h = hashlib.sha256()
    for f in sorted(os.listdir(HERE)) + [os.path.join(ROOT, "include", "topo_b200.h")]:
        p = f if os.path.isabs(f) else os.path.join(HERE, f)
        if p.endswith((".cu", ".cuh", ".h")):
            h.update(open(p, "rb").read())
    h.update(" ".join(list(FLAGS) + list(flags) + list(sources)).encode())
    return h.hexdigest()


def build(force=False, verbose=False, sources=None, target=OUT, extra_flags=(), obj_suffix=".o"):
    sources = [s for s in (sources or SOURCES) if os.path.exists(os.path.join(HERE, s))]
    digest = _digest(sources, extra_flags)
    stamp = target + ".stamp"
    if not force and os.path.exists(target) and os.path.exists(stamp) and open(stamp).read() == digest:
        return target
    objs = []
    procs = []
    for s in sources:
        o = os.path.join(HERE, s.replace(".cu", obj_suffix))
        cmd = ["nvcc", *FLAGS, *extra_flags, "-c", os.path.join(HERE, s), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if out.strip() and (verbose or p.returncode):
            print(f"--- {s} ---\n{out}")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call(["nvcc", "-shared", "-o", target, *objs])   # cudart is linked statically
    with open(stamp, "w") as f:
        f.write(digest)
    return target


def build_debug(force=False, verbose=False):
    """libtopo_b200_debug.so: the test / measurement twin of the product library (see OUT_DEBUG)."""
    return build(force, verbose, SOURCES + DEBUG_ONLY_SOURCES, OUT_DEBUG, ("-DTOPO_DEBUG_KERNELS=1",), ".dbg.o")


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--no-debug" not in sys.argv:
        print(build_debug(force="--force" in sys.argv, verbose="-v" in sys.argv))
