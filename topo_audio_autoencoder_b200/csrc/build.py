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
STAMP = OUT + ".stamp"
SOURCES = ["tables.cu", "gate.cu", "rectifier.cu", "operators.cu", "aggregate.cu", "combine.cu", "combine_tc.cu", "weight_images.cu", "combine_fwd16.cu", "combine_bwd_tc.cu", "gemm16_debug.cu", "distance.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
         "-I", os.path.join(ROOT, "include"), "-I", HERE]
if os.environ.get("TOPO_DEBUG_KERNELS", "0") not in ("", "0"):
    FLAGS.append("-DTOPO_DEBUG_KERNELS=1")      # ablation knobs + globaltimer stamps in the tensor-core kernels (scripts/ablate_*.py)


def _digest(sources):
    h = hashlib.sha256()
    for f in sorted(os.listdir(HERE)) + [os.path.join(ROOT, "include", "topo_b200.h")]:
        p = f if os.path.isabs(f) else os.path.join(HERE, f)
        if p.endswith((".cu", ".cuh", ".h")):
            h.update(open(p, "rb").read())
    h.update(" ".join(FLAGS + sources).encode())
    return h.hexdigest()


def build(force=False, verbose=False, sources=None):
    sources = [s for s in (sources or SOURCES) if os.path.exists(os.path.join(HERE, s))]
    digest = _digest(sources)
    if not force and os.path.exists(OUT) and os.path.exists(STAMP) and open(STAMP).read() == digest:
        return OUT
    objs = []
    procs = []
    for s in sources:
        o = os.path.join(HERE, s.replace(".cu", ".o"))
        cmd = ["nvcc", *FLAGS, "-c", os.path.join(HERE, s), "-o", o]
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
    subprocess.check_call(["nvcc", "-shared", "-o", OUT, *objs])   # cudart is linked statically
    with open(STAMP, "w") as f:
        f.write(digest)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
