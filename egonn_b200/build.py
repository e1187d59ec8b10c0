"""Build libegonn_b200.so in-tree with nvcc for sm_100a (no torch dependency in the library)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libegonn_b200.so")
SOURCES = ["api.cu", "coords.cu", "ops.cu", "forward.cu", "sconv_tc.cu", "sconv_ts.cu", "conv0_tc.cu", "comm.cu", "sort.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _newest_source_mtime() -> float:
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    paths.append(os.path.join(os.path.dirname(HERE), "include", "egonn_b200.h"))
    return max(os.path.getmtime(p) for p in paths)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_source_mtime():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = list(NVCC_FLAGS)
    if os.environ.get("EGN_TRACE_BUILD") == "1":           # clock64 stamps inside the tensor-core convolution (tools/trace_conv.py)
        flags.append("-DEGN_TRACE_BUILD")
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"== {src} ==\n{out}")
        failed |= p.returncode != 0
    with open(os.path.join(CSRC, "build.log"), "w") as f:
        f.write("\n".join(log))
    if failed or verbose:
        print("\n".join(log), file=sys.stderr)
    if failed:
        raise RuntimeError("nvcc failed; see egonn_b200/csrc/build.log")
    subprocess.check_call([nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-ldl"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
