"""Build libka9q_b200.so in-tree with nvcc for sm_100a (no GPU needed to compile).

    python -m ka9q_sdr_b200.build        # or: from ka9q_sdr_b200.build import build; build()
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libka9q_b200.so")
OBJDIR = os.path.join(HERE, "build")

CU_SOURCES = ["bigfft.cu", "chan_kernels.cu", "design.cu", "stream.cu", "mgpu.cu", "n0.cu", "dropin.cu", "decimate.cu", "frontend.cu"]
C_SOURCES = ["osc_host.c", "rtp_glue.c", "rx_host.c"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-fvisibility=default", "--threads", "2"]
# experiment hook: extra -D switches for kernel variants (e.g. KA9Q_B200_NVCC_EXTRA="-DFM_CTAS_PER_SM=6")
NVCC_FLAGS += os.environ.get("KA9Q_B200_NVCC_EXTRA", "").split()


def _stamp(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return r.stdout + r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "ka9q_b200.h"))
    objs = []
    jobs = []
    for src in CU_SOURCES + C_SOURCES:
        sp = os.path.join(CSRC, src)
        obj = os.path.join(OBJDIR, src + ".o")
        stamp_file = obj + ".stamp"
        stamp = _stamp([sp] + headers)
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
            continue
        if src.endswith(".cu"):
            cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", sp, "-o", obj]
        else:
            cmd = ["gcc", "-O2", "-std=gnu11", "-fPIC", "-c", sp, "-o", obj]
        jobs.append((cmd, stamp_file, stamp))
    procs = [(subprocess.Popen(c, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True), c, sf, st) for c, sf, st in jobs]
    for p, c, sf, st in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("build failed: " + " ".join(c))
        if verbose:
            sys.stderr.write(out)
        with open(sf, "w") as f:
            f.write(st)
    if jobs or not os.path.exists(LIB):
        _run(["nvcc", "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-Xlinker",
                                                      "-Bsymbolic", "-lpthread", "-ldl", "-lm"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
