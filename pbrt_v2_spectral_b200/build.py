"""Builds libspt.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

-fmad=false: the reference is built without FMA contraction (src/Makefile:24-29) and
hit/miss decisions must match it bit for bit; IEEE division / square root are nvcc defaults."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "spt_api.cu")
OUT = os.path.join(HERE, "libspt.so")
DEPS = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))] + [os.path.join(ROOT, "include", "spt.h")]


def nvcc_cmd(extra=()):
    return ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false",
            "-std=c++17", "--shared", "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"),
            "-I" + os.path.join(HERE, "csrc"), *extra, "-o", OUT, SRC]


def build(force=False, verbose=False):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    cmd = nvcc_cmd(("-Xptxas", "-v") if verbose else ())
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    return OUT


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(OUT)
