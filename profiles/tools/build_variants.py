"""A/B builds of libspt.so: the same sources with extra nvcc flags per translation unit, written to variants/<tag>/libspt.so
(git-ignored, shipped to the GPU box). Select one with SPT_LIB=variants/<tag>/libspt.so.
    python profiles/tools/build_variants.py tag1:unit.cu:-DFLAG=1,-DOTHER=2 tag2:..."""
import os, subprocess, sys, shutil
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pbrt_v2_spectral_b200 import build as B

def one(spec):
    tag, unit, flags = spec.split(":", 2)
    flags = [f for f in flags.split(",") if f]
    out = os.path.join(ROOT, "variants", tag)
    os.makedirs(out, exist_ok=True)
    B.build()
    objs = []
    for u, uflags in B.UNITS:
        o = os.path.join(B.OBJ, u[:-3] + ".o")
        if u == unit:
            o = os.path.join(out, u[:-3] + ".o")
            B._run(["nvcc", *B.COMMON, *uflags, *flags, "-c", os.path.join(B.CSRC, u), "-o", o], False)
        objs.append(o)
    B._run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", os.path.join(out, "libspt.so"), *objs], False)
    print("built", tag, flush=True)

if __name__ == "__main__":
    B.build()
    with ThreadPoolExecutor(4) as ex:
        list(ex.map(one, sys.argv[1:]))
