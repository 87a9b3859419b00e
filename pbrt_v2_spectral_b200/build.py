"""Builds libspt.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Three translation units, two floating-point regimes:
  spt_exact.cu   camera rays + BVH traversal, -fmad=false: the reference is built without FMA contraction
                 (src/Makefile:24-29) and hit/miss decisions must match it bit for bit; IEEE
                 division / square root are nvcc defaults.
  spt_shade.cu   shading, accumulation, film. Also -fmad=false: the reference's own formulas are ill-conditioned
                 in places (1 - cos of a small cone angle in the sphere-light pdf), so radiance only tracks the
                 reference to 2e-4 per sample when the rounding sequence is the same; FMA bought 5 % of one kernel.
  spt_api.cu     host side of the C ABI (no kernels).
  spt_build.cu   scene re-layout kernels run once per spt_scene_create (pair nodes, leaf flags, vertex pre-gather).
  spt_wide.cu    host-side builder of the fast (4-wide) traversal layout, no kernels."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

# Two libraries from the same sources: libspt.so (SPT_NBANDS = 32, the reference as shipped) and libspt30.so (-DSPT_NBANDS=30,
# the 30-band SampledSpectrum BASELINE.json's metric names: src/core/spectrum.h:41-43 with nSpectralSamples = 30).
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libspt.so")
OBJ = os.path.join(HERE, "build")
UNITS = [("spt_exact.cu", ["-fmad=false"]), ("spt_shade.cu", ["-fmad=false"]), ("spt_api.cu", ["-fmad=false"]),
         ("spt_build.cu", ["-fmad=false"]), ("spt_wide.cu", [])]
DEPS = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "spt.h")]
COMMON = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
          "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]


def _run(cmd, verbose):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed: " + " ".join(cmd))


def build(force=False, verbose=False, nbands=None):
    """nbands None: both libraries; 32 / 30: that one. Returns the path of the 32-band library (or of the one asked for)."""
    if nbands is None:
        with ThreadPoolExecutor(2) as ex:
            outs = list(ex.map(lambda nb: build(force, verbose, nb), (32, 30)))
        return outs[0]
    out_so = OUT if nbands == 32 else os.path.join(HERE, "libspt%d.so" % nbands)
    obj_dir = OBJ if nbands == 32 else os.path.join(OBJ, "nb%d" % nbands)
    band_flags = [] if nbands == 32 else ["-DSPT_NBANDS=%d" % nbands]
    return _build_one(force, verbose, out_so, obj_dir, band_flags)


def _build_one(force, verbose, OUT, OBJ, band_flags):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    extra = (["-Xptxas", "-v"] if verbose else []) + band_flags
    objs = [os.path.join(OBJ, u[:-3] + ".o") for u, _ in UNITS]
    # a unit is recompiled when its own source or any header is newer than its object
    headers = [d for d in DEPS if not d.endswith(".cu")]
    def stale(u, o):
        return force or not os.path.exists(o) or any(os.path.getmtime(o) < os.path.getmtime(d) for d in headers + [os.path.join(CSRC, u)])
    cmds = [["nvcc", *COMMON, *flags, *extra, "-c", os.path.join(CSRC, u), "-o", o] for (u, flags), o in zip(UNITS, objs) if stale(u, o)]
    with ThreadPoolExecutor(max(len(cmds), 1)) as ex:
        list(ex.map(lambda c: _run(c, verbose), cmds))
    _run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", OUT, *objs], verbose)
    return OUT


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(OUT)
