"""Loader for the plain-C oracle (oracle/spt_oracle.c). TEST INFRASTRUCTURE: only tests/, smoke()
and bench.py's cpu_baseline leg may import this."""
import ctypes as C
import os
import subprocess

import numpy as np

from pbrt_v2_spectral_b200 import ctypes_defs as D
from pbrt_v2_spectral_b200.scene_io import LoweredScene, load_container

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_ref", "liboracle.so" if D.NBANDS == 32 else "liboracle%d.so" % D.NBANDS)
GOLDEN_SMALL = os.path.join(ROOT, "tests", "golden")
GOLDEN_BIG = os.path.join(ROOT, "oracle", "_ref", "golden")

_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(ROOT, "oracle", "spt_oracle.c")
        if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "restate"], check=True)
        _lib = C.CDLL(ORACLE_SO)
        _lib.orc_nbands.restype = C.c_int
        assert _lib.orc_nbands() == D.NBANDS
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def camera_rays(scene, samples5):
    s = np.ascontiguousarray(samples5, np.float32)
    out = np.empty((len(s), 8), np.float32)
    lib().orc_camera_rays(C.byref(scene.camera), _p(s), C.c_uint64(len(s)), _p(out))
    return out


def trace_closest(scene, rays):
    r = np.ascontiguousarray(rays, np.float32)
    n = len(r)
    slot = np.empty(n, np.uint32); pid = np.empty(n, np.uint32); t = np.empty(n, np.float32)
    lib().orc_trace_closest(C.byref(scene.desc), _p(r), C.c_uint64(n), _p(slot), _p(pid), _p(t))
    return slot, pid, t


def trace_any(scene, rays):
    r = np.ascontiguousarray(rays, np.float32)
    hit = np.empty(len(r), np.uint8)
    lib().orc_trace_any(C.byref(scene.desc), _p(r), C.c_uint64(len(r)), _p(hit))
    return hit


def trace_counts(scene, rays):
    r = np.ascontiguousarray(rays, np.float32)
    nodes = C.c_uint64(0); prims = C.c_uint64(0)
    lib().orc_trace_closest_counted(C.byref(scene.desc), _p(r), C.c_uint64(len(r)), C.byref(nodes), C.byref(prims))
    return nodes.value, prims.value


def sample_floats(scene, integrator=None):
    integ = scene.params.integrator if integrator is None else integrator
    return int(lib().orc_sample_floats(C.byref(scene.desc), C.c_int32(integ)))


def shade_samples(scene, samples37, rng, max_depth=None, spp=None, integrator=None):
    s = np.ascontiguousarray(samples37, np.float32)
    integ = scene.params.integrator if integrator is None else integrator
    assert s.shape[1] == sample_floats(scene, integ), "sample vectors do not have the integrator's layout"
    g = np.ascontiguousarray(rng, np.float32)
    n = len(s)
    out = np.empty((n, D.NBANDS), np.float32)
    md = scene.params.max_depth if max_depth is None else max_depth
    lib().orc_shade_samples(C.byref(scene.desc), C.byref(scene.camera), C.c_int32(integ), C.c_int32(md),
                            C.c_int32(scene.params.spp if spp is None else spp), _p(s), _p(g),
                            C.c_int32(g.shape[1]), C.c_uint64(n), _p(out))
    return out


def ulp_sensitivity(scene, samples37, rng, tol=2e-4, seed=1):
    """Fraction of samples whose ORACLE radiance moves by more than `tol` (relative) when every sample value
    that feeds direction sampling is moved by one ulp: how ill-conditioned the reference's own formulas are
    on this scene (Blinn exponent 1000 under an HDR environment map: ~0.5 %; diffuse scenes: ~0.05 %). The
    CUDA path differs from glibc by an ulp or two in every sinf/cosf/powf/atan2f/acosf along a path, so its
    agreement with the reference cannot be better than a small multiple of this."""
    s = np.ascontiguousarray(samples37, np.float32)
    ref = shade_samples(scene, s, rng)
    p = s.copy()
    up = np.random.default_rng(seed).integers(0, 2, size=p[:, 5:].shape).astype(bool)
    p[:, 5:] = np.where(up, np.nextafter(p[:, 5:], np.float32(2)), np.nextafter(p[:, 5:], np.float32(-1))).astype(np.float32)
    L = shade_samples(scene, p, rng)
    err = np.abs(L - ref).max(axis=1) / np.maximum(np.abs(ref).max(axis=1), 1e-6)
    return float((err > tol).mean())


def film_add_samples(scene, xy, L, film=None):
    fd = film if film is not None else scene.film
    c = np.zeros((fd.y_pixel_count, fd.x_pixel_count, D.NBANDS), np.float32)
    w = np.zeros((fd.y_pixel_count, fd.x_pixel_count), np.float32)
    xy = np.ascontiguousarray(xy, np.float32); L = np.ascontiguousarray(L, np.float32)
    lib().orc_film_add_samples(C.byref(fd), C.byref(scene.tables), _p(xy), _p(L), C.c_uint64(len(xy)), _p(c), _p(w))
    return c, w


def gen_samples(seed, px, py, spp, shutter=(0.0, 1.0), n_rng=34):
    smp = np.empty((spp, 37), np.float32); rng = np.empty((spp, n_rng), np.float32)
    L = lib()
    for s in range(spp):
        L.orc_gen_sample(C.c_uint64(seed), C.c_int32(px), C.c_int32(py), C.c_int32(s), C.c_int32(spp),
                         C.c_float(shutter[0]), C.c_float(shutter[1]), C.c_int32(n_rng),
                         _p(smp[s]), _p(rng[s]))
    return smp, rng


def render(scene, params=None, film=None):
    fd = film if film is not None else scene.film
    rp = params if params is not None else scene.params
    c = np.zeros((fd.y_pixel_count, fd.x_pixel_count, D.NBANDS), np.float32)
    w = np.zeros((fd.y_pixel_count, fd.x_pixel_count), np.float32)
    lib().orc_render(C.byref(scene.desc), C.byref(scene.camera), C.byref(fd), C.byref(rp), _p(c), _p(w))
    return c, w


def golden_cases(big=True):
    """[(name, scene path, golden path)] of every golden set present (the tiny one always is)."""
    if D.NBANDS != 32:
        # the 30-band variant (SPT_NBANDS=30 in the environment): its own golden sets, made by the reference built with
        # nSpectralSamples = 30 (oracle/Makefile ref30); names end in the band count
        return [(f[:-7], os.path.join(GOLDEN_BIG, f[:-7] + ".spt"), os.path.join(GOLDEN_BIG, f))
                for f in (sorted(os.listdir(GOLDEN_BIG)) if os.path.isdir(GOLDEN_BIG) else []) if f.endswith("_small%d.golden" % D.NBANDS)]
    cases = [(n, os.path.join(GOLDEN_SMALL, n + ".spt"), os.path.join(GOLDEN_SMALL, n + ".golden"))
             for n in ("tiny", "tiny_tex", "tiny_direct") if os.path.exists(os.path.join(GOLDEN_SMALL, n + ".golden"))]
    if big and os.path.isdir(GOLDEN_BIG):
        for f in sorted(os.listdir(GOLDEN_BIG)):
            if f.endswith(".golden") and not f.endswith("_rays.golden") and not f[:-7].endswith("30"):
                name = f[:-7]
                cases.append((name, os.path.join(GOLDEN_BIG, name + ".spt"), os.path.join(GOLDEN_BIG, f)))
    return cases


def ray_cases(big=True):
    """[(name, scene path, golden path)] of the compact first-hit sets ({image_xy, prim_id, t_hit}, oracle_dump's
    SPT_DUMP_COMPACT): the committed mid-size ones (65 536 camera rays of BASELINE configs 1 and 2 at full resolution, xz
    containers) and, where oracle/_ref/golden is present, the 1 048 576-ray ones."""
    if D.NBANDS != 32:
        return []
    cases = [(n, os.path.join(GOLDEN_SMALL, n + ".spt.xz"), os.path.join(GOLDEN_SMALL, n + ".golden.xz"))
             for n in ("killeroo_rays_mid", "bunny_rays_mid") if os.path.exists(os.path.join(GOLDEN_SMALL, n + ".golden.xz"))]
    if big:
        cases += [(n, os.path.join(GOLDEN_BIG, n + ".spt"), os.path.join(GOLDEN_BIG, n + ".golden"))
                  for n in ("killeroo_rays", "bunny_rays") if os.path.exists(os.path.join(GOLDEN_BIG, n + ".golden"))]
    return cases


def compact_samples(g):
    """n x 5 camera sample floats of a compact set: {imageX, imageY, 0, 0, 0} (pinhole camera, static scene)."""
    s = np.zeros((len(g["image_xy"]), 5), np.float32)
    s[:, :2] = g["image_xy"]
    return s


def load_case(scene_path, golden_path):
    return LoweredScene.load(scene_path), load_container(golden_path)
