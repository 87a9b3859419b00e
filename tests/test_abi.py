"""CPU-side checks of the drop-in boundary: libspt.so builds/loads and exports every symbol
include/spt.h declares; struct mirrors agree with the C layout; the container round-trips."""
import ctypes as C
import os
import re
import subprocess

import numpy as np

import oracle_lib as O
from pbrt_v2_spectral_b200 import build, capi, ctypes_defs as D, scene_io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    build.build()
    hdr = open(os.path.join(ROOT, "include", "spt.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(spt_[a-z_]+)\s*\(", hdr)))
    assert declared == sorted(capi.SYMBOLS)
    lib = C.CDLL(capi.LIB_PATH)
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.spt_nbands() == D.NBANDS


def test_struct_sizes_match_c():
    src = r'''
    #include <stdio.h>
    #include "spt.h"
    int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(SptSceneDesc), sizeof(SptCameraDesc),
      sizeof(SptFilmDesc), sizeof(SptRenderParams), sizeof(SptStats), sizeof(SptSpectralTables), sizeof(SptQuadric),
      sizeof(SptXform), sizeof(SptMaterial), sizeof(SptLight), sizeof(SptLightShape)); return 0; }'''
    exe = "/tmp/spt_sizes"
    subprocess.run(["gcc", "-x", "c", "-", "-I" + os.path.join(ROOT, "include"), "-o", exe], input=src.encode(), check=True)
    got = [int(v) for v in subprocess.run([exe], capture_output=True, check=True).stdout.split()]
    want = [C.sizeof(D.SptSceneDesc), C.sizeof(D.SptCameraDesc), C.sizeof(D.SptFilmDesc), C.sizeof(D.SptRenderParams),
            C.sizeof(D.SptStats), C.sizeof(D.SptSpectralTables), D.SIZEOF_QUADRIC, D.SIZEOF_XFORM, D.SIZEOF_MATERIAL,
            D.SIZEOF_LIGHT, D.SIZEOF_LIGHT_SHAPE]
    assert got == want


def test_container_roundtrip(tmp_path):
    a = scene_io.load_container(os.path.join(O.GOLDEN_SMALL, "tiny.spt"))
    p = str(tmp_path / "rt.spt")
    scene_io.save_container(p, a)
    b = scene_io.load_container(p)
    assert a.keys() == b.keys()
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    s = scene_io.LoweredScene(b)
    assert s.n_prims > 300 and s.n_nodes > s.n_prims


def test_no_device_fails_loudly():
    """Without a CUDA device the product must refuse, not fall back (this container has none)."""
    if capi.device_count() > 0:
        return
    s = scene_io.LoweredScene.load(os.path.join(O.GOLDEN_SMALL, "tiny.spt"))
    try:
        capi.Scene(s)
    except capi.SptError as e:
        assert "no CUDA device" in str(e)
    else:
        raise AssertionError("scene creation succeeded without a GPU")
