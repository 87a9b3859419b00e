"""The drop-in boundary end to end: the reference's own binary with the 3-line `Renderer "gpupath"` registration
(oracle/_ref/bin/pbrt_gpupath = the unmodified reference objects + pbrt_v2_spectral_b200/host/{gpupath,lowering}.cpp)
parses a .pbrt file, builds the scene with the reference's classes, lowers it and renders through libspt.so.
 * a scene the GPU path does not implement is handed to the reference's SamplerRenderer, with a reason (no GPU needed);
 * a scene it implements gives the film spt_render gives through the Python binding (GPU)."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from pbrt_v2_spectral_b200 import capi, ctypes_defs as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "bin")
TINY = os.path.join(ROOT, "tests", "golden", "tiny.pbrt")
LIB = os.path.join(ROOT, "pbrt_v2_spectral_b200", "libspt.so")

needs_bins = pytest.mark.skipif(not (os.path.exists(os.path.join(BIN, "pbrt_gpupath")) and os.path.exists(os.path.join(BIN, "pbrt"))),
                                reason="oracle/_ref/bin not built (build() needs /root/reference)")


def _run(binary, scene_text, cwd, name):
    with open(os.path.join(cwd, name + ".pbrt"), "w") as f:
        f.write(scene_text.replace('"tiny.exr"', '"%s.exr"' % name))
    env = dict(os.environ, SPT_LIB=LIB)
    return subprocess.run([os.path.join(BIN, binary), "--quiet", "--ncores", "1", name + ".pbrt"], cwd=cwd, env=env,
                          capture_output=True, text=True, timeout=600)


@needs_bins
def test_unsupported_scene_falls_back_to_the_reference_renderer(tmp_path):
    s = open(TINY).read()
    # a procedural texture is not lowered: the whole scene must be rendered by the reference's own SamplerRenderer
    s = s.replace('Material "matte" "color Kd" [.55 .5 .45]',
                  'Texture "chk" "color" "checkerboard" "float uscale" [4] "float vscale" [4]\nMaterial "matte" "texture Kd" "chk"')
    assert '"chk"' in s
    cwd = str(tmp_path)
    ref = _run("pbrt", s, cwd, "ref")
    assert ref.returncode == 0, ref.stderr
    gpu = _run("pbrt_gpupath", s.replace("WorldBegin", 'Renderer "gpupath"\nWorldBegin', 1), cwd, "gpu")
    assert gpu.returncode == 0, gpu.stderr
    log = re.sub(r"\s+", " ", gpu.stderr + gpu.stdout)          # the reference's Error() wraps long lines
    assert "rendering with the CPU SamplerRenderer instead" in log
    assert "texture for Kd is neither constant nor an image map" in log
    a, b = capi.read_dat(os.path.join(cwd, "ref.dat")), capi.read_dat(os.path.join(cwd, "gpu.dat"))
    assert np.array_equal(a, b)          # same renderer, same task decomposition (--ncores 1): the same image


@needs_bins
@pytest.mark.gpu
@pytest.mark.parametrize("gpus", [None, 1, 2, 0])
def test_gpupath_binary_renders_what_spt_render_renders(tmp_path, gpus):
    """`Renderer "gpupath"` alone renders on one GPU through spt_render; `"integer gpus" [N]` (0 = every visible GPU) goes
    through spt_multi_* - the same image either way. The parameters are read like any renderer's (src/core/api.cpp:1374-1379)
    and raise no unused-parameter warning."""
    if gpus == 2 and capi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    extra = "" if gpus is None else ' "integer gpus" [%d]' % gpus
    s = open(TINY).read().replace("WorldBegin", 'Renderer "gpupath" "integer seed" [5]%s\nWorldBegin' % extra, 1)
    cwd = str(tmp_path)
    r = _run("pbrt_gpupath", s, cwd, "drop")
    assert r.returncode == 0, r.stderr
    log = re.sub(r"\s+", " ", r.stderr + r.stdout)
    assert "CPU SamplerRenderer instead" not in log
    assert "not used" not in log and "unused" not in log.lower(), log
    got = capi.read_dat(os.path.join(cwd, "drop.dat"))
    lowered, _ = O.load_case(*O.golden_cases(big=False)[0][1:])
    scene = capi.Scene(lowered)
    rp = D.SptRenderParams.from_buffer_copy(bytes(lowered.params))
    rp.seed = 5
    film = capi.Film(lowered.film)
    scene.render(film, rp)
    c, _ = film.download()
    film.close(); scene.close()
    want = np.maximum(c.astype(np.float64), 0.0)     # WriteImage clamps at zero (spectralImage.cpp:283-296)
    assert got.shape == want.shape
    assert np.allclose(got, want, rtol=1e-5, atol=1e-7)   # same kernels, same samples; film atomics may add in another order
