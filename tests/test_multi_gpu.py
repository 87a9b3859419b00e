"""-m gpu: several GPUs behind the C ABI (include/spt.h, "several GPUs"): the film kernel of every GPU adds straight into ONE
film over peer access. Runs on a one-GPU box too: spt_multi_* with one device, and the inter-process film (two processes, one
GPU) exercise the same code; the two-device cases are skipped there."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as O
from pbrt_v2_spectral_b200 import capi, ctypes_defs as D, multi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tiny():
    lowered, _ = O.load_case(*O.golden_cases(big=False)[0][1:])
    rp = D.SptRenderParams.from_buffer_copy(bytes(lowered.params))
    rp.spp = 8
    rp.seed = 11
    rp.tile_size = 8
    return lowered, rp


def _single(lowered, rp):
    scene = capi.Scene(lowered)
    film = capi.Film(lowered.film)
    scene.render(film, rp)
    c, w = film.download()
    film.close(); scene.close()
    return c, w


@pytest.mark.parametrize("ndev", [1, 2, 4])
def test_multi_render_equals_single_render(ndev):
    """spt_multi_render on N devices of this process = spt_render on one: all samples of a pixel are rendered by one GPU
    and summed in one warp before the film sees them, so the images agree to the last bit whatever N is."""
    if capi.device_count() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    lowered, rp = _tiny()
    c1, w1 = _single(lowered, rp)
    m = capi.MultiRenderer(lowered, devices=list(range(ndev)))
    assert m.device_count == ndev
    m.render(rp)
    c, w = m.film.download()
    total = sum(m.stats(k)["camera_samples"] for k in range(ndev))
    m.film.clear()
    m.render(rp)                        # a second frame into the cleared film
    c2, w2 = m.film.download()
    m.close()
    assert total == (rp.x_end - rp.x_start) * (rp.y_end - rp.y_start) * rp.spp
    assert np.array_equal(w, w1) and np.array_equal(w2, w1)
    assert np.allclose(c, c1, rtol=1e-5, atol=1e-6) and np.allclose(c2, c1, rtol=1e-5, atol=1e-6)


def test_more_ranks_than_tiles():
    """A rank that owns no tile (SIGFPE in the wave sizing) renders nothing and returns SPT_OK."""
    lowered, rp = _tiny()
    rp.tile_size = 32                   # 48x48 sample extent -> 2x2 tiles
    scene = capi.Scene(lowered)
    film = capi.Film(lowered.film)
    for r in range(7):
        rp.tile_rank, rp.tile_nranks = r, 7
        scene.render(film, rp)          # ranks 4..6 own nothing
    c, w = film.download()
    film.close(); scene.close()
    rp.tile_rank, rp.tile_nranks = 0, 1
    c1, w1 = _single(lowered, rp)
    assert np.array_equal(w, w1)
    assert np.allclose(c, c1, rtol=1e-5, atol=1e-6)


CHILD = r'''
import sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
import oracle_lib as O
from pbrt_v2_spectral_b200 import capi, ctypes_defs as D
lowered, _ = O.load_case(*O.golden_cases(big=False)[0][1:])
rp = D.SptRenderParams.from_buffer_copy(bytes(lowered.params))
rp.spp = 8; rp.seed = 11; rp.tile_size = 8; rp.tile_rank = 1; rp.tile_nranks = 2
capi.set_device(%(dev)d)
film = capi.Film(lowered.film, ipc_handle=bytes.fromhex(sys.argv[1]))
scene = capi.Scene(lowered)
scene.render(film, rp)
film.close(); scene.close()
print("child done")
'''


def test_film_shared_between_processes():
    """One process per GPU (the torchrun layout): rank 0 exports its film, another PROCESS opens it and renders its tile set
    into it (spt_film_ipc_export / spt_film_open_ipc); the film then holds the whole image."""
    lowered, rp = _tiny()
    c1, w1 = _single(lowered, rp)
    scene = capi.Scene(lowered)
    film = capi.Film(lowered.film)
    handle = film.ipc_export()
    dev = 1 if capi.device_count() > 1 else 0
    code = CHILD % {"root": ROOT, "tests": os.path.join(ROOT, "tests"), "dev": dev}
    child = subprocess.Popen([sys.executable, "-c", code, handle.hex()], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    rp.tile_rank, rp.tile_nranks = 0, 2
    scene.render(film, rp)
    out, err = child.communicate(timeout=300)
    assert child.returncode == 0 and "child done" in out, err
    c, w = film.download()
    film.close(); scene.close()
    assert np.array_equal(w, w1)
    assert np.allclose(c, c1, rtol=1e-5, atol=1e-6)
