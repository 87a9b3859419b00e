"""-m gpu: argument checking at the C ABI (a scene needs a device to exist, hence the marker): what the reference would
refuse or never produce is refused with a status and a message, nothing is rendered approximately."""
import os

import numpy as np
import pytest

import oracle_lib as O
from pbrt_v2_spectral_b200 import capi, ctypes_defs as D

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tiny():
    lowered, g = O.load_case(*O.golden_cases(big=False)[0][1:])
    scene = capi.Scene(lowered)
    film = capi.Film(lowered.film)
    yield lowered, scene, film
    film.close(); scene.close()


def _params(lowered, **kw):
    rp = D.SptRenderParams.from_buffer_copy(bytes(lowered.params))
    for k, v in kw.items():
        setattr(rp, k, v)
    return rp


def test_spp_must_be_a_power_of_two(tiny):
    lowered, scene, film = tiny
    with pytest.raises(capi.SptError, match="power of two"):       # LDSampler rounds up (lowdiscrepancy.cpp:38-46): the host passes the rounded count
        scene.render(film, _params(lowered, spp=3))


def test_unknown_integrator_is_refused(tiny):
    lowered, scene, film = tiny
    with pytest.raises(capi.SptError, match="unknown integrator"):
        scene.render(film, _params(lowered, integrator=7))


def test_tile_rank_out_of_range(tiny):
    lowered, scene, film = tiny
    with pytest.raises(capi.SptError, match="tile_rank"):
        scene.render(film, _params(lowered, tile_rank=2, tile_nranks=2))


def test_directlighting_one_with_specular_materials_is_unsupported():
    sp, gp = os.path.join(O.GOLDEN_BIG, "specular_small.spt"), os.path.join(O.GOLDEN_BIG, "specular_small.golden")
    if not os.path.exists(sp):
        pytest.skip("specular_small golden set not generated")
    lowered, _ = O.load_case(sp, gp)
    scene = capi.Scene(lowered)
    film = capi.Film(lowered.film)
    try:
        # strategy "all" walks the SpecularReflect / SpecularTransmit tree (directlighting.cpp:97-103) on the device; strategy "one"
        # with specular materials is not lowered, so refused
        with pytest.raises(capi.SptError, match="specular"):
            scene.render(film, _params(lowered, integrator=D.INTEGRATOR_DIRECT_ONE, spp=1))
        c, w = film.download()
        assert not np.any(w) and not np.any(c)                      # nothing was rendered
    finally:
        film.close(); scene.close()
