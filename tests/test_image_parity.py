"""-m gpu: converged-image parity against .dat files rendered by the unmodified reference binary
(oracle/make_golden.py --images). north_star: "within 1 % mean relative error per spectral band";
as SURVEY.md 8c measured, the per-pixel-averaged form cannot pass even reference-vs-reference
(3 % Monte-Carlo noise at 1024 spp), so the aggregate form is used: per band
    sum_pixels |gpu - ref| / sum_pixels ref  <= 1 %      (reference-vs-reference floor: 0.15 %)
plus the band-mean bias |mean(gpu) - mean(ref)| / mean(ref) <= 0.3 %, the sensitive detector of
systematic shading errors. Both films hold un-normalised sums: divided by spp here."""
import json
import os

import numpy as np
import pytest

import oracle_lib as O
from pbrt_v2_spectral_b200 import capi, ctypes_defs as D
from pbrt_v2_spectral_b200.scene_io import LoweredScene

pytestmark = pytest.mark.gpu

IMAGES = [("killeroo_small", 1024), ("bunny_small", 4096), ("metal_small", 512), ("envmap_small", 8192), ("synth_small", 2048),
          ("ssenv_small", 4096), ("specular_small", 2048),
          # configs 3 and 4 with their shipped floor: substrate + image-mapped Kd (EWA) + bump map
          ("metal_shipped_small", 8192), ("ssenv_shipped_small", 8192),
          # the shipped scenes under their own integrator: directlighting, strategy all
          ("killeroo_direct_small", 1024), ("bunny_direct_small", 1024),
          # the bunny's shipped measured BRDF: under the path integrator (config 2), and bunny.pbrt exactly as shipped
          ("bunny_measured_small", 4096), ("bunny_shipped_small", 1024),
          # the half-angle (MERL-format) measured BRDF
          ("tiny_merl_small", 32768),
          # directlighting with a glass and a mirror killeroo: the SpecularReflect / SpecularTransmit tree (directlighting.cpp:97-107)
          ("specular_direct_small", 1024)]
if D.NBANDS == 30:      # the 30-band library (SPT_NBANDS=30): the image the reference built with nSpectralSamples = 30 rendered
    IMAGES = [("killeroo_small30", 1024)]


FLOOR_PATH = os.path.join(O.GOLDEN_SMALL, "image_floor.json")
FLOOR = json.load(open(FLOOR_PATH)) if os.path.exists(FLOOR_PATH) else {}
N_SEEDS_HEAVY = 8


def _render(scene, lowered, spp, seed):
    rp = D.SptRenderParams.from_buffer_copy(bytes(lowered.params))
    rp.spp = spp
    rp.seed = seed
    film = capi.Film(lowered.film)
    scene.render(film, rp)
    c, w = film.download()
    st = scene.stats()
    film.close()
    return c.astype(np.float64) / spp, st


@pytest.mark.parametrize("name,spp", IMAGES, ids=[n for n, _ in IMAGES])
def test_converged_image_within_1_percent_per_band(name, spp):
    dat = os.path.join(O.GOLDEN_BIG, "%s_%dspp.ref.npy" % (name, spp))    # the reference's .dat as float32 [y][x][band]
    spt = os.path.join(O.GOLDEN_BIG, name + ".spt")
    if not (os.path.exists(dat) and os.path.exists(spt)):
        pytest.skip("reference image %s not generated (oracle/make_golden.py --images)" % dat)
    ref = np.load(dat).astype(np.float64) / spp
    lowered = LoweredScene.load(spt)
    scene = capi.Scene(lowered)
    gpu_spp = max(4 * spp, 4096)            # the GPU side's own noise is pushed below the reference's
    if lowered.desc.n_textures:
        # image textures are filtered with ray differentials scaled by 1/sqrt(spp) (samplerrenderer.cpp:91) and the bump
        # map's finite-difference step follows them (material.cpp:48-60): the reference's image itself depends on spp
        gpu_spp = spp
    l1_allowed, bias_allowed = 0.01, 0.003
    if name in FLOOR:
        # Heavy-tailed scene (Blinn exponent 1000 under an HDR map; a glossy measured BRDF that is only cosine-sampled): the
        # UNMODIFIED REFERENCE does not reproduce its own image to 1 % at this sample count. tests/golden/image_floor.json
        # holds the per-band L1 / bias between two independent reference renders (oracle/make_golden.py --floor: another
        # --ncores = another task split = other RNG streams). The comparison is made as sharp as the data allow - the mean of
        # the two reference renders against the mean of N_SEEDS_HEAVY independent GPU renders at the reference's spp - and must
        # then be within max(1 %, 1.2 x floor) / max(0.3 %, 1.2 x floor bias): no other criterion is substituted.
        fl = FLOOR[name]
        assert fl["spp"] == spp
        ref2 = np.load(os.path.join(O.GOLDEN_BIG, "%s_%dspp.ref2.npy" % (name, spp))).astype(np.float64) / spp
        ref = 0.5 * (ref + ref2)
        gpu_spp = spp
        img = np.zeros_like(ref)
        for k in range(N_SEEDS_HEAVY):
            one, st = _render(scene, lowered, gpu_spp, 2024 + k)          # 2024 = oracle/make_golden.py IMAGE_SEED
            img += one / N_SEEDS_HEAVY
            if k == 0:
                first = one
        l1_allowed, bias_allowed = max(l1_allowed, 1.2 * fl["l1_max"]), max(bias_allowed, 1.2 * fl["bias_max"])
        print("%s: reference vs reference at %d spp: L1 %.3f%%, bias %.3f%% -> allowed %.3f%% / %.3f%%" % (
            name, spp, 100 * fl["l1_max"], 100 * fl["bias_max"], 100 * l1_allowed, 100 * bias_allowed))
        oracle_img = os.path.join(O.GOLDEN_BIG, "%s_%dspp.oracle.npy" % (name, spp))
        if os.path.exists(oracle_img):
            # on top: the CPU oracle (bit-identical to the reference per sample) rendered this frame with the product's sampler
            # and seed 2024; the GPU image of the same samples must reproduce it
            oimg = np.load(oracle_img).astype(np.float64)
            l1o = np.abs(first - oimg).sum((0, 1)) / oimg.sum((0, 1))
            print("%s: against the oracle's render of the same samples: L1 max %.3f%%" % (name, 100 * l1o.max()))
            assert l1o.max() <= 0.003, l1o
    else:
        img, st = _render(scene, lowered, gpu_spp, 2024)
    scene.close()
    print("%s: %d spp rendered in %.1f ms (%.0f Msamples/s)" % (name, gpu_spp, st["render_ms"], st["camera_samples"] / st["render_ms"] / 1e3))
    assert img.shape == ref.shape
    l1 = np.abs(img - ref).sum((0, 1)) / ref.sum((0, 1))
    bias = np.abs(img.mean((0, 1)) - ref.mean((0, 1))) / ref.mean((0, 1))
    print("%s: per-band aggregate L1 error max %.3f%%, band-mean bias max %.3f%%" % (name, 100 * l1.max(), 100 * bias.max()))
    assert l1.max() <= l1_allowed, l1
    assert bias.max() <= bias_allowed, bias
