"""-m gpu: converged-image parity against .dat files rendered by the unmodified reference binary
(oracle/make_golden.py --images). north_star: "within 1 % mean relative error per spectral band";
as SURVEY.md 8c measured, the per-pixel-averaged form cannot pass even reference-vs-reference
(3 % Monte-Carlo noise at 1024 spp), so the aggregate form is used: per band
    sum_pixels |gpu - ref| / sum_pixels ref  <= 1 %      (reference-vs-reference floor: 0.15 %)
plus the band-mean bias |mean(gpu) - mean(ref)| / mean(ref) <= 0.3 %, the sensitive detector of
systematic shading errors. Both films hold un-normalised sums: divided by spp here."""
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
          ("bunny_measured_small", 4096), ("bunny_shipped_small", 1024)]


@pytest.mark.parametrize("name,spp", IMAGES, ids=[n for n, _ in IMAGES])
def test_converged_image_within_1_percent_per_band(name, spp):
    dat = os.path.join(O.GOLDEN_BIG, "%s_%dspp.ref.npy" % (name, spp))    # the reference's .dat as float32 [y][x][band]
    spt = os.path.join(O.GOLDEN_BIG, name + ".spt")
    if not (os.path.exists(dat) and os.path.exists(spt)):
        pytest.skip("reference image %s not generated (oracle/make_golden.py --images)" % dat)
    ref = np.load(dat).astype(np.float64) / spp
    lowered = LoweredScene.load(spt)
    scene = capi.Scene(lowered)
    rp = D.SptRenderParams.from_buffer_copy(bytes(lowered.params))
    gpu_spp = max(4 * spp, 4096)            # the GPU side's own noise is pushed below the reference's
    noise = os.path.join(O.GOLDEN_BIG, "%s_%dspp.noise.json" % (name, spp))
    if os.path.exists(noise):
        gpu_spp = spp                       # heavy-tailed scene: rendered with the very samples of the oracle's fixture (below)
    if lowered.desc.n_textures:
        # image textures are filtered with ray differentials scaled by 1/sqrt(spp) (samplerrenderer.cpp:91) and the bump
        # map's finite-difference step follows them (material.cpp:48-60): the reference's image itself depends on spp
        gpu_spp = spp
    rp.spp = gpu_spp
    rp.seed = 2024                          # = oracle/make_golden.py IMAGE_SEED
    film = capi.Film(lowered.film)
    scene.render(film, rp)
    c, w = film.download()
    st = scene.stats()
    film.close(); scene.close()
    print("%s: %d spp rendered in %.1f ms (%.0f Msamples/s)" % (name, gpu_spp, st["render_ms"], st["camera_samples"] / st["render_ms"] / 1e3))
    img = c.astype(np.float64) / gpu_spp
    assert img.shape == ref.shape
    l1 = np.abs(img - ref).sum((0, 1)) / ref.sum((0, 1))
    bias = np.abs(img.mean((0, 1)) - ref.mean((0, 1))) / ref.mean((0, 1))
    print("%s: per-band aggregate L1 error max %.3f%%, band-mean bias max %.3f%%" % (name, 100 * l1.max(), 100 * bias.max()))
    l1_allowed, bias_allowed = 0.01, 0.003
    if os.path.exists(noise):
        # heavy-tailed scene (oracle/make_golden.py HEAVY_TAILED): two independent estimates at this sample count differ by
        # more than 1 % - measured with the CPU oracle, which is bit-identical to the reference per sample, rendering this
        # frame with the product's sampler and seed. The GPU image must reproduce the oracle's image, and be as close to the
        # reference render as the oracle's is.
        import json
        nz = json.load(open(noise))
        assert nz["seed"] == rp.seed and nz["spp"] == gpu_spp
        oimg = np.load(os.path.join(O.GOLDEN_BIG, "%s_%dspp.oracle.npy" % (name, spp))).astype(np.float64)
        l1o = np.abs(img - oimg).sum((0, 1)) / oimg.sum((0, 1))
        print("%s: against the oracle's render of the same samples: L1 max %.3f%% (oracle vs reference: L1 %.3f%%, bias %.3f%%)" % (
            name, 100 * l1o.max(), 100 * nz["l1_max"], 100 * nz["bias_max"]))
        assert l1o.max() <= 0.003, l1o
        l1_allowed, bias_allowed = max(l1_allowed, 1.05 * nz["l1_max"]), max(bias_allowed, 1.05 * nz["bias_max"])
    assert l1.max() <= l1_allowed, l1
    assert bias.max() <= bias_allowed, bias
