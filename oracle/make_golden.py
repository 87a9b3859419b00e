#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: derive the BASELINE.json config scenes from the reference's shipped
scene files and generate golden vectors by RUNNING THE REFERENCE
(oracle/_ref/bin/oracle_dump, oracle/_ref/bin/pbrt). Needs /root/reference; outputs go to
oracle/_ref/scenes (derived .pbrt + copied geometry), oracle/_ref/golden (.spt/.golden/.dat) and,
for the small committed fixture, tests/golden/.

    python oracle/make_golden.py            # scenes + lowered scenes + per-sample golden vectors
    python oracle/make_golden.py --images   # also reference .dat renders (slow)
"""
import argparse
import os
import re
import shutil
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("SPT_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
SCENES = os.path.join(OUT, "scenes")
GOLDEN = os.path.join(OUT, "golden")
TESTS_GOLDEN = os.path.join(REPO, "tests", "golden")
LOWERED = os.path.join(REPO, "assets", "_lowered")


def read(p):
    with open(p) as f:
        return f.read()


def write(p, s):
    os.makedirs(os.path.dirname(p), exist_ok=True)
    with open(p, "w") as f:
        f.write(s)


def set_res(s, w, h):
    # replace an active resolution spec, or add one to the Film line
    s2, n = re.subn(r'^(\s*)"integer xresolution" \[\d+\] "integer yresolution" \[\d+\]',
                    r'\1"integer xresolution" [%d] "integer yresolution" [%d]' % (w, h), s, count=1, flags=re.M)
    if n:
        return s2
    s2, n = re.subn(r'(Film "image")( "integer xresolution" \[\d+\] "integer yresolution" \[\d+\])?',
                    r'\1 "integer xresolution" [%d] "integer yresolution" [%d]' % (w, h), s, count=1)
    assert n
    return s2


def set_spp(s, spp):
    s2, n = re.subn(r'Sampler "lowdiscrepancy" "integer pixelsamples" \[\d+\]',
                    'Sampler "lowdiscrepancy" "integer pixelsamples" [%d]' % spp, s)
    if not n:
        s2 = s.replace("WorldBegin", 'Sampler "lowdiscrepancy" "integer pixelsamples" [%d]\nWorldBegin' % spp, 1)
    return s2


def set_filename(s, name):
    s2, n = re.subn(r'"string filename" "[^"]*\.exr"', '"string filename" "%s.exr"' % name, s, count=1)
    if not n:
        s2 = re.sub(r'(Film "image")', r'\1 "string filename" "%s.exr"' % name, s, count=1)
    return s2


def killeroo(w, h, spp, name, maxdepth=5):
    s = read(os.path.join(REF, "scenes/killeroo-simple.pbrt"))
    s = s.replace('SurfaceIntegrator "directlighting"', 'SurfaceIntegrator "path" "integer maxdepth" [%d]' % maxdepth)
    return set_filename(set_spp(set_res(s, w, h), spp), name)


def bunny(w, h, spp, name, maxdepth=5, measured=False):
    s = read(os.path.join(REF, "scenes/bunny.pbrt"))
    s = s.replace('Film "image"', 'Film "image" "integer xresolution" [%d] "integer yresolution" [%d]\n'
                  'Sampler "lowdiscrepancy" "integer pixelsamples" [%d]\n'
                  'SurfaceIntegrator "path" "integer maxdepth" [%d]' % (w, h, spp, maxdepth), 1)
    assert 'Material "measured" "string filename" "brdfs/mystique.brdf"' in s
    if not measured:
        # the first golden sets were made before the measured BRDF was lowered: plastic stands in
        s = s.replace('Material "measured" "string filename" "brdfs/mystique.brdf"',
                      'Material "plastic" "color Kd" [.3 .25 .4] "color Ks" [.4 .4 .4] "float roughness" [.08]')
    return set_filename(s, name)


def bunny_measured(w, h, spp, name, maxdepth=5):
    """BASELINE config 2 with the bunny's shipped material: the measured BRDF brdfs/mystique.brdf (IrregIsotropicBRDF)."""
    return bunny(w, h, spp, name, maxdepth, measured=True)


def bunny_shipped(w, h, spp, name, maxdepth=5):
    """bunny.pbrt AS SHIPPED: default directlighting integrator, measured BRDF; only resolution and spp are set."""
    s = read(os.path.join(REF, "scenes/bunny.pbrt"))
    s = s.replace('Film "image"', 'Film "image" "integer xresolution" [%d] "integer yresolution" [%d]\n'
                  'Sampler "lowdiscrepancy" "integer pixelsamples" [%d]' % (w, h, spp), 1)
    assert 'Material "measured"' in s
    return set_filename(s, name)


def metal(w, h, spp, name, maxdepth=5):
    s = read(os.path.join(REF, "scenes/metal.pbrt"))
    s = re.sub(r'^Renderer "metropolis".*\n', '', s, flags=re.M)
    s = re.sub(r'^\s*"bool dodirectseparately".*\n', '', s, flags=re.M)
    # uffizi map is absent from the reference checkout: constant-L infinite light
    s = re.sub(r'\n\s*"string mapname" \["textures/uffizi_latlong.exr"\]', '', s)
    # floor: substrate + image textures + bump are "next" (SURVEY.md 8f N2): matte Kd 0.5
    s = re.sub(r'Texture "tmap".*?"texture bumpmap" "sbump" \n', 'Material "matte" "color Kd" [.5 .5 .5]\n', s, flags=re.S)
    s = s.replace("WorldBegin", 'SurfaceIntegrator "path" "integer maxdepth" [%d]\nWorldBegin' % maxdepth, 1)
    return set_filename(set_spp(set_res(s, w, h), spp), name)


def exr_to_pfm(src, dst):
    """The reference build here has no OpenEXR; its PFM reader does not flip rows
    (src/core/imageio.cpp:601-678), so the EXR's row order is kept."""
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    import cv2
    import numpy as np
    img = cv2.imread(src, cv2.IMREAD_UNCHANGED)
    assert img is not None, src
    rgb = np.ascontiguousarray(img[..., 2::-1].astype("<f4"))
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    with open(dst, "wb") as f:
        f.write(b"PF\n%d %d\n-1.0\n" % (rgb.shape[1], rgb.shape[0]))
        f.write(rgb.tobytes())


def envmap(w, h, spp, name, maxdepth=5):
    """BASELINE config 4 (ss-envmap): infinite area light with the grace lat-long map, importance sampled."""
    s = read(os.path.join(REF, "scenes/ss-envmap.pbrt"))
    s = re.sub(r'SurfaceIntegrator "dipolesubsurface".*\n\s*"float maxerror".*\n', 'SurfaceIntegrator "path" "integer maxdepth" [%d]\n' % maxdepth, s)
    s = s.replace("textures/grace_latlong.exr", "textures/grace_latlong.pfm")
    # floor: substrate + image textures + bump are "next" (SURVEY.md 8f N2): matte Kd 0.5
    s = re.sub(r'Texture "tmap".*?"texture bumpmap" "sbump" \n', 'Material "matte" "color Kd" [.5 .5 .5]\n', s, flags=re.S)
    # the teapot's `subsurface` material is a specular reflection BSDF under the path integrator
    # (subsurface.cpp:52-56); specular BxDFs are not lowered yet: plastic stands in
    s = re.sub(r'Material "subsurface".*\n\s*"color sigma_prime_s".*\n',
               'Material "plastic" "color Kd" [.4 .35 .3] "color Ks" [.5 .5 .5] "float roughness" [.05]\n', s)
    return set_filename(set_spp(set_res(s, w, h), spp), name)


def synth(w, h, spp, name, maxdepth=5, ntris=200000, chunks=8):
    """BASELINE config 5 recipe at test size: random triangle soup in [-1,1]^3, half matte /
    half plastic, a sphere area light plus a constant infinite light (two lights: exercises the light choice)."""
    import numpy as np
    rng = np.random.default_rng(12345)
    out = ['LookAt 0 -4 0  0 0 0  0 0 1', 'Camera "perspective" "float fov" [40]',
           'Film "image" "integer xresolution" [%d] "integer yresolution" [%d] "string filename" "%s.exr"' % (w, h, name),
           'Sampler "lowdiscrepancy" "integer pixelsamples" [%d]' % spp, 'PixelFilter "box"',
           'SurfaceIntegrator "path" "integer maxdepth" [%d]' % maxdepth, 'WorldBegin',
           'AttributeBegin', 'AreaLightSource "diffuse" "color L" [50 50 50]', 'Translate 0 0 3',
           'Shape "sphere" "float radius" [.2]', 'AttributeEnd',
           'LightSource "infinite" "color L" [.5 .5 .5]']
    per = ntris // chunks
    for c in range(chunks):
        centre = rng.uniform(-1, 1, (per, 3))
        def edge():
            d = rng.normal(size=(per, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
            return d * rng.uniform(0.002, 0.02, (per, 1))
        P = np.stack([centre, centre + edge(), centre + edge()], 1).reshape(-1, 3).astype(np.float32)
        kd = rng.uniform(0.2, 0.8, 3)
        if c % 2 == 0:
            out.append('Material "matte" "color Kd" [%g %g %g]' % tuple(kd))
        else:
            out.append('Material "plastic" "color Kd" [%g %g %g] "color Ks" [.3 .3 .3] "float roughness" [%g]' % (*kd, rng.uniform(0.01, 0.3)))
        out.append('Shape "trianglemesh" "integer indices" [' + " ".join(map(str, range(3 * per))) + ']')
        out.append('"point P" [' + " ".join("%.9g" % v for v in P.ravel()) + ']')
    out.append('WorldEnd')
    return "\n".join(out) + "\n"


def ssenv(w, h, spp, name, maxdepth=5):
    """BASELINE config 4 with the teapot's shipped `subsurface` material: under the path integrator its BSDF is a
    specular reflection with dielectric Fresnel (subsurface.cpp:40-58). Floor as in envmap()."""
    s = read(os.path.join(REF, "scenes/ss-envmap.pbrt"))
    s = re.sub(r'SurfaceIntegrator "dipolesubsurface".*\n\s*"float maxerror".*\n', 'SurfaceIntegrator "path" "integer maxdepth" [%d]\n' % maxdepth, s)
    s = s.replace("textures/grace_latlong.exr", "textures/grace_latlong.pfm")
    s = re.sub(r'Texture "tmap".*?"texture bumpmap" "sbump" \n', 'Material "matte" "color Kd" [.5 .5 .5]\n', s, flags=re.S)
    return set_filename(set_spp(set_res(s, w, h), spp), name)


def metal_shipped(w, h, spp, name, maxdepth=5):
    """BASELINE config 3 as shipped: Au teapot on the `substrate` floor with the lines.exr image map as Kd and,
    scaled by -0.25, as bump map. Only what cannot run here is replaced: the Metropolis
    renderer lines (path integrator instead, as BASELINE.json names it), .exr -> .pfm (no OpenEXR in this
    build), and the uffizi environment map that is absent from the checkout -> the grace map."""
    s = read(os.path.join(REF, "scenes/metal.pbrt"))
    s = re.sub(r'^Renderer "metropolis".*\n', '', s, flags=re.M)
    s = re.sub(r'^\s*"bool dodirectseparately".*\n', '', s, flags=re.M)
    s = s.replace("textures/uffizi_latlong.exr", "textures/grace_latlong.pfm").replace("textures/lines.exr", "textures/lines.pfm")
    s = s.replace("WorldBegin", 'SurfaceIntegrator "path" "integer maxdepth" [%d]\nWorldBegin' % maxdepth, 1)
    assert "substrate" in s and "sbump" in s
    return set_filename(set_spp(set_res(s, w, h), spp), name)


def ssenv_shipped(w, h, spp, name, maxdepth=5):
    """BASELINE config 4 as shipped (subsurface teapot, substrate floor with image-mapped Kd and bump map, grace
    environment light): path integrator instead of dipolesubsurface, .exr -> .pfm."""
    s = read(os.path.join(REF, "scenes/ss-envmap.pbrt"))
    s = re.sub(r'SurfaceIntegrator "dipolesubsurface".*\n\s*"float maxerror".*\n', 'SurfaceIntegrator "path" "integer maxdepth" [%d]\n' % maxdepth, s)
    s = s.replace("textures/grace_latlong.exr", "textures/grace_latlong.pfm").replace("textures/lines.exr", "textures/lines.pfm")
    assert "substrate" in s and "sbump" in s
    return set_filename(set_spp(set_res(s, w, h), spp), name)


def specular(w, h, spp, name, maxdepth=5):
    """killeroo-simple with one killeroo of glass and one mirror: specular reflection / transmission,
    emitted light seen through specular bounces (path.cpp:55-56)."""
    s = killeroo(w, h, spp, name, maxdepth)
    s = s.replace('Material "plastic" "color Kd" [.4 .2 .2] "color Ks" [.5 .5 .5]', 'Material "glass" "color Kr" [.9 .9 .9] "color Kt" [.9 .8 .7] "float index" [1.5]')
    s = s.replace('Material "plastic" "color Ks" [.3 .3 .3] "color Kd" [.4 .5 .4]', 'Material "mirror" "color Kr" [.8 .85 .9]')
    assert "glass" in s and "mirror" in s
    return s


def killeroo_direct(w, h, spp, name, maxdepth=5):
    """killeroo-simple.pbrt AS SHIPPED: SurfaceIntegrator "directlighting" (strategy all), the sphere light's
    nsamples 8."""
    s = read(os.path.join(REF, "scenes/killeroo-simple.pbrt"))
    assert 'SurfaceIntegrator "directlighting"' in s
    return set_filename(set_spp(set_res(s, w, h), spp), name)


def specular_direct(w, h, spp, name, maxdepth=5):
    """killeroo-simple.pbrt under its shipped directlighting integrator with one killeroo of glass and one mirror: the
    SpecularReflect / SpecularTransmit recursion of DirectLightingIntegrator::Li (directlighting.cpp:97-107)."""
    s = killeroo_direct(w, h, spp, name, maxdepth)
    s = s.replace('Material "plastic" "color Kd" [.4 .2 .2] "color Ks" [.5 .5 .5]', 'Material "glass" "color Kr" [.9 .9 .9] "color Kt" [.9 .8 .7] "float index" [1.5]')
    s = s.replace('Material "plastic" "color Ks" [.3 .3 .3] "color Kd" [.4 .5 .4]', 'Material "mirror" "color Kr" [.8 .85 .9]')
    assert "glass" in s and "mirror" in s and 'SurfaceIntegrator "directlighting"' in s
    return s


def killeroo_direct_one(w, h, spp, name, maxdepth=5):
    """killeroo-simple.pbrt with the directlighting integrator's other strategy ("one": UniformSampleOneLight)."""
    s = killeroo_direct(w, h, spp, name, maxdepth)
    s = s.replace('SurfaceIntegrator "directlighting"', 'SurfaceIntegrator "directlighting" "string strategy" ["one"]')
    return s


def bunny_direct(w, h, spp, name, maxdepth=5):
    """bunny.pbrt as shipped (default directlighting integrator, point light + disk area light with nsamples 4) except
    its measured BRDF ("next"), for which plastic stands in."""
    s = read(os.path.join(REF, "scenes/bunny.pbrt"))
    s = s.replace('Film "image"', 'Film "image" "integer xresolution" [%d] "integer yresolution" [%d]\n'
                  'Sampler "lowdiscrepancy" "integer pixelsamples" [%d]' % (w, h, spp), 1)
    s = s.replace('Material "measured" "string filename" "brdfs/mystique.brdf"',
                  'Material "plastic" "color Kd" [.3 .25 .4] "color Ks" [.4 .4 .4] "float roughness" [.08]')
    return set_filename(s, name)


def tiny_pattern_pfm(dst):
    """A small procedural image (24 x 20: not a power of two, so the reference's MIPMap resamples it) for the committed
    texture fixture: own authoring, written on the spot - the lowered scene carries the MIP pyramid the reference built."""
    import numpy as np
    y, x = np.mgrid[0:20, 0:24]
    r = 0.15 + 0.8 * (((x // 3) + (y // 4)) % 2)
    g = 0.2 + 0.6 * np.abs(np.sin(0.7 * x)) * (y / 19.0)
    b = 0.9 - 0.7 * ((x + 2 * y) % 5) / 4.0
    rgb = np.stack([r, g, b], -1).astype("<f4")
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    with open(dst, "wb") as f:
        f.write(b"PF\n24 20\n-1.0\n")
        f.write(rgb.tobytes())


def tiny_tex(w, h, spp, name, maxdepth=5):
    """Committed fixture for the extended materials: the tiny scene with a substrate floor whose Kd is an
    EWA-filtered image map and whose bump map is the same image scaled by a constant, and with a trilinear-filtered image
    map + bump map on the sphere that has per-vertex normals and uvs (dndu / dndv)."""
    s = read(os.path.join(TESTS_GOLDEN, "tiny.pbrt"))
    tex = ('Texture "tmap" "color" "imagemap" "string filename" "textures/tiny_pattern.pfm" "float uscale" 2 "float vscale" 2\n'
           'Texture "tbump-tex" "float" "imagemap" "string filename" "textures/tiny_pattern.pfm" "float uscale" 2 "float vscale" 2\n'
           'Texture "sbump" "float" "scale" "texture tex1" "tbump-tex" "float tex2" [-.05]\n'
           'Texture "tri" "color" "imagemap" "string filename" "textures/tiny_pattern.pfm" "bool trilinear" ["true"] "float uscale" 3\n'
           'Texture "tribump" "float" "imagemap" "string filename" "textures/tiny_pattern.pfm" "bool trilinear" ["true"] "float uscale" 3 "float scale" [.02]\n')
    a = 'Material "matte" "color Kd" [.55 .5 .45]'
    b = 'Material "plastic" "color Kd" [.45 .2 .15] "color Ks" [.5 .5 .5] "float roughness" [.04]'
    assert a in s and b in s
    s = s.replace(a, tex + 'Material "substrate" "texture Kd" "tmap" "color Ks" [.4 .4 .4] "float uroughness" [.05] "float vroughness" [.2] '
                  '"texture bumpmap" "sbump"')
    s = s.replace(b, 'Material "plastic" "texture Kd" "tri" "color Ks" [.5 .5 .5] "float roughness" [.04] "texture bumpmap" "tribump"')
    return set_filename(set_spp(set_res(s, w, h), spp), name)


def tiny_direct(w, h, spp, name, maxdepth=5):
    """Committed fixture for the directlighting integrator: the tiny scene under strategy "all", its area
    lights with 4, 2 and 1 samples, plus the point light."""
    s = read(os.path.join(TESTS_GOLDEN, "tiny.pbrt"))
    s = s.replace('SurfaceIntegrator "path" "integer maxdepth" [5]', 'SurfaceIntegrator "directlighting"')
    s = s.replace('AreaLightSource "area" "color L" [40 36 30]', 'AreaLightSource "area" "color L" [40 36 30] "integer nsamples" [4]')
    s = s.replace('AreaLightSource "area" "color L" [9 10 12]', 'AreaLightSource "area" "color L" [9 10 12] "integer nsamples" [2]')
    assert "directlighting" in s and '"integer nsamples" [4]' in s
    return set_filename(set_spp(set_res(s, w, h), spp), name)


def synth_merl(dst):
    """A MERL-format half-angle table (measured.cpp:122-170 reads: int dims[3] = 90, 90, 180, then per colour channel
    90*90*180 doubles, which it scales by 1/1500, 1.15/1500, 1.66/1500): no such file ships with the reference, so an
    analytic glossy lobe over a coloured diffuse base is written on the spot (own authoring). theta_h is sampled on the
    format's square-root scale."""
    import numpy as np
    if os.path.exists(dst):
        return
    ih, idd, ip = np.meshgrid(np.arange(90), np.arange(90), np.arange(180), indexing="ij")
    theta_h = ((ih + 0.5) / 90.0) ** 2 * (np.pi / 2)
    theta_d = (idd + 0.5) / 90.0 * (np.pi / 2)
    phi_d = (ip + 0.5) / 180.0 * np.pi
    lobe = 6.0 * np.exp(-(theta_h / 0.12) ** 2) * (1.0 + 2.0 * (theta_d / (np.pi / 2)) ** 4)
    aniso = 1.0 + 0.15 * np.cos(2 * phi_d)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    with open(dst, "wb") as f:
        f.write(np.array([90, 90, 180], np.int32).tobytes())
        for base, scale in ((0.55, 1.0 / 1500), (0.30, 1.15 / 1500), (0.18, 1.66 / 1500)):
            brdf = base / np.pi + lobe * aniso
            f.write((brdf / scale).astype(np.float64).tobytes())


def tiny_merl(w, h, spp, name, maxdepth=5):
    """Fixture for the half-angle (MERL) measured BRDF: the tiny scene with RegularHalfangleBRDF on the
    plastic sphere."""
    s = read(os.path.join(TESTS_GOLDEN, "tiny.pbrt"))
    b = 'Material "plastic" "color Kd" [.45 .2 .15] "color Ks" [.5 .5 .5] "float roughness" [.04]'
    assert b in s
    s = s.replace(b, 'Material "measured" "string filename" "brdfs/synth_merl.binary"')
    return set_filename(set_spp(set_res(s, w, h), spp), name)


def tiny(w, h, spp, name, maxdepth=5):
    s = read(os.path.join(TESTS_GOLDEN, "tiny.pbrt"))
    return set_filename(set_spp(set_res(s, w, h), spp), name)


# name -> (builder, w, h, spp, dump pixels, n_rng, reference-image spp or 0)
CONFIGS = {
    # config 1 of BASELINE.json at full size (bench workload) and a small golden-vector variant
    "killeroo_path":   (killeroo, 700, 700, 64, 0, 0, 0),
    "killeroo_small":  (killeroo, 176, 176, 4, 6000, 40, 1024),
    # config 2: first-hit parity on bunny (all camera rays of a 4-spp 320x240 frame)
    "bunny_path":      (bunny, 640, 480, 256, 0, 0, 0),
    "bunny_small":     (bunny, 320, 240, 4, 8000, 40, 4096),
    # config 3 (metal teapot, Au SPDs) with the substitutions noted above
    "metal_small":     (metal, 200, 200, 4, 6000, 40, 512),
    # config 4 (ss-envmap): grace environment map, Distribution2D importance sampling
    "envmap_small":    (envmap, 200, 200, 4, 6000, 40, 8192),
    # config 5 recipe at test size: 200 000 random triangles, two lights
    "synth_small":     (synth, 256, 144, 4, 6000, 40, 2048),
    # config 4 with the shipped subsurface teapot (specular reflection BSDF); glass + mirror killeroos
    "ssenv_small":     (ssenv, 200, 200, 4, 6000, 40, 4096),
    "specular_small":  (specular, 176, 176, 4, 6000, 40, 2048),
    # configs 3 and 4 with their shipped floor: substrate (FresnelBlend + Anisotropic), image-mapped Kd (EWA), bump map
    "metal_shipped_small":  (metal_shipped, 200, 200, 4, 6000, 40, 8192),
    "ssenv_shipped_small":  (ssenv_shipped, 200, 200, 4, 6000, 40, 8192),
    # configs 3 and 4 at BASELINE's full size, shipped floor: optional bench workloads (bench.py --workload metal_path / ssenv_path)
    "metal_path":      (metal_shipped, 400, 400, 512, 0, 0, 0),
    "ssenv_path":      (ssenv_shipped, 1920, 1080, 1024, 0, 0, 0),
    # the shipped scenes under their own integrator (directlighting, strategy all): SURVEY.md 8f N3
    "killeroo_direct_small": (killeroo_direct, 176, 176, 4, 6000, 8, 1024),
    "bunny_direct_small":    (bunny_direct, 320, 240, 4, 6000, 8, 1024),
    "killeroo_direct":       (killeroo_direct, 700, 700, 64, 0, 0, 0),
    "killeroo_direct_one_small": (killeroo_direct_one, 176, 176, 4, 3000, 8, 0),
    # directlighting with specular materials: the SpecularReflect / SpecularTransmit recursion
    "specular_direct_small": (specular_direct, 176, 176, 4, 3000, 8, 1024),
    # the bunny with its shipped measured BRDF: under the path integrator (config 2) and as shipped (directlighting)
    "bunny_measured_small":  (bunny_measured, 320, 240, 4, 6000, 40, 4096),
    "bunny_shipped_small":   (bunny_shipped, 320, 240, 4, 6000, 8, 1024),
    "bunny_shipped":         (bunny_shipped, 640, 480, 256, 0, 0, 0),
    # config 5 recipe at 1 M triangles (BVH + pair nodes + vertices = 176 MB, larger than L2): optional bench workload
    "synth_1m":        (lambda w, h, spp, name, maxdepth=5: synth(w, h, spp, name, maxdepth, ntris=1000000, chunks=10), 1024, 576, 16, 0, 0, 0),
    # first hits at BASELINE resolution (configs 1 and 2): 262 144 random pixels x 4 spp = 1 048 576 camera rays each, compact
    # records (COMPACT below); the _mid variants (16 384 x 4 = 65 536 rays) are committed under tests/golden/ with the lowered
    # scene xz-compressed, so that a fresh clone checks first hits on the real scenes, not only on the tiny fixtures
    "killeroo_rays":     (killeroo, 700, 700, 4, 262144, 0, 0),
    "bunny_rays":        (bunny, 640, 480, 4, 262144, 0, 0),
    "killeroo_rays_mid": (killeroo, 700, 700, 4, 16384, 0, 0),
    "bunny_rays_mid":    (bunny, 640, 480, 4, 16384, 0, 0),
    # the 30-band variant (BASELINE.json's metric: "30-band spectral path trace"): the same scene through the reference built
    # with nSpectralSamples = 30 (oracle/Makefile ref30 -> oracle/_ref/bin30); BANDS30 below routes these to that build
    "killeroo_path30":  (killeroo, 700, 700, 64, 0, 0, 0),
    "killeroo_small30": (killeroo, 176, 176, 4, 6000, 40, 1024),
    # the half-angle (MERL) measured BRDF on the tiny scene; the 17 MB table keeps it out of tests/golden
    "tiny_merl_small": (tiny_merl, 64, 64, 4, 1500, 40, 32768),
    # small committed fixture
    "tiny":            (tiny, 48, 48, 4, 700, 40, 0),
    # two more committed fixtures (tests/golden/): extended materials, directlighting
    "tiny_tex":        (tiny_tex, 48, 48, 4, 300, 40, 0),
    "tiny_direct":     (tiny_direct, 48, 48, 4, 300, 8, 0),
}


# small golden scene -> the full-size bench workload (assets/_lowered/, product data) it shares every table with: same scene
# file, other resolution / spp. The golden scene is stored as a delta of the workload, never the other way round: bench.py
# reads nothing under oracle/.
BANDS30 = {"killeroo_path30", "killeroo_small30"}
COMPACT = {"killeroo_rays", "bunny_rays", "killeroo_rays_mid", "bunny_rays_mid"}
COMMITTED = {"killeroo_rays_mid", "bunny_rays_mid"}
DELTA_BASE = {"killeroo_small30": "killeroo_path30", "killeroo_rays": "killeroo_path", "bunny_rays": "bunny_path", "killeroo_small": "killeroo_path", "bunny_small": "bunny_path", "metal_shipped_small": "metal_path",
              "ssenv_shipped_small": "ssenv_path", "killeroo_direct_small": "killeroo_direct",
              "bunny_shipped_small": "bunny_shipped"}


def write_delta(path, name):
    """Replace a golden lowered scene by {base_scene, camera, film, params, film_filename} when every other array is
    byte-identical to the full-size workload's (checked): keeps the snapshot shipped to the GPU box small."""
    base = DELTA_BASE.get(name)
    base_path = os.path.join(LOWERED, (base or "") + ".spt")
    if not base or not os.path.exists(base_path):
        return
    sys.path.insert(0, REPO)
    import numpy as np
    from pbrt_v2_spectral_b200.scene_io import LoweredScene, load_container, save_container
    a, b = load_container(path), LoweredScene.load_arrays(base_path)
    own = ("camera", "film", "params", "film_filename")
    if set(a) != set(b) or any(not np.array_equal(a[k], b[k]) for k in a if k not in own):
        return
    delta = {"base_scene": np.frombuffer(os.path.relpath(base_path, os.path.dirname(path)).encode(), np.uint8)}
    delta.update({k: a[k] for k in own if k in a})
    save_container(path, delta)


# Scenes whose per-sample radiance is heavy-tailed (Blinn exponent 1000 under an HDR environment map): two independent
# 8192-spp estimates of the image differ by more than the 1 % the converged-image test asks for. For these the CPU oracle
# (bit-identical to the reference per sample) renders the same frame with the PRODUCT's sampler and seed: the GPU image
# must reproduce that image, and its distance to the reference render must be the oracle's (tests/test_image_parity.py).
HEAVY_TAILED = {"metal_shipped_small", "bunny_measured_small"}     # the measured lacquer BRDF is glossy and only cosine-sampled
IMAGE_SEED = 2024


# The reference's SpectralImageFilm::WriteImage adds `splatScale * splatC[nSpectralSamples]` to every band of a pixel
# (src/film/spectralImage.cpp:307-313): one float PAST the splat array, i.e. Pixel::pad, which Pixel() never initialises
# (src/film/spectralImage.h:74-84). The film is allocated at WorldEnd, after the scene's textures were read and their
# temporary buffers freed, so on scenes with an environment map `pad` holds stale texels: a constant, band-independent,
# spp-independent offset per pixel (0.36 per pixel on the grace-map scenes here = 65 samples' worth of radiance at 64 spp,
# 0.4 % of the image at 8192 spp - found as a "bias" that two reference renders shared and eight GPU seeds did not).
# Undefined behaviour is not a parity target: every reference IMAGE is rendered with glibc's MALLOC_PERTURB_=255, which
# hands out zero-filled allocations, so the term reads as zero; the binary itself stays unmodified.
REF_RENDER_ENV = dict(os.environ, MALLOC_PERTURB_="255")


def pad_garbage(name, spp_tag):
    """What the uninitialised Pixel::pad added to an EXISTING reference image of config `name`: the difference of two 1-spp
    renders of the same scene (same samples, one task) without and with zero-filled allocations. The film is allocated
    before any sampler or thread exists, so the garbage does not depend on spp or --ncores."""
    sys.path.insert(0, REPO)
    import numpy as np
    from pbrt_v2_spectral_b200 import capi
    build, w, h, _spp, _npix, _nrng, _img = CONFIGS[name]
    out = []
    for tag, env in (("g", dict(os.environ)), ("z", REF_RENDER_ENV)):
        iname = "%s_pad%s" % (name, tag)
        write(os.path.join(SCENES, iname + ".pbrt"), build(w, h, 1, iname))
        subprocess.run([os.path.join(OUT, "bin/pbrt"), "--quiet", "--ncores", "1", iname + ".pbrt"], cwd=SCENES, env=env, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        out.append(capi.read_dat(os.path.join(SCENES, iname + ".dat")))
        os.remove(os.path.join(SCENES, iname + ".dat")); os.remove(os.path.join(SCENES, iname + ".pbrt"))
    return out[0] - out[1]


def fix_pad():
    """Subtract the pad garbage from the reference images rendered before REF_RENDER_ENV existed (once: a marker is kept)."""
    import json
    import numpy as np
    for f in sorted(os.listdir(GOLDEN)):
        m = re.match(r"(.+)_(\d+)spp\.(ref2?)\.npy$", f)
        if not m or m.group(1) not in CONFIGS:
            continue
        marker = os.path.join(GOLDEN, f[:-4] + ".padfix.json")
        if os.path.exists(marker):
            continue
        E = pad_garbage(m.group(1), m.group(2))
        img = np.load(os.path.join(GOLDEN, f))
        if E.max() > 0:
            np.save(os.path.join(GOLDEN, f), (img.astype(np.float64) - E).astype(np.float32))
        with open(marker, "w") as fp:
            json.dump({"pad_mean": float(E.mean()), "pad_max": float(E.max()), "share_of_image": float(E.sum() / max(img.sum(), 1e-30)),
                       "how": "oracle/make_golden.py --fix-pad: 1-spp reference renders without / with MALLOC_PERTURB_=255"}, fp)
        print("%-40s pad garbage: mean %.4g per pixel and band, %.3f %% of the image" % (f, E.mean(), 100 * E.sum() / max(img.sum(), 1e-30)), flush=True)


def dat_to_npy(dat, npy):
    """The reference's .dat ([band][x][y] float64 sums, spectralImage.cpp:319-369) kept as [y][x][band] float32: half the
    bytes in the snapshot shipped to the GPU box; the 2^-24 relative rounding is far below the images' Monte-Carlo noise."""
    sys.path.insert(0, REPO)
    import numpy as np
    from pbrt_v2_spectral_b200 import capi
    np.save(npy, capi.read_dat(dat).astype(np.float32))
    os.remove(dat)


def noise_floor(name, spp):
    sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
    import json
    import numpy as np
    import oracle_lib as O
    from pbrt_v2_spectral_b200 import capi, ctypes_defs as D
    sc, _ = O.load_case(os.path.join(GOLDEN, name + ".spt"), os.path.join(GOLDEN, name + ".golden"))
    ref = np.load(os.path.join(GOLDEN, "%s_%dspp.ref.npy" % (name, spp))).astype(np.float64) / spp
    rp = D.SptRenderParams.from_buffer_copy(bytes(sc.params)); rp.spp = spp; rp.seed = IMAGE_SEED
    t0 = time.time()
    c, _w = O.render(sc, rp)
    img = c.astype(np.float64) / spp
    np.save(os.path.join(GOLDEN, "%s_%dspp.oracle.npy" % (name, spp)), img.astype(np.float32))
    l1 = np.abs(img - ref).sum((0, 1)) / ref.sum((0, 1))
    bias = np.abs(img.mean((0, 1)) - ref.mean((0, 1))) / ref.mean((0, 1))
    with open(os.path.join(GOLDEN, "%s_%dspp.noise.json" % (name, spp)), "w") as f:
        json.dump({"l1_max": float(l1.max()), "bias_max": float(bias.max()), "seed": IMAGE_SEED, "spp": spp,
                   "how": "oracle/make_golden.py --images: orc_render (CPU oracle, product sampler) vs the reference .dat"}, f)
    print("%-16s oracle render %d spp in %.1fs: L1 %.3f%%, bias %.3f%% against the reference render" % (
        name, spp, time.time() - t0, 100 * l1.max(), 100 * bias.max()), flush=True)


# Reference-vs-reference floor of the converged-image check for the heavy-tailed scenes: a SECOND, independent render by the
# unmodified reference (another --ncores gives another task split, hence other per-task RNG seeds and sample scrambles:
# src/renderers/samplerrenderer.cpp:203-214,66-71) at the same spp. Their mutual per-band L1 / band-mean bias is what the
# metric reads when both images are correct; tests/test_image_parity.py compares the GPU image with the mean of the two
# renders against max(1 %, 1.2 x floor). The numbers are committed (tests/golden/image_floor.json), the images travel in
# oracle/_ref/golden.
FLOOR_NCORES = {"metal_shipped_small": 16, "bunny_measured_small": 32}


def reference_floor(name, spp):
    import json
    import numpy as np
    build, w, h, _spp, _npix, _nrng, img_spp = CONFIGS[name]
    assert spp == img_spp
    iname = "%s_%dspp" % (name, spp)
    ref1 = os.path.join(GOLDEN, iname + ".ref.npy")
    ref2 = os.path.join(GOLDEN, iname + ".ref2.npy")
    n2 = FLOOR_NCORES[name]
    if not os.path.exists(ref2):
        i2 = iname + "_b"
        write(os.path.join(SCENES, i2 + ".pbrt"), build(w, h, spp, i2))
        t0 = time.time()
        subprocess.run([os.path.join(OUT, "bin/pbrt"), "--quiet", "--ncores", str(n2), i2 + ".pbrt"],
                       cwd=SCENES, env=REF_RENDER_ENV, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        dat_to_npy(os.path.join(SCENES, i2 + ".dat"), ref2)
        os.remove(os.path.join(SCENES, i2 + ".pbrt"))
        print("%-16s second reference render (--ncores %d) in %.1fs" % (name, n2, time.time() - t0), flush=True)
    a = np.load(ref1).astype(np.float64) / spp
    b = np.load(ref2).astype(np.float64) / spp
    assert not np.array_equal(a, b), "the two reference renders are identical: same task split"
    l1 = np.abs(a - b).sum((0, 1)) / (0.5 * (a + b)).sum((0, 1))
    bias = np.abs(a.mean((0, 1)) - b.mean((0, 1))) / (0.5 * (a + b)).mean((0, 1))
    fp = os.path.join(TESTS_GOLDEN, "image_floor.json")
    floor = json.load(open(fp)) if os.path.exists(fp) else {}
    floor[name] = {"spp": spp, "l1_max": float(l1.max()), "bias_max": float(bias.max()),
                   "how": "two renders of oracle/_ref/bin/pbrt (unmodified reference), --ncores %s vs %d: per-band "
                          "sum|a-b| / sum mean(a,b) and |mean a - mean b| / mean" % (read(os.path.join(GOLDEN, iname + ".txt")).split()[0].split("=")[1], n2)}
    with open(fp, "w") as f:
        json.dump(floor, f, indent=1, sort_keys=True)
    print("%-16s reference vs reference at %d spp: L1 %.3f%%, bias %.3f%%" % (name, spp, 100 * l1.max(), 100 * bias.max()), flush=True)


def share_arrays(path, donor, keys=("tex_texels",)):
    """Replace arrays of the container `path` that are byte-identical in `donor` by "<array>@" references to it (the loader
    resolves them): the shipped-floor scenes carry the same 22 MB texel pool."""
    if not (os.path.exists(path) and os.path.exists(donor)) or os.path.abspath(path) == os.path.abspath(donor):
        return
    sys.path.insert(0, REPO)
    import numpy as np
    from pbrt_v2_spectral_b200.scene_io import load_container, save_container
    a, b = load_container(path), load_container(donor)
    changed = False
    for k in keys:
        if k in a and k in b and a[k].size > (1 << 18) and np.array_equal(a[k], b[k]):
            del a[k]
            a[k + "@"] = np.frombuffer(os.path.relpath(donor, os.path.dirname(path)).encode(), np.uint8)
            changed = True
    if changed:
        save_container(path, a)


# scene files that repeat the geometry of a bench workload (other materials / integrator): every array over 256 KB that is
# byte-identical in the donor becomes a "<array>@" reference - the snapshot shipped to the GPU box is capped at 512 MiB
SHARE_DONOR = {"specular_direct_small": "killeroo_path", "bunny_direct_small": "bunny_path", "bunny_measured_small": "bunny_path", "killeroo_direct_one_small": "killeroo_path",
               "specular_small": "killeroo_path", "bunny_shipped": "bunny_path", "killeroo_direct": "killeroo_path", "killeroo_path30": "killeroo_path"}


def share_all(path, donor):
    if not (os.path.exists(path) and os.path.exists(donor)):
        return
    sys.path.insert(0, REPO)
    import numpy as np
    from pbrt_v2_spectral_b200.scene_io import load_container, save_container
    a, b = load_container(path), load_container(donor)
    if "base_scene" in a or "base_scene" in b:
        return
    changed = False
    for k in list(a):
        if not k.endswith("@") and k in b and a[k].nbytes > (1 << 18) and a[k].dtype == b[k].dtype and a[k].shape == b[k].shape and np.array_equal(a[k], b[k]):
            del a[k]
            a[k + "@"] = np.frombuffer(os.path.relpath(donor, os.path.dirname(path)).encode(), np.uint8)
            changed = True
    if changed:
        save_container(path, a)


def with_gpupath(s):
    return s.replace("WorldBegin", 'Renderer "gpupath"\nWorldBegin', 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", action="store_true", help="also render reference .dat images")
    ap.add_argument("--only", default="", help="comma-separated config names")
    ap.add_argument("--fix-pad", action="store_true", help="only: subtract the uninitialised Pixel::pad term from existing reference images")
    ap.add_argument("--floor", action="store_true", help="only: second reference render + reference-vs-reference floor of the heavy-tailed scenes")
    args = ap.parse_args()
    if args.fix_pad:
        fix_pad()
        return
    if args.floor:
        for name in sorted(HEAVY_TAILED):
            if not args.only or name in args.only.split(","):
                reference_floor(name, CONFIGS[name][6])
        return
    names = [n for n in CONFIGS if not args.only or n in args.only.split(",")]
    names.sort(key=lambda n: (n in DELTA_BASE, n == "ssenv_path"))   # full-size workloads first (metal_path before ssenv_path): the golden variants refer to them
    # the synthetic scenes are written from scratch and lowered by the BUILT reference (oracle/_ref/bin/oracle_dump,
    # which travels with the repo): they can be generated where the reference source tree is absent (the GPU box)
    need_ref = any(not n.startswith("synth") for n in names)
    if need_ref and not os.path.isdir(REF):
        sys.exit("reference tree %s not present: golden vectors can only be generated where it is" % REF)
    os.makedirs(SCENES, exist_ok=True)
    os.makedirs(GOLDEN, exist_ok=True)
    if need_ref:
        for d in ("geometry", "spds", "brdfs"):
            dst = os.path.join(SCENES, d)
            if not os.path.isdir(dst):
                shutil.copytree(os.path.join(REF, "scenes", d), dst)
        for tex in ("grace_latlong", "lines"):
            pfm = os.path.join(SCENES, "textures", tex + ".pfm")
            if not os.path.exists(pfm):
                exr_to_pfm(os.path.join(REF, "scenes", "textures", tex + ".exr"), pfm)
    tiny_pattern_pfm(os.path.join(SCENES, "textures", "tiny_pattern.pfm"))
    if any(n.startswith("tiny_merl") for n in names):
        synth_merl(os.path.join(SCENES, "brdfs", "synth_merl.binary"))
    for name in names:
        build, w, h, spp, npix, nrng, img_spp = CONFIGS[name]
        s = build(w, h, spp, name)
        write(os.path.join(SCENES, name + ".pbrt"), s)
        write(os.path.join(SCENES, name + ".gpu.pbrt"), with_gpupath(s))
        prefix = os.path.join(TESTS_GOLDEN if (name.startswith("tiny") and not name.endswith("_small")) or name in COMMITTED else GOLDEN, name)
        env = dict(os.environ, SPT_DUMP_PREFIX=prefix, SPT_DUMP_PIXELS=str(max(npix, 1)),
                   SPT_DUMP_NRNG=str(max(nrng, 1)), SPT_DUMP_LI="1" if npix else "0",
                   SPT_DUMP_COMPACT="1" if name in COMPACT else "0")
        t0 = time.time()
        bindir = "bin30" if name in BANDS30 else "bin"
        subprocess.run([os.path.join(OUT, bindir, "oracle_dump"), "--quiet", name + ".gpu.pbrt"],
                       cwd=SCENES, env=env, check=True, stdout=subprocess.DEVNULL)
        if not npix:
            # full-size workloads: only the lowered scene is kept, where bench.py looks for it
            os.remove(prefix + ".golden")
            os.makedirs(LOWERED, exist_ok=True)
            dst = os.path.join(LOWERED, name + ".spt")
            shutil.move(prefix + ".spt", dst)
            if name == "ssenv_path":
                share_arrays(dst, os.path.join(LOWERED, "metal_path.spt"))
        if npix and name in DELTA_BASE:
            write_delta(prefix + ".spt", name)
        if name in ("envmap_small", "ssenv_small"):       # the grace map's tables (10 MB) are those of the config-4 workload
            share_arrays(prefix + ".spt", os.path.join(LOWERED, "ssenv_path.spt"), keys=("env_rgb", "env_cdf", "env_func"))
        if name in COMMITTED:
            # committed fixtures: xz-compressed containers (the loader opens .xz transparently)
            import lzma
            for ext in (".spt", ".golden"):
                with open(prefix + ext, "rb") as f:
                    raw = f.read()
                with open(prefix + ext + ".xz", "wb") as f:
                    f.write(lzma.compress(raw, preset=9 | lzma.PRESET_EXTREME))
                os.remove(prefix + ext)
        if name in SHARE_DONOR:
            share_all(os.path.join(LOWERED if not npix else GOLDEN, name + ".spt"), os.path.join(LOWERED, SHARE_DONOR[name] + ".spt"))
        print("%-16s lowered + golden in %.1fs" % (name, time.time() - t0), flush=True)
        if name.startswith("synth"):                 # tens of MB of text per scene: regenerable, not shipped
            for f in (name + ".pbrt", name + ".gpu.pbrt"):
                if not (args.images and img_spp) or f.endswith(".gpu.pbrt"):
                    os.remove(os.path.join(SCENES, f))
        if args.images and img_spp:
            iname = "%s_%dspp" % (name, img_spp)
            write(os.path.join(SCENES, iname + ".pbrt"), build(w, h, img_spp, iname))
            t0 = time.time()
            ncores = os.cpu_count()
            subprocess.run([os.path.join(OUT, bindir, "pbrt"), "--quiet", "--ncores", str(ncores), iname + ".pbrt"],
                           cwd=SCENES, env=REF_RENDER_ENV, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            with open(os.path.join(GOLDEN, iname + ".ref.padfix.json"), "w") as fp:
                fp.write('{"how": "rendered with MALLOC_PERTURB_=255: Pixel::pad reads as zero"}')
            dat_to_npy(os.path.join(SCENES, iname + ".dat"), os.path.join(GOLDEN, iname + ".ref.npy"))
            if name.startswith("synth"):
                for f in (name + ".pbrt", iname + ".pbrt"):
                    if os.path.exists(os.path.join(SCENES, f)):
                        os.remove(os.path.join(SCENES, f))
            write(os.path.join(GOLDEN, iname + ".txt"), "ncores=%d seconds=%.1f\n" % (ncores, time.time() - t0))
            if name in HEAVY_TAILED:
                noise_floor(name, img_spp)
            print("%-16s reference image %d spp in %.1fs" % (name, img_spp, time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
