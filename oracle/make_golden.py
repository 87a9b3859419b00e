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


def bunny(w, h, spp, name, maxdepth=5):
    s = read(os.path.join(REF, "scenes/bunny.pbrt"))
    s = s.replace('Film "image"', 'Film "image" "integer xresolution" [%d] "integer yresolution" [%d]\n'
                  'Sampler "lowdiscrepancy" "integer pixelsamples" [%d]\n'
                  'SurfaceIntegrator "path" "integer maxdepth" [%d]' % (w, h, spp, maxdepth), 1)
    # the shipped bunny material is a measured BRDF ("next"); plastic stands in
    s = s.replace('Material "measured" "string filename" "brdfs/mystique.brdf"',
                  'Material "plastic" "color Kd" [.3 .25 .4] "color Ks" [.4 .4 .4] "float roughness" [.08]')
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
    # small committed fixture
    "tiny":            (tiny, 48, 48, 4, 700, 40, 0),
}


def with_gpupath(s):
    return s.replace("WorldBegin", 'Renderer "gpupath"\nWorldBegin', 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", action="store_true", help="also render reference .dat images")
    ap.add_argument("--only", default="", help="comma-separated config names")
    args = ap.parse_args()
    if not os.path.isdir(REF):
        sys.exit("reference tree %s not present: golden vectors can only be generated where it is" % REF)
    os.makedirs(SCENES, exist_ok=True)
    os.makedirs(GOLDEN, exist_ok=True)
    for d in ("geometry", "spds", "brdfs"):
        dst = os.path.join(SCENES, d)
        if not os.path.isdir(dst):
            shutil.copytree(os.path.join(REF, "scenes", d), dst)
    names = [n for n in CONFIGS if not args.only or n in args.only.split(",")]
    for name in names:
        build, w, h, spp, npix, nrng, img_spp = CONFIGS[name]
        s = build(w, h, spp, name)
        write(os.path.join(SCENES, name + ".pbrt"), s)
        write(os.path.join(SCENES, name + ".gpu.pbrt"), with_gpupath(s))
        prefix = os.path.join(TESTS_GOLDEN if name == "tiny" else GOLDEN, name)
        env = dict(os.environ, SPT_DUMP_PREFIX=prefix, SPT_DUMP_PIXELS=str(max(npix, 1)),
                   SPT_DUMP_NRNG=str(max(nrng, 1)), SPT_DUMP_LI="1" if npix else "0")
        t0 = time.time()
        subprocess.run([os.path.join(OUT, "bin/oracle_dump"), "--quiet", name + ".gpu.pbrt"],
                       cwd=SCENES, env=env, check=True, stdout=subprocess.DEVNULL)
        if not npix:
            # full-size workloads: only the lowered scene is kept, where bench.py looks for it
            os.remove(prefix + ".golden")
            os.makedirs(LOWERED, exist_ok=True)
            shutil.move(prefix + ".spt", os.path.join(LOWERED, name + ".spt"))
        print("%-16s lowered + golden in %.1fs" % (name, time.time() - t0), flush=True)
        if args.images and img_spp:
            iname = "%s_%dspp" % (name, img_spp)
            write(os.path.join(SCENES, iname + ".pbrt"), build(w, h, img_spp, iname))
            t0 = time.time()
            ncores = os.cpu_count()
            subprocess.run([os.path.join(OUT, "bin/pbrt"), "--quiet", "--ncores", str(ncores), iname + ".pbrt"],
                           cwd=SCENES, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            shutil.move(os.path.join(SCENES, iname + ".dat"), os.path.join(GOLDEN, iname + ".dat"))
            write(os.path.join(GOLDEN, iname + ".txt"), "ncores=%d seconds=%.1f\n" % (ncores, time.time() - t0))
            print("%-16s reference image %d spp in %.1fs" % (name, img_spp, time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
