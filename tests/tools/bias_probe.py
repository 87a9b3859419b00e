"""TEST INFRASTRUCTURE. Where does a band-mean difference between the reference's render and the product's come from?
Renders variants of a scene with the unmodified reference binary and with the CPU oracle running the PRODUCT's sampler
(orc_render: what the GPU reproduces to 0.03 %), at a small resolution, and prints per-band mean ratios.
    python tests/tools/bias_probe.py <variant> [res] [spp] [nseeds]"""
import os, re, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import make_golden as G
import oracle_lib as O
from pbrt_v2_spectral_b200 import capi, ctypes_defs as D

variant = sys.argv[1]
res = int(sys.argv[2]) if len(sys.argv) > 2 else 100
spp = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
nseeds = int(sys.argv[4]) if len(sys.argv) > 4 else 2
name = "probe_" + variant
s = G.metal_shipped(res, res, spp, name)
if "plastic_teapot" in variant:
    s = re.sub(r'Material "metal"  "float roughness" \[.001\]\s*"spectrum eta" "[^"]*"\s*"spectrum k" "[^"]*"',
               'Material "plastic" "color Kd" [.4 .35 .3] "color Ks" [.5 .5 .5] "float roughness" [.05]', s)
    assert "plastic" in s
if "rough_metal" in variant:
    s = s.replace('"float roughness" [.001]', '"float roughness" [.05]')
if "matte_floor" in variant:
    s = re.sub(r'Material "substrate" "texture Kd" "tmap"\s*"color Ks" \[.5 .5 .5\] "float uroughness" \[.05\]\s*"float vroughness" \[.05\]\s*"texture bumpmap" "sbump"',
               'Material "matte" "color Kd" [.5 .5 .5]', s)
    assert '"matte"' in s
if "nobump" in variant:
    s = s.replace('"texture bumpmap" "sbump"', '')
if "constkd" in variant:
    s = s.replace('"texture Kd" "tmap"', '"color Kd" [.5 .5 .5]')
if "constlight" in variant:
    s = re.sub(r'\n\s*"string mapname" \["textures/grace_latlong.pfm"\]', '', s)
if "depth1" in variant:
    s = s.replace('"integer maxdepth" [5]', '"integer maxdepth" [1]')
if "depth0" in variant:
    s = s.replace('"integer maxdepth" [5]', '"integer maxdepth" [0]')
G.write(os.path.join(G.SCENES, name + ".pbrt"), s)
G.write(os.path.join(G.SCENES, name + ".gpu.pbrt"), G.with_gpupath(s))
prefix = os.path.join("/tmp/probe", name)
env = dict(os.environ, SPT_DUMP_PREFIX=prefix, SPT_DUMP_PIXELS="1", SPT_DUMP_NRNG="1", SPT_DUMP_LI="0")
subprocess.run([os.path.join(G.OUT, "bin/oracle_dump"), "--quiet", name + ".gpu.pbrt"], cwd=G.SCENES, env=env, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
refs = []
for k, nc in enumerate((8, 16)[:max(1, min(2, nseeds))]):
    t0 = time.time()
    subprocess.run([os.path.join(G.OUT, "bin/pbrt"), "--quiet", "--ncores", str(nc), name + ".pbrt"], cwd=G.SCENES, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    refs.append(capi.read_dat(os.path.join(G.SCENES, name + ".dat")) / spp)
    print("reference render %d: %.1fs" % (k, time.time() - t0), flush=True)
os.remove(os.path.join(G.SCENES, name + ".dat"))
for f in (name + ".pbrt", name + ".gpu.pbrt"):
    os.remove(os.path.join(G.SCENES, f))
ref = np.mean(refs, 0)
sc, _ = O.load_case(prefix + ".spt", prefix + ".golden")
imgs = []
for k in range(nseeds):
    rp = D.SptRenderParams.from_buffer_copy(bytes(sc.params)); rp.spp = spp; rp.seed = 900 + k
    t0 = time.time()
    c, _w = O.render(sc, rp)
    imgs.append(c.astype(np.float64) / spp)
    print("oracle render %d: %.1fs" % (k, time.time() - t0), flush=True)
img = np.mean(imgs, 0)
def ratio(a, b): return a.mean((0, 1)) / b.mean((0, 1)) - 1
np.set_printoptions(precision=4, suppress=True, linewidth=200)
print("ABS reference band means:", ref.mean((0, 1))[::8], " oracle:", img.mean((0, 1))[::8])
np.set_printoptions(precision=2, suppress=True, linewidth=200)
if len(refs) > 1:
    print("ref0 vs ref1 band-mean ratio-1 (%):", 100 * ratio(refs[0], refs[1])[::4])
if len(imgs) > 1:
    print("orc0 vs orc1 band-mean ratio-1 (%):", 100 * ratio(imgs[0], imgs[1])[::4])
print("%s %dx%d %dspp: oracle(product sampler) vs reference band-mean ratio-1 (%%):" % (variant, res, res, spp), 100 * ratio(img, ref)[::4])
# where: rows of the image (top = background / teapot, bottom = floor)
rows = np.array_split(np.arange(res), 4)
for r in rows:
    print("  rows %3d-%3d: %s" % (r[0], r[-1], 100 * (img[r].mean((0, 1)) / ref[r].mean((0, 1)) - 1)[::8]))
