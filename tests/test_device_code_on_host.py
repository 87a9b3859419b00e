"""The product's DEVICE code (pbrt_v2_spectral_b200/csrc) compiled for the host with g++ and compared with the reference's
golden vectors and with the oracle - which is pinned bit-exactly against the reference - on a machine without a GPU: the
traversal kernel itself and the scene re-layout kernels (a warp of one lane), the kd-tree look-up of the
measured BRDF (with its 3-nearest-neighbour shortcut to the reference's final search radius), the image-texture filters
(EWA, trilinear), and the first-vertex shading frame (ray differentials, bump map, image-mapped Kd). Same compiler flags
and libm on both sides, so the comparison is bit for bit. The CUDA build of the same headers is what `-m gpu` tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM_SRC = os.path.join(ROOT, "tests", "host_shim", "device_on_host.cpp")
SHIM_SO = os.path.join(ROOT, "oracle", "_ref", "libdevhost.so")      # built artefact, next to the oracle's


def _build(so, src):
    deps = [src, os.path.join(ROOT, "tests", "host_shim", "fake", "cuda_runtime.h"), os.path.join(ROOT, "include", "spt.h")]
    csrc = os.path.join(ROOT, "pbrt_v2_spectral_b200", "csrc")
    deps += [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h", "spt_build.cu"))]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.run(["g++", "-O2", "-m64", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-std=c++17", "-w",
                        "-I" + os.path.join(ROOT, "tests", "host_shim", "fake"), "-I" + os.path.join(ROOT, "include"), "-I" + csrc,
                        "-o", so, src, "-lm"], check=True)
    return C.CDLL(so)


@pytest.fixture(scope="module")
def trace_shim():
    return _build(os.path.join(ROOT, "oracle", "_ref", "libtracehost.so"), os.path.join(ROOT, "tests", "host_shim", "trace_on_host.cpp"))


ALL_CASES = O.golden_cases()


@pytest.mark.parametrize("case", ALL_CASES, ids=[c[0] for c in ALL_CASES])
def test_traversal_kernel_and_relayout_bit_exact(trace_shim, case):
    """The traversal KERNEL source (k_trace_multi: pair-node walk, min/max slab form, leaf tests) and the re-layout kernels
    (leaf flags, pair nodes, vertex pre-gather), compiled for the host as a warp of one lane, against the reference's golden
    vectors: first-hit primitive ids and distances of camera and secondary rays bit for bit, any-hit verdicts."""
    name, sp, gp = case
    scene, g = O.load_case(sp, gp)
    rays = np.ascontiguousarray(g["rays"], np.float32)
    slot = np.empty(len(rays), np.uint32); t = np.empty(len(rays), np.float32)
    assert trace_shim.hd_trace(C.byref(scene.desc), _p(rays), len(rays), 0, _p(slot), _p(t), 0) == 0
    pid = np.where(slot == 0xffffffff, 0, scene.a["prim_id"][np.minimum(slot, len(scene.a["prim_id"]) - 1)])
    assert np.array_equal(pid, g["prim_id"])
    assert np.array_equal(t.view(np.uint32), g["t_hit"].view(np.uint32))
    m = g["prim_id"] != 0
    unbounded = np.ascontiguousarray(g["rays2"][m]); unbounded[:, 7] = np.inf
    s2 = np.empty(len(unbounded), np.uint32); t2 = np.empty(len(unbounded), np.float32)
    trace_shim.hd_trace(C.byref(scene.desc), _p(unbounded), len(unbounded), 0, _p(s2), _p(t2), 0)
    pid2 = np.where(s2 == 0xffffffff, 0, scene.a["prim_id"][np.minimum(s2, len(scene.a["prim_id"]) - 1)])
    assert np.array_equal(pid2, g["prim_id2"][m])
    assert np.array_equal(t2.view(np.uint32), g["t_hit2"][m].view(np.uint32))
    seg = np.ascontiguousarray(g["rays2"][m])
    s3 = np.empty(len(seg), np.uint32)
    trace_shim.hd_trace(C.byref(scene.desc), _p(seg), len(seg), 1, _p(s3), None, 0)
    assert np.array_equal((s3 != 0xffffffff).astype(np.uint8), g["any2"][m])


@pytest.mark.parametrize("case", ALL_CASES, ids=[c[0] for c in ALL_CASES])
def test_fast_traversal_differs_only_on_ties(trace_shim, case):
    """The FAST layout (csrc/wide.h: 4-wide nodes collapsed from the reference's flattened tree, children entered nearest
    first) against the reference's golden first hits: where the closest hit is unique it must be the same primitive at the
    bit-identical distance. It may differ only where two primitives lie at (nearly) the same distance - the reference keeps the
    last one it tests, a nearest-first walk another - or where the reference's own box test drops a grazing hit; both are
    counted and bounded. Any-hit verdicts do not depend on the order at all."""
    name, sp, gp = case
    scene, g = O.load_case(sp, gp)
    m = g["prim_id"] != 0
    unbounded = np.ascontiguousarray(g["rays2"][m]); unbounded[:, 7] = np.inf
    for rays, want_id, want_t in ((np.ascontiguousarray(g["rays"], np.float32), g["prim_id"], g["t_hit"]), (unbounded, g["prim_id2"][m], g["t_hit2"][m])):
        slot = np.empty(len(rays), np.uint32); t = np.empty(len(rays), np.float32)
        assert trace_shim.hd_trace(C.byref(scene.desc), _p(rays), len(rays), 0, _p(slot), _p(t), 1) == 0
        pid = np.where(slot == 0xffffffff, 0, scene.a["prim_id"][np.minimum(slot, len(scene.a["prim_id"]) - 1)])
        same = (pid == want_id) & (t.view(np.uint32) == want_t.view(np.uint32))
        diff = ~same
        # a difference must be a tie (same distance to 1e-5 relative, another primitive) or a hit the reference's box test dropped
        with np.errstate(invalid="ignore"):
            tie = diff & (pid != 0) & (want_id != 0) & (np.abs(t - want_t) <= 1e-5 * np.abs(want_t))
            found_more = diff & (pid != 0) & ((want_id == 0) | (t < want_t))
        if diff.any():
            print("%s: fast traversal differs on %d of %d rays (%d ties, %d the reference's box test dropped)" % (name, diff.sum(), len(diff), tie.sum(), found_more.sum()))
        assert not (diff & ~tie & ~found_more).any(), "%s: fast traversal LOST hits" % name
        assert diff.mean() < 2e-3, "%s: %d of %d rays differ" % (name, diff.sum(), len(diff))
    seg = np.ascontiguousarray(g["rays2"][m])
    s3 = np.empty(len(seg), np.uint32)
    trace_shim.hd_trace(C.byref(scene.desc), _p(seg), len(seg), 1, _p(s3), None, 1)
    anyhit = (s3 != 0xffffffff).astype(np.uint8)
    # an any-hit ray may only gain a hit the reference's box test dropped
    assert not ((anyhit == 0) & (g["any2"][m] == 1)).any()
    assert (anyhit != g["any2"][m]).mean() < 1e-3


@pytest.fixture(scope="module")
def shim():
    return _build(SHIM_SO, SHIM_SRC)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _unit(rng, n, upper=True):
    v = rng.normal(size=(n, 3)).astype(np.float32)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    if upper:
        v[:, 2] = np.abs(v[:, 2])
    return np.ascontiguousarray(v, np.float32)


def _case(name):
    sp, gp = os.path.join(O.GOLDEN_BIG, name + ".spt"), os.path.join(O.GOLDEN_BIG, name + ".golden")
    if not os.path.exists(sp):
        sp, gp = os.path.join(O.GOLDEN_SMALL, name + ".spt"), os.path.join(O.GOLDEN_SMALL, name + ".golden")
    if not os.path.exists(sp):
        pytest.skip("golden set %s not generated" % name)
    return O.load_case(sp, gp)


def test_measured_brdf_lookup_bit_exact(shim):
    scene, _ = _case("bunny_measured_small")
    rng = np.random.default_rng(3)
    n = 4000
    wo, wi = _unit(rng, n), _unit(rng, n, upper=False)
    wi[:100] = wo[:100] * np.float32(-1.0)                     # degenerate pairs (wo + wi = 0 in the remap)
    wi[100:200, 2] = 0.0                                       # grazing
    a = np.empty((n, O.D.NBANDS), np.float32); b = np.empty_like(a)
    shim.hd_measured_f(C.byref(scene.desc), 0, _p(wo), _p(wi), n, _p(a))
    O.lib().orc_measured_f(C.byref(scene.desc), 0, _p(wo), _p(wi), n, _p(b))
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.isfinite(b).mean() > 0.9 and b[np.isfinite(b)].max() > 0


@pytest.mark.parametrize("name", ["tiny_tex", "metal_shipped_small"])
def test_image_texture_filters_bit_exact(shim, name):
    scene, _ = _case(name)
    rng = np.random.default_rng(5)
    for tex in range(scene.desc.n_textures):
        n = 3000
        uvd = np.empty((n, 6), np.float32)
        uvd[:, :2] = rng.uniform(-1.5, 2.5, (n, 2))
        scale = 10.0 ** rng.uniform(-5, -0.5, (n, 1))          # footprints from far below a texel to many texels
        uvd[:, 2:] = rng.normal(size=(n, 4)) * scale
        uvd[:300, 2:] = 0.0                                    # no differentials: level-0 bilinear (later path vertices)
        uvd[300:400, 4:] = 0.0                                 # degenerate minor axis
        ch = int(np.frombuffer(scene.a["textures"].tobytes(), np.int32).reshape(-1, 16)[tex, 0])
        a = np.empty((n, ch), np.float32); b = np.empty_like(a)
        shim.hd_tex_evaluate(C.byref(scene.desc), tex, _p(uvd), n, _p(a))
        O.lib().orc_tex_evaluate(C.byref(scene.desc), tex, _p(uvd), n, _p(b))
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), "texture %d of %s" % (tex, name)


@pytest.mark.parametrize("name", ["tiny_tex", "metal_shipped_small", "killeroo_small"])
def test_first_vertex_frame_bit_exact(shim, name):
    """camera_ray_diff -> shape_record -> compute_differentials -> Material::Bump / Kd look-up -> BSDF frame."""
    scene, g = _case(name)
    slot, pid, t = O.trace_closest(scene, g["rays"])
    hit = np.flatnonzero(slot != 0xffffffff)[:4000]
    smp = np.ascontiguousarray(g["samples"][hit, :5], np.float32)
    s = np.ascontiguousarray(slot[hit]); tt = np.ascontiguousarray(t[hit])
    spp = int(g["meta"][0])
    a = np.zeros((len(hit), 12), np.float32); b = np.zeros_like(a)
    shim.hd_first_vertex_frame(C.byref(scene.desc), C.byref(scene.camera), spp, _p(smp), _p(s), _p(tt), len(hit), _p(a))
    O.lib().orc_first_vertex_frame(C.byref(scene.desc), C.byref(scene.camera), spp, _p(smp), _p(s), _p(tt), len(hit), _p(b))
    assert np.any(b != 0)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("name", ["tiny", "killeroo_small", "envmap_small", "bunny_small"])
def test_light_sampling_bit_exact(shim, name):
    """Light::Sample_L and Light::Pdf of every light (sphere / disk / triangle-set area lights, point lights, the
    importance-sampled environment map) from first-hit points of the golden camera rays."""
    scene, g = _case(name)
    slot, pid, t = O.trace_closest(scene, g["rays"])
    hit = np.flatnonzero(slot != 0xffffffff)[:3000]
    r = g["rays"][hit]
    p = np.ascontiguousarray(r[:, :3] + r[:, 3:6] * t[hit, None], np.float32)
    rng = np.random.default_rng(9)
    u = rng.random((len(hit), 3)).astype(np.float32)
    # a third of the samples at the rim of the cone Sphere::Sample draws from (u1 -> 0, sphere.cpp:228-252): there the float
    # quadratic of Sphere::Intersect misses the sphere by rounding, Sample falls back to the point of closest approach, and
    # what ShapeSet::Sample's re-intersection (light.cpp:141-149) makes of that point decides whether the sample is black
    u[::3, 0] *= 0.02
    w = _unit(rng, len(hit), upper=False)
    for light in range(scene.desc.n_lights):
        a = np.zeros((len(hit), 9), np.float32); b = np.zeros_like(a)
        shim.hd_light_sample(C.byref(scene.desc), light, _p(p), _p(u), len(hit), _p(a))
        O.lib().orc_light_sample(C.byref(scene.desc), light, _p(p), _p(u), len(hit), _p(b))
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), "Sample_L of light %d of %s" % (light, name)
        # Light::Pdf towards the sampled directions (non-zero for area / infinite lights) and towards random ones
        for dirs in (np.ascontiguousarray(b[:, :3]), w):
            pa = np.zeros(len(hit), np.float32); pb = np.zeros_like(pa)
            shim.hd_light_pdf(C.byref(scene.desc), light, _p(p), _p(dirs), len(hit), _p(pa))
            O.lib().orc_light_pdf(C.byref(scene.desc), light, _p(p), _p(dirs), len(hit), _p(pb))
            assert np.array_equal(pa.view(np.uint32), pb.view(np.uint32)), "Pdf of light %d of %s" % (light, name)


def test_generated_samples_match_the_restated_sampler(trace_shim):
    """The product's own sampler (hashed (0,2)-sequences, csrc/sampler.cuh) and K1's sample -> film position -> camera ray
    chain, run on the host, against the oracle's restatement of it (orc_gen_sample): the whole-render tests on the GPU
    compare images made by the two and rely on the samples being the same numbers."""
    scene, _ = _case("tiny")
    spp, seed = 8, 1234567
    n = 64 * spp
    xy = np.empty((n, 2), np.float32); rays = np.empty((n, 8), np.float32)
    trace_shim.hd_gen_tile(C.byref(scene.camera), C.c_uint64(seed), 16, 8, spp, _p(xy), _p(rays))
    px, py = np.floor(xy[:, 0]).astype(int), np.floor(xy[:, 1]).astype(int)
    assert px.min() == 16 and px.max() == 23 and py.min() == 8 and py.max() == 15
    u11 = np.empty(11, np.float32)
    for i in range(0, n, 7):
        s = i & (spp - 1)
        smp, rng = O.gen_samples(seed, int(px[i]), int(py[i]), spp, shutter=(scene.camera.shutter_open, scene.camera.shutter_close), n_rng=34)
        assert np.array_equal(smp[s, :2].view(np.uint32), xy[i].view(np.uint32))
        ref_ray = O.camera_rays(scene, smp[s:s + 1, :5])[0]
        assert np.array_equal(ref_ray.view(np.uint32), rays[i].view(np.uint32))
        for b in range(6):
            trace_shim.hd_bounce_dims(C.c_uint64(seed), int(px[i]), int(py[i]), s, spp, b, 1, _p(u11))
            if b < 3:
                one, two = smp[s, 5 + 4 * b:9 + 4 * b], smp[s, 19 + 6 * b:25 + 6 * b]
                want = np.array([one[1], two[0], two[1], one[0], two[2], two[3], one[2], two[4], two[5], one[3]], np.float32)
                assert np.array_equal(u11[:10].view(np.uint32), want.view(np.uint32))
            else:
                k = (b - 3) * 10 + (b - 4 if b > 4 else 0)          # RNG draws consumed before bounce b (path.cpp:82,97)
                assert np.array_equal(u11[:10].view(np.uint32), rng[s, k:k + 10].view(np.uint32))
                assert u11[10].view(np.uint32) == rng[s, k + 10].view(np.uint32)


ASAN_CHILD = r'''
import sys, ctypes as C
sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
import numpy as np, oracle_lib as O
lib = C.CDLL(%(so)r)
p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
n = 0
for name, sp, gp in O.golden_cases():
    scene, g = O.load_case(sp, gp)
    rays = np.ascontiguousarray(g["rays"], np.float32)
    m = g["prim_id"] != 0
    seg = np.ascontiguousarray(g["rays2"][m])
    for wide in (0, 1):
        slot = np.empty(len(rays), np.uint32); t = np.empty(len(rays), np.float32)
        assert lib.hd_trace(C.byref(scene.desc), p(rays), len(rays), 0, p(slot), p(t), wide) == 0
        s3 = np.empty(len(seg), np.uint32)
        lib.hd_trace(C.byref(scene.desc), p(seg), len(seg), 1, p(s3), None, wide)
    n += 1
print("asan ok", n)
'''


def test_traversal_kernels_under_address_sanitizer(tmp_path):
    """compute-sanitizer is closed on this GPU pool (profiles/r02_sanitizer_memcheck.log holds the refusal), so the memory
    safety of the traversal code is checked where it can be: the kernel source (k_trace_multi, exact and wide), the re-layout
    kernels and the wide-tree builder compiled for the host with -fsanitize=address,undefined and run over every golden set -
    out-of-bounds node / stack / vertex accesses and undefined shifts would abort the child."""
    import sys
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan not available")
    so = str(tmp_path / "libtracehost_asan.so")
    csrc = os.path.join(ROOT, "pbrt_v2_spectral_b200", "csrc")
    subprocess.run(["g++", "-O1", "-g", "-m64", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-std=c++17", "-w",
                    "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer",
                    "-I" + os.path.join(ROOT, "tests", "host_shim", "fake"), "-I" + os.path.join(ROOT, "include"), "-I" + csrc,
                    "-o", so, os.path.join(ROOT, "tests", "host_shim", "trace_on_host.cpp"), "-lm"], check=True)
    code = ASAN_CHILD % {"root": ROOT, "tests": os.path.join(ROOT, "tests"), "so": so}
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0"))
    assert r.returncode == 0 and "asan ok" in r.stdout, (r.stdout + r.stderr)[-3000:]
