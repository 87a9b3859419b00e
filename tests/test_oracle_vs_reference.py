"""Pins the plain-C oracle (oracle/spt_oracle.c) against vectors produced by the UNMODIFIED
reference (oracle/oracle_dump.cpp linked with the reference objects): camera rays, first hits
(bit-exact ids, exact t), secondary closest/any-hit rays, and per-sample path-traced radiance
for the reference's own LDPixelSample vectors and RNG draws."""
import numpy as np
import pytest

import oracle_lib as O

CASES = O.golden_cases()


@pytest.fixture(scope="module", params=CASES, ids=[c[0] for c in CASES])
def case(request):
    name, sp, gp = request.param
    scene, g = O.load_case(sp, gp)
    return name, scene, g


def test_camera_rays_bit_exact(case):
    _, scene, g = case
    rays = O.camera_rays(scene, g["samples"][:, :5])
    # K1 parity: same fp32 operations in the same order -> identical bits
    assert np.array_equal(rays.view(np.uint32), g["rays"].view(np.uint32))


def test_first_hit_ids_and_t(case):
    _, scene, g = case
    slot, pid, t = O.trace_closest(scene, g["rays"])
    assert np.array_equal(pid, g["prim_id"])          # bit-exact ids (north_star)
    assert np.array_equal(t.view(np.uint32), g["t_hit"].view(np.uint32))


def test_secondary_rays(case):
    _, scene, g = case
    m = g["prim_id"] != 0
    unbounded = g["rays2"][m].copy()
    unbounded[:, 7] = np.inf      # the dump traced the closest hit with maxt = INFINITY, any-hit with the segment
    slot, pid, t = O.trace_closest(scene, unbounded)
    assert np.array_equal(pid, g["prim_id2"][m])
    assert np.array_equal(t.view(np.uint32), g["t_hit2"][m].view(np.uint32))
    hit = O.trace_any(scene, g["rays2"][m])
    assert np.array_equal(hit, g["any2"][m])


def test_path_radiance(case):
    name, scene, g = case
    if "L" not in g:
        pytest.skip("no radiance in this golden set")
    L = O.shade_samples(scene, g["samples"], g["rng"])
    ref = g["L"]
    # same libm, same fp32 operation order as the reference build: bit-exact radiance
    assert np.array_equal(L.view(np.uint32), ref.view(np.uint32)), "%s: %d of %d samples differ" % (
        name, (L != ref).any(axis=1).sum(), len(L))


RAY_CASES = O.ray_cases()


@pytest.mark.parametrize("rcase", RAY_CASES, ids=[c[0] for c in RAY_CASES])
def test_first_hits_at_baseline_resolution(rcase):
    """BASELINE configs 1 (700x700) and 2 (640x480) at full resolution: the reference's own first hits of 65 536 (committed)
    and 1 048 576 camera rays - ids and distances bit for bit."""
    name, sp, gp = rcase
    scene, g = O.load_case(sp, gp)
    assert scene.camera.lens_radius == 0.0
    rays = O.camera_rays(scene, O.compact_samples(g))
    slot, pid, t = O.trace_closest(scene, rays)
    assert len(pid) >= (1 << 20 if not name.endswith("_mid") else 1 << 16)
    assert np.array_equal(pid, g["prim_id"])
    assert np.array_equal(t.view(np.uint32), g["t_hit"].view(np.uint32))
    assert (pid != 0).mean() > 0.2          # not a set of misses
