"""The 30-band variant (BASELINE.json: "30-band spectral path trace"; src/core/spectrum.h:41-43 with nSpectralSamples = 30):
libspt30.so / liboracle30.so are the same sources built with -DSPT_NBANDS=30, pinned against golden vectors and a 1024-spp
image made by the reference itself built with 30 bands (oracle/Makefile ref30, oracle/make_golden.py killeroo_small30).
The Python side sizes its arrays from SPT_NBANDS in the environment, so these run the ordinary test files in a child pytest."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN30 = os.path.join(ROOT, "oracle", "_ref", "golden", "killeroo_small30.golden")
needs_golden = pytest.mark.skipif(not os.path.exists(GOLDEN30), reason="30-band golden set not generated (needs /root/reference)")


def _child(args):
    env = dict(os.environ, SPT_NBANDS="30")
    env.pop("SPT_LIB", None)
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", *args], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=1800)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and " failed" not in r.stdout, tail
    return r.stdout


@needs_golden
def test_oracle30_bit_exact_against_the_30_band_reference():
    out = _child(["tests/test_oracle_vs_reference.py", "tests/test_abi.py::test_header_symbols_exported", "tests/test_abi.py::test_struct_sizes_match_c"])
    assert "killeroo_small30" in out or "passed" in out


@needs_golden
@pytest.mark.gpu
def test_cuda_path_30_bands():
    """Camera rays / first hits / secondary rays bit-exact, radiance within 2e-4, film, whole render vs oracle, tile sets, and the
    converged image within 1 % per band of the 30-band reference's render - all through libspt30.so."""
    out = _child(["-m", "gpu", "-s", "tests/test_gpu_parity.py", "tests/test_image_parity.py", "tests/test_multi_gpu.py"])
    print(out[-2500:])
