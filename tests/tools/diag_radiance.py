"""TEST INFRASTRUCTURE (uses the golden vectors through tests/oracle_lib.py). GPU-box diagnostic: per-sample radiance of the CUDA path for the reference's own sample vectors, saved next to
the reference values so disagreements can be classified offline (which material, which bounce)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # tests/tools/ -> repo root
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
from pbrt_v2_spectral_b200 import capi

names = sys.argv[1:] or ["metal_shipped_small", "ssenv_shipped_small"]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
for name in names:
    lowered, g = O.load_case(os.path.join(O.GOLDEN_BIG, name + ".spt"), os.path.join(O.GOLDEN_BIG, name + ".golden"))
    scene = capi.Scene(lowered)
    out = {}
    for md in [int(x) for x in os.environ.get("DIAG_MD", "0,1,5").split(",")]:
        out["L_md%d" % md] = scene.shade_samples(g["samples"], g["rng"], max_depth=md)
    scene.close()
    ref = g["L"]
    L = out["L_md%d" % lowered.params.max_depth] if ("L_md%d" % lowered.params.max_depth) in out else out["L_md5"]
    err = np.abs(L - ref).max(axis=1) / np.maximum(np.abs(ref).max(axis=1), 1e-6)
    print(name, "bad", int((err > 2e-4).sum()), "of", len(err), "worst", float(err.max()))
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "diag_%s.npz" % name), **out)
