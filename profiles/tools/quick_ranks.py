"""One GPU renders the tile set of rank 0 of N = 1, 2, 4, 8 with the default wave / lane heuristics: the per-rank device
time an N-GPU run can reach. Env knobs (SPT_MERGE_TRACE, SPT_STACK_SMEM, SPT_FETCH_THRESHOLD, SPT_LANES) are read by the library."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pbrt_v2_spectral_b200 import capi, ctypes_defs as D, multi
from pbrt_v2_spectral_b200.scene_io import LoweredScene

root = os.path.join(os.path.dirname(__file__), "..", "..")
wl = sys.argv[1] if len(sys.argv) > 1 else "killeroo_path"
lowered = LoweredScene.load(os.path.join(root, "assets", "_lowered", wl + ".spt"))
scene = capi.Scene(lowered)
film = capi.Film(lowered.film)
out = {}
for nranks in (1, 2, 4, 8):
    rp = D.SptRenderParams.from_buffer_copy(bytes(lowered.params)); rp.seed = 1
    rp = multi.rank_params(rp, 0, nranks); rp.wave_pixels = 0
    for _ in range(3):
        scene.render(film, rp)
    ms = []
    for _ in range(10):
        scene.render(film, rp); ms.append(scene.stats()["render_ms"])
    out[nranks] = float(np.median(ms))
base = out[1]
print(wl, {k: os.environ.get(k) for k in ("SPT_MERGE_TRACE", "SPT_STACK_SMEM", "SPT_FETCH_THRESHOLD", "SPT_LANES") if os.environ.get(k)},
      " ".join("N=%d %.3f ms (eff %.3f)" % (n, t, base / n / t) for n, t in out.items()), flush=True)
