"""One GPU renders, in turn, the tile set of every rank of an 8-GPU run for two tile sizes: device time per rank.
The slowest rank sets the frame time of the real run; shows what the tile size does to the balance."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pbrt_v2_spectral_b200 import capi, ctypes_defs as D, multi
from pbrt_v2_spectral_b200.scene_io import LoweredScene

root = os.path.join(os.path.dirname(__file__), "..", "..")
lowered = LoweredScene.load(os.path.join(root, "assets", "_lowered", "killeroo_path.spt"))
scene = capi.Scene(lowered)
film = capi.Film(lowered.film)
for nranks in (8, 4):
    for tile in (32, 16, 8):
        times = []
        for r in range(nranks):
            rp = D.SptRenderParams.from_buffer_copy(bytes(lowered.params)); rp.seed = 1
            rp = multi.rank_params(rp, r, nranks, tile)
            for _ in range(2):
                scene.render(film, rp)
            ms = []
            for _ in range(8):
                scene.render(film, rp); ms.append(scene.render_ms())
            times.append(float(np.median(ms)))
        print("ranks %d tile %2d: per-rank ms %s  max %.3f  mean %.3f  (ideal %.3f)" % (
            nranks, tile, " ".join("%.2f" % t for t in times), max(times), sum(times) / len(times), 42.65 / nranks), flush=True)
