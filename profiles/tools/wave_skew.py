"""Two waves of unequal size on the two lanes: the drains of one lane's persistent trace launches then fall into the body of the
other lane's kernels instead of coinciding with its drains. One GPU renders rank 0's tile set of N GPUs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pbrt_v2_spectral_b200 import capi, ctypes_defs as D, multi
from pbrt_v2_spectral_b200.scene_io import LoweredScene

root = os.path.join(os.path.dirname(__file__), "..", "..")
lowered = LoweredScene.load(os.path.join(root, "assets", "_lowered", "killeroo_path.spt"))
scene = capi.Scene(lowered)
film = capi.Film(lowered.film)
for nranks in (8, 4, 1):
    rp0 = D.SptRenderParams.from_buffer_copy(bytes(lowered.params)); rp0.seed = 1
    rp0 = multi.rank_params(rp0, 0, nranks)
    ntiles = 22 * 22
    local = ((ntiles + nranks - 1) // nranks) * 1024
    for first in (0, 50, 53, 56, 60, 65, 70, 100):
        rp = D.SptRenderParams.from_buffer_copy(bytes(rp0))
        rp.wave_pixels = 0 if first == 0 else (local * first + 99) // 100
        for _ in range(3): scene.render(film, rp)
        ms = []
        for _ in range(12):
            scene.render(film, rp); ms.append(scene.stats()["render_ms"])
        print("N=%d first wave %3d %% of the pixels (wave_pixels %d): %.3f ms" % (nranks, first, rp.wave_pixels, float(np.median(ms))), flush=True)
