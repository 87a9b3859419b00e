"""One GPU renders the tile set of rank 0 of N (what each rank of an N-GPU run does per frame) under different lane /
wave settings: device render time per frame. Shows where strong scaling loses time without paying for N GPUs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pbrt_v2_spectral_b200 import capi, ctypes_defs as D, multi
from pbrt_v2_spectral_b200.scene_io import LoweredScene

root = os.path.join(os.path.dirname(__file__), "..", "..")
lowered = LoweredScene.load(os.path.join(root, "assets", "_lowered", "killeroo_path.spt"))
scene = capi.Scene(lowered)
film = capi.Film(lowered.film)
for nranks in (1, 2, 4, 8):
    base = None
    for lanes in (4, 2, 1):
        for waves_per_lane in (1, 2, 4):
            rp = D.SptRenderParams.from_buffer_copy(bytes(lowered.params))
            rp.seed = 1
            rp = multi.rank_params(rp, 0, nranks)
            scene.set_lanes(lanes)
            tiles = ((rp.x_end - rp.x_start + 31) // 32) * ((rp.y_end - rp.y_start + 31) // 32)
            local_pixels = ((tiles + nranks - 1) // nranks) * 1024
            nw = lanes * waves_per_lane
            rp.wave_pixels = (local_pixels + nw - 1) // nw
            for _ in range(3):
                scene.render(film, rp)
            ms = []
            for _ in range(12):
                scene.render(film, rp)
                ms.append(scene.stats()["render_ms"])
            t = float(np.median(ms))
            if base is None:
                base = t
            print("ranks %d lanes %d waves/lane %d (wave %7d px): %.3f ms per frame  (ideal %.3f)" % (
                nranks, lanes, waves_per_lane, rp.wave_pixels, t, 43.1 / nranks), flush=True)
    rp = D.SptRenderParams.from_buffer_copy(bytes(lowered.params)); rp.seed = 1
    rp = multi.rank_params(rp, 0, nranks); rp.wave_pixels = 0
    scene.set_lanes(4)
    for _ in range(3):
        scene.render(film, rp)
    ms = []
    for _ in range(12):
        scene.render(film, rp); ms.append(scene.stats()["render_ms"])
    print("ranks %d DEFAULT heuristic: %.3f ms per frame, lanes used %d" % (nranks, float(np.median(ms)), scene.stats()["lanes_used"]), flush=True)
