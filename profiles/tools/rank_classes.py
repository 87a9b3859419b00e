"""Per-class kernel times of rank 0's tile set of N GPUs (one GPU, every wave on one stream so that each kernel sits between
its own events) beside 1/N of the whole frame's: where the per-rank time of an N-GPU run goes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pbrt_v2_spectral_b200 import capi, ctypes_defs as D, multi
from pbrt_v2_spectral_b200.scene_io import LoweredScene

root = os.path.join(os.path.dirname(__file__), "..", "..")
lowered = LoweredScene.load(os.path.join(root, "assets", "_lowered", "killeroo_path.spt"))
names = ["gen", "trace", "shade", "addlight", "film", "advance"]
res = {}
for lanes in (1, 2):
    scene = capi.Scene(lowered); scene.set_lanes(lanes)
    film = capi.Film(lowered.film)
    for nranks in (1, 8):
        rp = D.SptRenderParams.from_buffer_copy(bytes(lowered.params)); rp.seed = 1
        rp = multi.rank_params(rp, 0, nranks); rp.wave_pixels = 0
        for _ in range(3): scene.render(film, rp)
        tot = []; cls = np.zeros(8); launches = None
        for _ in range(8):
            scene.render(film, rp); st = scene.stats()
            tot.append(st["render_ms"]); cls[:len(st["class_ms"])] += np.array(st["class_ms"]); launches = st["class_launches"]
        res[(lanes, nranks)] = (float(np.median(tot)), cls / 8, launches)
    film.close(); scene.close()
for (lanes, nranks), (t, cls, launches) in res.items():
    print("lanes %d N=%d render %.3f ms  classes(ms) %s sum %.3f  launches %s" % (lanes, nranks, t, np.round(cls[:6], 3).tolist(), cls[:6].sum(), list(launches)[:6]))
t1, c1, _ = res[(1, 1)]; t8, c8, _ = res[(1, 8)]
print("one stream: class time at N=8 minus 1/8 of the full frame's:", np.round(c8[:6] - c1[:6] / 8, 3).tolist(), "total", round(float((c8[:6] - c1[:6] / 8).sum()), 3))
