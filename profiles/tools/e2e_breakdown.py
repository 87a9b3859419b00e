"""Where the end-to-end (host-buffer) time of one render goes. Run on the GPU box."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pbrt_v2_spectral_b200 import capi, ctypes_defs as D
from pbrt_v2_spectral_b200.scene_io import LoweredScene

lowered = LoweredScene.load(os.path.join(os.path.dirname(__file__), "..", "..", "assets", "_lowered", "killeroo_path.spt"))
rp = D.SptRenderParams.from_buffer_copy(bytes(lowered.params))
fd = lowered.film
c = np.empty((fd.y_pixel_count, fd.x_pixel_count, D.NBANDS), np.float32)
w = np.empty((fd.y_pixel_count, fd.x_pixel_count), np.float32)
for it in range(3):
    t = [time.perf_counter()]
    sc = capi.Scene(lowered); t.append(time.perf_counter())
    f = capi.Film(fd); t.append(time.perf_counter())
    sc.render(f, rp); t.append(time.perf_counter())
    f.download((c, w)); t.append(time.perf_counter())
    f.close(); sc.close(); t.append(time.perf_counter())
    names = ["scene_create", "film_create", "render(host wall)", "film_download", "destroy"]
    print("iter %d: " % it + ", ".join("%s %.1f ms" % (n, 1e3 * (b - a)) for n, a, b in zip(names, t, t[1:])),
          "| device render %.1f ms" % sc_stats["render_ms"] if False else "")
