"""Where the end-to-end (host-buffer) time of one render goes. Run on the GPU box."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pbrt_v2_spectral_b200 import capi, ctypes_defs as D
from pbrt_v2_spectral_b200.scene_io import LoweredScene

lowered = LoweredScene.load(os.path.join(os.path.dirname(__file__), "..", "..", "assets", "_lowered", "killeroo_path.spt"))
rp = D.SptRenderParams.from_buffer_copy(bytes(lowered.params))
fd = lowered.film
cp = capi.HostBuffer((fd.y_pixel_count, fd.x_pixel_count, D.NBANDS)); wp = capi.HostBuffer((fd.y_pixel_count, fd.x_pixel_count))
for it in range(6):
    t = [time.perf_counter()]
    sc = capi.Scene(lowered); t.append(time.perf_counter())
    f = capi.Film(fd); t.append(time.perf_counter())
    sc.render(f, rp); t.append(time.perf_counter())
    dev_ms = sc.stats()["render_ms"]
    f.download((cp.array, wp.array)); t.append(time.perf_counter())
    f.close(); sc.close(); t.append(time.perf_counter())
    names = ["scene_create", "film_create", "render(host wall)", "film_download", "destroy"]
    print("iter %d: " % it + ", ".join("%s %.1f ms" % (n, 1e3 * (b - a)) for n, a, b in zip(names, t, t[1:])),
          "| device render %.1f ms | timed total %.1f ms" % (dev_ms, 1e3 * (t[4] - t[0])))
