"""ctypes binding of the C ABI in include/spt.h (libspt.so). This is the Python-side host of the
drop-in boundary: the same calls the C++ `Renderer "gpupath"` makes
(pbrt_v2_spectral_b200/host/gpupath.cpp). There is no CPU fallback here or in the library: a
missing libspt.so or a missing CUDA device raises."""
import ctypes as C
import os

import numpy as np

from . import ctypes_defs as D

HERE = os.path.dirname(os.path.abspath(__file__))
# libspt.so: 32 bands, the reference as shipped; libspt30.so: the 30-band variant (SPT_NBANDS=30 in the environment selects it
# and sizes the Python-side arrays). SPT_LIB: A/B builds of the same ABI (profiles/tools)
LIB_PATH = os.environ.get("SPT_LIB") or os.path.join(HERE, "libspt.so" if D.NBANDS == 32 else "libspt%d.so" % D.NBANDS)

# every symbol include/spt.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "spt_nbands", "spt_last_error", "spt_device_count", "spt_set_device", "spt_host_alloc", "spt_host_free", "spt_trim",
    "spt_scene_create", "spt_scene_destroy", "spt_scene_set_traversal", "spt_scene_enable_counters", "spt_scene_set_lanes", "spt_get_stats", "spt_last_render_ms",
    "spt_camera_rays", "spt_trace_closest", "spt_trace_any", "spt_trace_closest_dev", "spt_trace_any_dev",
    "spt_shade_samples",
    "spt_film_create", "spt_film_create_external", "spt_film_destroy", "spt_film_clear", "spt_film_clear_idle",
    "spt_film_add_samples", "spt_film_download", "spt_film_device_ptr", "spt_film_write_dat",
    "spt_render", "spt_render_begin", "spt_render_end",
    "spt_film_ipc_export", "spt_film_open_ipc",
    "spt_multi_create", "spt_multi_destroy", "spt_multi_device_count", "spt_multi_render", "spt_multi_film", "spt_multi_get_stats",
    "spt_multi_last_render_ms",
]
IPC_HANDLE_BYTES = 64


class SptError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SptError("libspt.so is not built (python -m pbrt_v2_spectral_b200.build); there is no fallback path")
        L = C.CDLL(LIB_PATH)
        L.spt_last_error.restype = C.c_char_p
        L.spt_host_alloc.restype = C.c_void_p
        L.spt_host_alloc.argtypes = [C.c_uint64]
        L.spt_host_free.argtypes = [C.c_void_p]
        L.spt_trim.restype = None
        L.spt_scene_create.restype = C.c_void_p
        L.spt_scene_create.argtypes = [C.POINTER(D.SptSceneDesc)]
        L.spt_scene_destroy.argtypes = [C.c_void_p]
        L.spt_scene_enable_counters.argtypes = [C.c_void_p, C.c_int]
        L.spt_scene_set_traversal.argtypes = [C.c_void_p, C.c_int]
        L.spt_scene_set_lanes.argtypes = [C.c_void_p, C.c_int]
        L.spt_get_stats.argtypes = [C.c_void_p, C.POINTER(D.SptStats)]
        L.spt_last_render_ms.restype = C.c_double
        L.spt_last_render_ms.argtypes = [C.c_void_p]
        L.spt_camera_rays.argtypes = [C.POINTER(D.SptCameraDesc), C.c_void_p, C.c_uint64, C.c_void_p]
        L.spt_trace_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.spt_trace_any.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.spt_trace_closest_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.spt_trace_any_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.spt_shade_samples.argtypes = [C.c_void_p, C.POINTER(D.SptCameraDesc), C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                        C.c_int32, C.c_uint64, C.c_void_p]
        L.spt_film_create.restype = C.c_void_p
        L.spt_film_create.argtypes = [C.POINTER(D.SptFilmDesc)]
        L.spt_film_create_external.restype = C.c_void_p
        L.spt_film_create_external.argtypes = [C.POINTER(D.SptFilmDesc), C.c_void_p]
        L.spt_film_destroy.argtypes = [C.c_void_p]
        L.spt_film_clear.argtypes = [C.c_void_p]
        L.spt_film_clear_idle.argtypes = [C.c_void_p]
        L.spt_film_add_samples.argtypes = [C.c_void_p, C.POINTER(D.SptSpectralTables), C.c_void_p, C.c_void_p, C.c_uint64]
        L.spt_film_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.spt_film_device_ptr.restype = C.c_void_p
        L.spt_film_device_ptr.argtypes = [C.c_void_p]
        L.spt_film_write_dat.argtypes = [C.c_void_p, C.c_char_p]
        L.spt_render.argtypes = [C.c_void_p, C.POINTER(D.SptCameraDesc), C.c_void_p, C.POINTER(D.SptRenderParams)]
        L.spt_render_begin.argtypes = [C.c_void_p, C.POINTER(D.SptCameraDesc), C.c_void_p, C.POINTER(D.SptRenderParams)]
        L.spt_render_end.argtypes = [C.c_void_p]
        L.spt_film_ipc_export.argtypes = [C.c_void_p, C.c_void_p]
        L.spt_film_open_ipc.restype = C.c_void_p
        L.spt_film_open_ipc.argtypes = [C.POINTER(D.SptFilmDesc), C.c_void_p]
        L.spt_multi_create.restype = C.c_void_p
        L.spt_multi_create.argtypes = [C.POINTER(D.SptSceneDesc), C.POINTER(D.SptFilmDesc), C.c_int, C.c_void_p]
        L.spt_multi_destroy.argtypes = [C.c_void_p]
        L.spt_multi_destroy.restype = None
        L.spt_multi_device_count.argtypes = [C.c_void_p]
        L.spt_multi_render.argtypes = [C.c_void_p, C.POINTER(D.SptCameraDesc), C.POINTER(D.SptRenderParams)]
        L.spt_multi_film.restype = C.c_void_p
        L.spt_multi_film.argtypes = [C.c_void_p]
        L.spt_multi_get_stats.argtypes = [C.c_void_p, C.c_int, C.POINTER(D.SptStats)]
        L.spt_multi_last_render_ms.restype = C.c_double
        L.spt_multi_last_render_ms.argtypes = [C.c_void_p]
        if L.spt_nbands() != D.NBANDS:
            raise SptError("libspt.so was built for %d bands, the Python side expects %d" % (L.spt_nbands(), D.NBANDS))
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise SptError("spt error %d: %s" % (rc, (lib().spt_last_error() or b"").decode()))


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def device_count():
    return lib().spt_device_count()


def set_device(i):
    _check(lib().spt_set_device(int(i)))


class HostBuffer:
    """Page-locked host array (spt_host_alloc) viewed as numpy; free with close()."""

    def __init__(self, shape, dtype=np.float32):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = lib().spt_host_alloc(n)
        if not self.ptr:
            raise SptError("spt_host_alloc: " + (lib().spt_last_error() or b"").decode())
        buf = (C.c_char * n).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def close(self):
        if self.ptr:
            self.array = None
            lib().spt_host_free(self.ptr)
            self.ptr = None


def trim():
    lib().spt_trim()


def camera_rays(camera, samples5):
    s = np.ascontiguousarray(samples5, np.float32)
    out = np.empty((len(s), 8), np.float32)
    _check(lib().spt_camera_rays(C.byref(camera), _p(s), len(s), _p(out)))
    return out


class Scene:
    """Device-resident scene (spt_scene_create). `lowered` is a scene_io.LoweredScene."""

    def __init__(self, lowered):
        self.lowered = lowered
        self.h = lib().spt_scene_create(C.byref(lowered.desc))
        if not self.h:
            raise SptError("spt_scene_create: " + (lib().spt_last_error() or b"").decode())

    def close(self):
        if self.h:
            lib().spt_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def enable_counters(self, on=True):
        _check(lib().spt_scene_enable_counters(self.h, 1 if on else 0))

    def set_lanes(self, lanes):
        _check(lib().spt_scene_set_lanes(self.h, int(lanes)))

    def render_ms(self):
        """Device time of the last spt_render (first launch -> film final), without folding the per-kernel times."""
        return float(lib().spt_last_render_ms(self.h))

    def stats(self):
        st = D.SptStats()
        _check(lib().spt_get_stats(self.h, C.byref(st)))
        out = {}
        for k, _ in D.SptStats._fields_:
            v = getattr(st, k)
            out[k] = list(v) if hasattr(v, "__len__") else v
        return out

    def set_traversal(self, fast):
        """fast = False: the bit-exact pair-node walk (default); True: the 4-wide fast layout (include/spt.h)."""
        _check(lib().spt_scene_set_traversal(self.h, 1 if fast else 0))

    def trace_closest(self, rays):
        r = np.ascontiguousarray(rays, np.float32)
        n = len(r)
        slot = np.empty(n, np.uint32); pid = np.empty(n, np.uint32); t = np.empty(n, np.float32)
        _check(lib().spt_trace_closest(self.h, _p(r), n, _p(slot), _p(pid), _p(t)))
        return slot, pid, t

    def trace_any(self, rays):
        r = np.ascontiguousarray(rays, np.float32)
        hit = np.empty(len(r), np.uint8)
        _check(lib().spt_trace_any(self.h, _p(r), len(r), _p(hit)))
        return hit

    def trace_closest_dev(self, rays_ptr, n, slot_ptr, t_ptr):
        _check(lib().spt_trace_closest_dev(self.h, rays_ptr, n, slot_ptr, t_ptr))

    def trace_any_dev(self, rays_ptr, n, hit_ptr):
        _check(lib().spt_trace_any_dev(self.h, rays_ptr, n, hit_ptr))

    def shade_samples(self, samples37, rng, max_depth=None, camera=None, spp=None, integrator=None):
        s = np.ascontiguousarray(samples37, np.float32)
        g = np.ascontiguousarray(rng, np.float32) if rng is not None else None
        out = np.empty((len(s), D.NBANDS), np.float32)
        md = self.lowered.params.max_depth if max_depth is None else max_depth
        cam = camera if camera is not None else self.lowered.camera
        nspp = self.lowered.params.spp if spp is None else spp
        integ = self.lowered.params.integrator if integrator is None else integrator
        _check(lib().spt_shade_samples(self.h, C.byref(cam), integ, md, nspp, _p(s), _p(g), g.shape[1] if g is not None else 0,
                                       len(s), _p(out)))
        return out

    def render(self, film, params=None, camera=None):
        rp = params if params is not None else self.lowered.params
        cam = camera if camera is not None else self.lowered.camera
        _check(lib().spt_render(self.h, C.byref(cam), film.h, C.byref(rp)))

    def render_begin(self, film, params=None, camera=None):
        """Enqueue a frame and return (spt_render_begin); up to four frames may be in flight."""
        rp = params if params is not None else self.lowered.params
        cam = camera if camera is not None else self.lowered.camera
        _check(lib().spt_render_begin(self.h, C.byref(cam), film.h, C.byref(rp)))

    def render_end(self):
        """Wait for the oldest frame in flight (spt_render_end)."""
        _check(lib().spt_render_end(self.h))


class Film:
    """SpectralImageFilm accumulator on the device (spt_film_create[_external])."""

    def __init__(self, desc, device_ptr=None, ipc_handle=None, borrowed=None):
        """device_ptr: accumulate into caller-provided device memory; ipc_handle: the 64 bytes another process's
        Film.ipc_export() returned - this film then adds into THAT film's pixels over NVLink (spt_film_open_ipc);
        borrowed: an SptFilm* owned by someone else (MultiRenderer.film)."""
        self.desc = desc
        self.owner = borrowed is None
        if borrowed is not None:
            self.h = borrowed
        elif ipc_handle is not None:
            hb = (C.c_uint8 * IPC_HANDLE_BYTES).from_buffer_copy(bytes(ipc_handle))
            self.h = lib().spt_film_open_ipc(C.byref(desc), hb)
        elif device_ptr is None:
            self.h = lib().spt_film_create(C.byref(desc))
        else:
            self.h = lib().spt_film_create_external(C.byref(desc), C.c_void_p(device_ptr))
        if not self.h:
            raise SptError("spt_film_create: " + (lib().spt_last_error() or b"").decode())

    def ipc_export(self):
        hb = (C.c_uint8 * IPC_HANDLE_BYTES)()
        _check(lib().spt_film_ipc_export(self.h, hb))
        return bytes(hb)

    def close(self):
        if self.h:
            if self.owner:
                lib().spt_film_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def shape(self):
        return (self.desc.y_pixel_count, self.desc.x_pixel_count)

    def clear(self):
        _check(lib().spt_film_clear(self.h))

    def clear_idle(self):
        """Zero the film without waiting for frames running into other films (spt_film_clear_idle)."""
        _check(lib().spt_film_clear_idle(self.h))

    def add_samples(self, tables, xy, L):
        xy = np.ascontiguousarray(xy, np.float32); L = np.ascontiguousarray(L, np.float32)
        _check(lib().spt_film_add_samples(self.h, C.byref(tables), _p(xy), _p(L), len(xy)))

    def download(self, into=None):
        h, w = self.shape
        if into is None:
            c = np.empty((h, w, D.NBANDS), np.float32); wt = np.empty((h, w), np.float32)
        else:
            c, wt = into
        _check(lib().spt_film_download(self.h, _p(c), _p(wt)))
        return c, wt

    def device_ptr(self):
        return lib().spt_film_device_ptr(self.h)

    def write_dat(self, path):
        _check(lib().spt_film_write_dat(self.h, path.encode()))


def _stats_dict(st):
    d = {k: getattr(st, k) for k, _ in D.SptStats._fields_ if k not in ("class_ms", "class_launches", "class_rays", "pad_")}
    d["class_ms"] = list(st.class_ms); d["class_launches"] = list(st.class_launches); d["class_rays"] = list(st.class_rays)
    return d


class MultiRenderer:
    """The whole job on several GPUs of this process (spt_multi_*): scene replicated, image tile sets dealt round-robin, every
    GPU's film kernel adding straight into ONE film on the first device over NVLink."""

    def __init__(self, lowered, n_devices=0, devices=None):
        self.lowered = lowered
        arr = (C.c_int * len(devices))(*devices) if devices else None
        self.h = lib().spt_multi_create(C.byref(lowered.desc), C.byref(lowered.film), len(devices) if devices else int(n_devices), arr)
        if not self.h:
            raise SptError("spt_multi_create: " + (lib().spt_last_error() or b"").decode())
        self.film = Film(lowered.film, borrowed=lib().spt_multi_film(self.h))

    @property
    def device_count(self):
        return lib().spt_multi_device_count(self.h)

    def render(self, params=None, camera=None):
        rp = params if params is not None else self.lowered.params
        cam = camera if camera is not None else self.lowered.camera
        _check(lib().spt_multi_render(self.h, C.byref(cam), C.byref(rp)))

    def stats(self, index=0):
        st = D.SptStats()
        _check(lib().spt_multi_get_stats(self.h, index, C.byref(st)))
        return _stats_dict(st)

    def render_ms(self):
        return lib().spt_multi_last_render_ms(self.h)

    def close(self):
        if self.h:
            self.film.close()
            lib().spt_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def read_dat(path):
    """Reads a SpectralImageFilm .dat (src/film/spectralImage.cpp:319-369) -> [y][x][band] float64.
    Line 2 (lens info) is uninitialised in the reference for non-lens cameras and is skipped."""
    with open(path, "rb") as f:
        w, h, nb = (int(v) for v in f.readline().split())
        f.readline()
        data = np.frombuffer(f.read(), dtype=np.float64, count=w * h * nb)
    return data.reshape(nb, w, h).transpose(2, 1, 0).copy()
