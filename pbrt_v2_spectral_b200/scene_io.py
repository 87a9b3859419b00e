"""Reader/writer for the lowered-scene container (pbrt_v2_spectral_b200/host/lowering.h:
SptContainerWriter) and the holder that turns it into a ctypes SptSceneDesc for the C ABI."""
import ctypes as C
import struct

import numpy as np

from . import ctypes_defs as D

_DTYPES = {0: np.uint8, 1: np.int32, 2: np.uint32, 3: np.float32, 4: np.uint64}
_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


def load_container(path):
    """name -> numpy array, for a file written by SptContainerWriter."""
    with open(path, "rb") as f:
        buf = f.read()
    if path.endswith(".xz"):
        import lzma
        buf = lzma.decompress(buf)
    if buf[:8] != b"SPTSCN01":
        raise ValueError("%s: not an SPTSCN01 container" % path)
    (count,) = struct.unpack_from("<I", buf, 8)
    pos = 12
    out = {}
    for _ in range(count):
        (nl,) = struct.unpack_from("<I", buf, pos); pos += 4
        name = buf[pos:pos + nl].decode(); pos += nl
        dtype, ndim = struct.unpack_from("<II", buf, pos); pos += 8
        dims = struct.unpack_from("<%dQ" % ndim, buf, pos); pos += 8 * ndim
        (nbytes,) = struct.unpack_from("<Q", buf, pos); pos += 8
        arr = np.frombuffer(buf, dtype=_DTYPES[dtype], count=nbytes // np.dtype(_DTYPES[dtype]).itemsize, offset=pos)
        pos += nbytes + (8 - (nbytes + nl) % 8) % 8
        if ndim == 2 and arr.size == dims[0] * dims[1]:
            arr = arr.reshape(dims)
        out[name] = arr.copy()
    return out


def save_container(path, arrays):
    """Inverse of load_container (used by the synthetic-scene generator)."""
    with open(path, "wb") as f:
        f.write(b"SPTSCN01")
        f.write(struct.pack("<I", len(arrays)))
        for name, arr in arrays.items():
            arr = np.ascontiguousarray(arr)
            nb = name.encode()
            f.write(struct.pack("<I", len(nb))); f.write(nb)
            dims = arr.shape if arr.ndim == 2 else (arr.size,)
            f.write(struct.pack("<II", _CODES[arr.dtype], len(dims)))
            f.write(struct.pack("<%dQ" % len(dims), *dims))
            raw = arr.tobytes()
            f.write(struct.pack("<Q", len(raw))); f.write(raw)
            f.write(b"\0" * ((8 - (len(raw) + len(nb)) % 8) % 8))


def _from_bytes(struct_type, arr):
    raw = np.ascontiguousarray(arr).tobytes()
    if len(raw) != C.sizeof(struct_type):
        raise ValueError("%s: %d bytes in file, %d expected" % (struct_type.__name__, len(raw), C.sizeof(struct_type)))
    return struct_type.from_buffer_copy(raw)


class LoweredScene:
    """A lowered scene held in host memory (numpy) + the ctypes descriptors that point into it."""

    def __init__(self, arrays):
        self.a = {k: np.ascontiguousarray(v) for k, v in arrays.items()}
        a = self.a
        for key, dt in (("textures", np.uint8), ("tex_texels", np.float32), ("ewa_weight_lut", np.float32),
                        ("brdfs", np.uint8), ("brdf_nodes", np.uint8), ("brdf_spectra", np.float32), ("merl_rgb", np.float32)):
            a.setdefault(key, np.zeros(0, dt))
        if int(a["nbands"][0]) != D.NBANDS:
            raise ValueError("scene has %d bands, library expects %d" % (int(a["nbands"][0]), D.NBANDS))
        for key, size in (("quadrics", D.SIZEOF_QUADRIC), ("xforms", D.SIZEOF_XFORM), ("materials", D.SIZEOF_MATERIAL),
                          ("lights", D.SIZEOF_LIGHT), ("light_shapes", D.SIZEOF_LIGHT_SHAPE), ("textures", D.SIZEOF_TEXTURE)):
            if a[key].size % size:
                raise ValueError("table %s: size %d is not a multiple of %d" % (key, a[key].size, size))
        self.camera = _from_bytes(D.SptCameraDesc, a["camera"])
        self.film = _from_bytes(D.SptFilmDesc, a["film"])
        self.params = _from_bytes(D.SptRenderParams, a["params"])
        self.tables = _from_bytes(D.SptSpectralTables, a["tables"])
        self.film_filename = a["film_filename"].tobytes().decode() if "film_filename" in a else ""
        self.desc = self._make_desc()

    @staticmethod
    def load_arrays(path):
        """name -> array of a scene file with its references resolved (recursively):
        "<array>@" = file name (relative to this file) of another container holding <array> - scenes that share a big table,
        the shipped-floor scenes' 22 MB texel pool, keep one copy (oracle/make_golden.py share_arrays);
        "base_scene" = a DELTA file {base_scene, camera, film, params, film_filename}: the full-size bench workloads share
        every scene table with their small golden variant and differ in resolution and sample count only."""
        import os
        here = os.path.dirname(os.path.abspath(path))
        a = load_container(path)
        for key in [k for k in a if k.endswith("@")]:
            other = os.path.normpath(os.path.join(here, a.pop(key).tobytes().decode()))
            a[key[:-1]] = LoweredScene.load_arrays(other)[key[:-1]]
        if "base_scene" in a:
            full = LoweredScene.load_arrays(os.path.normpath(os.path.join(here, a["base_scene"].tobytes().decode())))
            full.update({k: v for k, v in a.items() if k != "base_scene"})
            a = full
        return a

    @classmethod
    def load(cls, path):
        return cls(LoweredScene.load_arrays(path))

    def _ptr(self, key):
        arr = self.a[key]
        return arr.ctypes.data if arr.size else None

    def _make_desc(self):
        a = self.a
        d = D.SptSceneDesc()
        d.nbands = D.NBANDS
        d.n_nodes = a["bvh_nodes"].size // 32
        d.bvh_nodes = self._ptr("bvh_nodes")
        d.n_prims = a["prim_kind"].size
        for k in ("prim_kind", "prim_flags", "prim_id", "prim_data", "prim_material", "prim_light", "prim_xform",
                  "tri_vidx", "P", "N", "UV", "quadrics", "xforms", "materials", "lights", "light_shapes",
                  "env_rgb", "env_func", "env_cdf", "env_func_int", "env_marg_func", "env_marg_cdf",
                  "textures", "tex_texels", "ewa_weight_lut", "brdfs", "brdf_nodes", "brdf_spectra", "merl_rgb"):
            setattr(d, k, self._ptr(k))
        d.n_tris = a["tri_vidx"].size // 3
        d.n_verts = a["P"].size // 3
        d.n_quadrics = a["quadrics"].size // D.SIZEOF_QUADRIC
        d.n_xforms = a["xforms"].size // D.SIZEOF_XFORM
        d.n_materials = a["materials"].size // D.SIZEOF_MATERIAL
        d.n_lights = a["lights"].size // D.SIZEOF_LIGHT
        d.n_light_shapes = a["light_shapes"].size // D.SIZEOF_LIGHT_SHAPE
        d.tables = self.tables
        d.env_w, d.env_h = int(a["env_dims"][0]), int(a["env_dims"][1])
        d.env_marg_int = float(a["env_marg_int"][0])
        d.n_textures = a["textures"].size // D.SIZEOF_TEXTURE
        d.n_texels = a["tex_texels"].size
        d.n_brdfs = a["brdfs"].size // D.SIZEOF_BRDF_TABLE
        d.n_brdf_nodes = a["brdf_nodes"].size // D.SIZEOF_KD_NODE
        d.n_merl_floats = a["merl_rgb"].size
        return d

    @property
    def n_prims(self):
        return int(self.desc.n_prims)

    @property
    def n_nodes(self):
        return int(self.desc.n_nodes)
