"""ctypes mirrors of the plain-C structs in include/spt.h (kept field-for-field in the same order)."""
import ctypes as C
import os

# valid bands of the library in use (libspt.so: 32 as the reference ships; SPT_NBANDS=30 selects libspt30.so) and the row
# pitch of every spectrum crossing the boundary (include/spt.h: SPT_BAND_PITCH; entries beyond NBANDS are zero)
NBANDS = int(os.environ.get("SPT_NBANDS", "32"))
BAND_PITCH = 32


class SptSpectralTables(C.Structure):
    _fields_ = [("cie_y", C.c_float * BAND_PITCH), ("yint", C.c_float), ("rgb_illum", (C.c_float * BAND_PITCH) * 7),
                ("rgb_refl", (C.c_float * BAND_PITCH) * 7)]


class SptSceneDesc(C.Structure):
    _fields_ = [
        ("nbands", C.c_int32),
        ("n_nodes", C.c_uint32), ("bvh_nodes", C.c_void_p),
        ("n_prims", C.c_uint32),
        ("prim_kind", C.c_void_p), ("prim_flags", C.c_void_p), ("prim_id", C.c_void_p),
        ("prim_data", C.c_void_p), ("prim_material", C.c_void_p), ("prim_light", C.c_void_p),
        ("prim_xform", C.c_void_p),
        ("n_tris", C.c_uint32), ("tri_vidx", C.c_void_p),
        ("n_verts", C.c_uint32), ("P", C.c_void_p), ("N", C.c_void_p), ("UV", C.c_void_p),
        ("n_quadrics", C.c_uint32), ("quadrics", C.c_void_p),
        ("n_xforms", C.c_uint32), ("xforms", C.c_void_p),
        ("n_materials", C.c_uint32), ("materials", C.c_void_p),
        ("n_lights", C.c_uint32), ("lights", C.c_void_p),
        ("n_light_shapes", C.c_uint32), ("light_shapes", C.c_void_p),
        ("tables", SptSpectralTables),
        ("env_w", C.c_int32), ("env_h", C.c_int32),
        ("env_rgb", C.c_void_p), ("env_func", C.c_void_p), ("env_cdf", C.c_void_p),
        ("env_func_int", C.c_void_p), ("env_marg_func", C.c_void_p), ("env_marg_cdf", C.c_void_p),
        ("env_marg_int", C.c_float),
        ("n_textures", C.c_uint32), ("textures", C.c_void_p),
        ("n_texels", C.c_uint64), ("tex_texels", C.c_void_p),
        ("ewa_weight_lut", C.c_void_p),
        ("n_brdfs", C.c_uint32), ("brdfs", C.c_void_p),
        ("n_brdf_nodes", C.c_uint32), ("brdf_nodes", C.c_void_p), ("brdf_spectra", C.c_void_p),
        ("n_merl_floats", C.c_uint64), ("merl_rgb", C.c_void_p),
    ]


class SptCameraDesc(C.Structure):
    _fields_ = [("raster_to_camera", C.c_float * 16), ("camera_to_world", C.c_float * 16),
                ("lens_radius", C.c_float), ("focal_distance", C.c_float),
                ("shutter_open", C.c_float), ("shutter_close", C.c_float),
                ("dx_camera", C.c_float * 3), ("dy_camera", C.c_float * 3)]


class SptFilmDesc(C.Structure):
    _fields_ = [("x_resolution", C.c_int32), ("y_resolution", C.c_int32),
                ("x_pixel_start", C.c_int32), ("y_pixel_start", C.c_int32),
                ("x_pixel_count", C.c_int32), ("y_pixel_count", C.c_int32),
                ("filter_xwidth", C.c_float), ("filter_ywidth", C.c_float),
                ("filter_inv_xwidth", C.c_float), ("filter_inv_ywidth", C.c_float),
                ("filter_table", C.c_float * 256)]


class SptRenderParams(C.Structure):
    _fields_ = [("spp", C.c_int32), ("max_depth", C.c_int32),
                ("x_start", C.c_int32), ("x_end", C.c_int32), ("y_start", C.c_int32), ("y_end", C.c_int32),
                ("seed", C.c_uint64),
                ("tile_rank", C.c_int32), ("tile_nranks", C.c_int32), ("tile_size", C.c_int32),
                ("wave_pixels", C.c_int32), ("skip_border", C.c_int32), ("integrator", C.c_int32)]


INTEGRATOR_PATH, INTEGRATOR_DIRECT_ALL, INTEGRATOR_DIRECT_ONE = 0, 1, 2


# K_TRACE_PATH times the one traversal launch of a bounce (its path rays + the previous bounce's shadow / MIS rays); the
# shadow / MIS classes only count rays
K_GEN, K_TRACE_PATH, K_SHADE, K_TRACE_SHADOW, K_TRACE_MIS, K_ACCUMULATE, K_FILM, K_ADVANCE, K_CLASSES = 0, 1, 2, 3, 4, 5, 6, 7, 8
K_KERNELS = ["k_gen_camera", "k_trace_multi", "k_compact_hits + k_shade", "k_trace_multi", "k_trace_multi", "k_addlight", "k_film_add", "k_advance"]
K_NAMES = ["gen_camera", "trace", "shade", "trace_any_shadow", "trace_closest_mis", "addlight", "film_add", "advance"]


class SptStats(C.Structure):
    _fields_ = [("camera_samples", C.c_uint64), ("closest_rays", C.c_uint64), ("any_rays", C.c_uint64),
                ("node_visits_closest", C.c_uint64), ("prim_tests_closest", C.c_uint64),
                ("node_visits_any", C.c_uint64), ("prim_tests_any", C.c_uint64),
                ("kernel_launches", C.c_uint64),
                ("render_ms", C.c_double), ("trace_ms", C.c_double),
                ("class_ms", C.c_double * K_CLASSES), ("class_launches", C.c_uint64 * K_CLASSES),
                ("class_rays", C.c_uint64 * K_CLASSES), ("mis_rays_elided", C.c_uint64), ("first_vertices", C.c_uint64),
                ("lanes_used", C.c_int32), ("pad_", C.c_int32),
                ("first_launch_ms", C.c_double * K_CLASSES), ("first_launch_units", C.c_uint64 * K_CLASSES)]


# row sizes of the table structs (bytes), for sanity checks against the container file
SIZEOF_QUADRIC = 32
SIZEOF_XFORM = 128
SIZEOF_MATERIAL = 16 + 2 * 4 * BAND_PITCH + 16
SIZEOF_TEXTURE = 64
SIZEOF_BRDF_TABLE = 32
SIZEOF_KD_NODE = 32
SIZEOF_LIGHT = 32 + 4 * BAND_PITCH + 16
SIZEOF_LIGHT_SHAPE = 16
