/* spt.h — C ABI of the B200 spectral path-tracing core (libspt.so).
 *
 * This is the drop-in boundary for ONE hot path of scienstanford/pbrt-v2-spectral:
 *   Renderer::Render (src/core/renderer.h:35-46) as implemented by
 *   SamplerRenderer::Render / SamplerRendererTask::Run (src/renderers/samplerrenderer.cpp:188-222, :60-164),
 *   i.e. LDSampler -> PerspectiveCamera::GenerateRayDifferential -> BVHAccel::Intersect/IntersectP ->
 *   PathIntegrator::Li (+UniformSampleOneLight/EstimateDirect, BSDF::Sample_f/f/Pdf) ->
 *   SpectralImageFilm::AddSample.
 * The reference's parser, pbrtApi, Shape/Material/Light/Camera classes and BVH *build* stay on
 * the CPU; the host side (pbrt_v2_spectral_b200/host/lowering.cpp, compiled against the
 * reference's own headers) lowers the built Scene into the flat buffers described here and calls
 * this ABI. The same ABI is bound from Python with ctypes (pbrt_v2_spectral_b200/capi.py).
 *
 * Conventions: plain C, caller-owned HOST buffers unless a name ends in _dev, opaque handles,
 * int status returns (0 = ok, negative = error; text via spt_last_error()), thread-compatible
 * (not thread-safe) per handle. There is NO CPU fallback: every entry point that computes needs a
 * CUDA device and fails with SPT_ERR_CUDA otherwise.
 *
 * Spectra: SPT_NBANDS samples per spectrum, the reference's SampledSpectrum
 * (src/core/spectrum.h:41-43, 269-450: 32 bands 395-715 nm as shipped; 30 is a build option), in rows of SPT_BAND_PITCH floats.
 */
#ifndef SPT_H
#define SPT_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#ifndef SPT_NBANDS
#define SPT_NBANDS 32
#endif
/* Every spectrum crossing this boundary is stored as a ROW of SPT_BAND_PITCH floats - one 128-byte line, which the device reads
 * and writes with one coalesced access: entries 0 .. SPT_NBANDS-1 are the reference's SampledSpectrum samples, the rest is
 * padding and MUST BE ZERO. The library is built once per band count (libspt.so: 32 as the reference ships; libspt30.so:
 * the 30-band variant BASELINE.json names, nSpectralSamples = 30 in src/core/spectrum.h:41-43); SptSceneDesc::nbands must match. */
#define SPT_BAND_PITCH 32
#if SPT_NBANDS > SPT_BAND_PITCH || SPT_NBANDS < 1
#error "SPT_NBANDS must be 1..32"
#endif

#define SPT_OK            0
#define SPT_ERR_ARG      -1
#define SPT_ERR_CUDA     -2
#define SPT_ERR_UNSUPP   -3
#define SPT_ERR_IO       -4

/* ---- primitive table: one row per slot of BVHAccel::primitives (src/accelerators/bvh.h:58-63),
 *      i.e. in BVH leaf order (src/accelerators/bvh.cpp:177-182). ---- */
enum { SPT_PRIM_TRIANGLE = 0, SPT_PRIM_SPHERE = 1, SPT_PRIM_DISK = 2 };
/* prim_flags bits */
enum {
    SPT_PF_FLIP_NORMAL = 1,   /* Shape::ReverseOrientation ^ TransformSwapsHandedness (src/core/diffgeom.cpp:44-46) */
    SPT_PF_HAS_N       = 2,   /* TriangleMesh::n != NULL  (src/shapes/trianglemesh.cpp:285-289) */
    SPT_PF_HAS_UV      = 4,   /* TriangleMesh::uvs != NULL (src/shapes/trianglemesh.h:78-92) */
    SPT_PF_REVERSE     = 8,   /* Shape::ReverseOrientation alone (src/shapes/sphere.cpp:231,250) */
    SPT_PF_HAS_S       = 16   /* TriangleMesh::s != NULL */
};

/* Quadric record (sphere: src/shapes/sphere.cpp:31-41; disk: src/shapes/disk.cpp:31-39). */
typedef struct SptQuadric {
    int32_t kind;        /* SPT_PRIM_SPHERE / SPT_PRIM_DISK */
    int32_t xform;       /* row of xforms[] holding ObjectToWorld and its inverse */
    float radius;        /* sphere radius | disk radius */
    float zmin, zmax;    /* sphere clipping | disk: zmin = height, zmax = innerRadius */
    float thetaMin, thetaMax;
    float phiMax;
} SptQuadric;

/* ObjectToWorld and WorldToObject as the reference holds them: Transform::m and ::mInv, row-major
 * (src/core/transform.h:184-241). */
typedef struct SptXform { float m[16]; float minv[16]; } SptXform;

/* Material table row = the BSDF a reference material builds when all its textures are constant
 * (de-duplicated by value).
 *   MATTE   spec0 = Kd.Clamp(), p0 = sigma clamped to [0,90]        (src/materials/matte.cpp:34-60)
 *   PLASTIC spec0 = Kd.Clamp(), spec1 = Ks.Clamp(), p0 = roughness  (src/materials/plastic.cpp:34-61)
 *   METAL   spec0 = eta, spec1 = k, p0 = roughness                  (src/materials/metal.cpp:44-68)
 *   MIRROR  spec0 = Kr.Clamp(): SpecularReflection(Kr, FresnelNoOp) (src/materials/mirror.cpp:34-55)
 *   GLASS   spec0 = Kr.Clamp(), spec1 = Kt.Clamp(), p0 = index: SpecularReflection(Kr, FresnelDielectric(1, index))
 *           + SpecularTransmission(Kt, 1, index), each only if its spectrum is not black (src/materials/glass.cpp:34-58).
 *           `subsurface` lowers to GLASS with Kt = 0, index = eta: its BSDF under the path integrator
 *           (src/materials/subsurface.cpp:40-58)
 *   SUBSTRATE spec0 = Kd.Clamp(), spec1 = Ks.Clamp(), p0 = uroughness, p1 = vroughness: FresnelBlend(Kd, Ks,
 *           Anisotropic(1/u, 1/v))                                   (src/materials/substrate.cpp:34-56)
 * Textured parameters: tex_kd >= 0 replaces spec0 of MATTE / PLASTIC / SUBSTRATE by the image
 * texture textures[tex_kd] evaluated at the hit (Kd->Evaluate(dgs).Clamp()); tex_bump >= 0 is the displacement
 * texture handed to Material::Bump (src/core/material.cpp:39-82). -1: the constant in the row / the constant-0
 * displacement every reference material carries.
 */
enum { SPT_MAT_MATTE = 0, SPT_MAT_PLASTIC = 1, SPT_MAT_METAL = 2, SPT_MAT_MIRROR = 3, SPT_MAT_GLASS = 4,
       SPT_MAT_SUBSTRATE = 5,
       SPT_MAT_MEASURED = 6    /* measured BRDF brdfs[brdf]: IrregIsotropicBRDF (theta-phi .brdf data, src/materials/measured.cpp:77-120,
                                  185-205, src/core/reflection.cpp:239-263) or RegularHalfangleBRDF (MERL .binary data,
                                  measured.cpp:122-170, reflection.cpp:267-300); spectra unused */
};
typedef struct SptMaterial {
    int32_t type;
    float p0, p1, p2;
    float spec0[SPT_BAND_PITCH];
    float spec1[SPT_BAND_PITCH];
    int32_t tex_kd, tex_bump;
    int32_t brdf;                   /* MEASURED: row of brdfs[]; -1 otherwise */
    int32_t pad_;
} SptMaterial;

/* Measured isotropic BRDF (theta-phi .brdf file): the kd-tree the reference built over the samples' BRDFRemap
 * coordinates (src/core/kdtree.h:91-140), node for node, so that a radius search visits and sums the samples in the
 * reference's order. Node k's sample spectrum is the row brdf_spectra[(node_first + k) * SPT_BAND_PITCH ...]. */
typedef struct SptKdNode {
    float split_pos;
    uint32_t bits;                  /* splitAxis (3 = leaf) | hasLeftChild << 2 | rightChild << 3; the left child is node k+1 */
    float p[3];                     /* IrregIsotropicBRDFSample::p */
    float pad_[3];
} SptKdNode;
/* One measured BRDF. Theta-phi data (IrregIsotropicBRDF): kd-tree nodes [node_first, node_first + n_nodes), n_nodes > 0.
 * Half-angle data (MERL .binary -> RegularHalfangleBRDF, src/core/reflection.cpp:267-300, src/materials/measured.cpp:122-170):
 * n_nodes == 0 and the table of n_theta_h x n_theta_d x n_phi_d RGB triples the reference loaded (already scaled and clamped
 * at zero, :147-161) starts at merl_rgb[rgb_offset]; entry index = phiD + n_phi_d * (thetaD + thetaH * n_theta_d). */
typedef struct SptBrdfTable {
    uint32_t node_first, n_nodes;
    uint32_t n_theta_h, n_theta_d, n_phi_d;
    uint32_t pad_;
    uint64_t rgb_offset;
} SptBrdfTable;

/* Image texture = ImageTexture<RGBSpectrum,Spectrum> / ImageTexture<float,float> over a UVMapping2D
 * (src/textures/imagemap.h:60-100, src/core/texture.cpp:80-90) with the MIPMap pyramid the reference built
 * (src/core/mipmap.h:120-215: power-of-two resample, box-filtered levels), optionally under ScaleTextures whose
 * other operand is a constant float (src/textures/scale.h:40-58; float textures only).
 * Texels: level 0 starts at tex_texels[texel_offset], row-major [t][s], `channels` floats per texel (RGB as
 * RGBSpectrum::ToRGB gives it, or the float); level l+1 (max(1,w/2) x max(1,h/2)) follows level l. */
enum { SPT_WRAP_REPEAT = 0, SPT_WRAP_BLACK = 1, SPT_WRAP_CLAMP = 2 };    /* ImageWrap, src/core/mipmap.h:37-41 */
typedef struct SptTexture {
    int32_t channels;               /* 3: spectrum image map, 1: float image map */
    int32_t width, height;          /* level 0 */
    int32_t n_levels;
    int32_t wrap;                   /* SPT_WRAP_* */
    int32_t trilinear;              /* MIPMap::doTrilinear */
    int32_t no_filter;              /* MIPMap::noFiltering (fork addition, mipmap.h:217-226) */
    float   max_aniso;
    float   su, sv, du, dv;         /* UVMapping2D */
    float   scale;                  /* product of the constant operands of enclosing ScaleTextures (1: none) */
    int32_t pad_;
    uint64_t texel_offset;
} SptTexture;

/* Light table row.
 *   AREA     spectrum = Lemit; shapes [shape_first, shape_first+shape_count) of light_shapes[]
 *            (src/lights/diffuse.cpp:61-78, src/core/light.cpp:106-172)
 *   POINT    spectrum = Intensity; pos = LightToWorld(0,0,0)         (src/lights/point.cpp:34-49)
 *   INFINITE spectrum unused (already folded into the texels); env map + Distribution2D tables
 *            in the scene desc; xform = LightToWorld                 (src/lights/infinite.cpp:60-226)
 */
enum { SPT_LIGHT_AREA = 0, SPT_LIGHT_POINT = 1, SPT_LIGHT_INFINITE = 2 };
typedef struct SptLight {
    int32_t type;
    int32_t shape_first, shape_count;
    int32_t xform;
    float pos[3];
    float sum_area;                 /* ShapeSet::sumArea as the reference accumulated it */
    float spectrum[SPT_BAND_PITCH];
    int32_t n_samples;              /* Sampler::RoundSize(Light::nSamples): samples per camera hit under the
                                       directlighting integrator (src/integrators/directlighting.cpp:53-58); >= 1 */
    int32_t pad_[3];
} SptLight;

/* One entry of a ShapeSet (src/core/light.cpp:106-128): a refined, intersectable shape. */
typedef struct SptLightShape {
    int32_t kind;        /* SPT_PRIM_* */
    int32_t flags;       /* SPT_PF_* */
    int32_t data;        /* triangle: index into tri_vidx/3 ; quadric: row of quadrics[] */
    float area;          /* Shape::Area() as the reference computed it */
} SptLightShape;

/* Tables of SampledSpectrum statics the path needs (src/core/spectrum.h:297-351,417-422;
 * src/core/spectrum.cpp:136-176). Order of rgb_illum: White,Cyan,Magenta,Yellow,Red,Green,Blue. */
typedef struct SptSpectralTables {
    float cie_y[SPT_BAND_PITCH];
    float yint;
    float rgb_illum[7][SPT_BAND_PITCH];
    float rgb_refl[7][SPT_BAND_PITCH];  /* rgbRefl2Spect*, same order: FromRGB(rgb, SPECTRUM_REFLECTANCE) of image textures */
} SptSpectralTables;

typedef struct SptSceneDesc {
    int32_t nbands;                 /* must equal SPT_NBANDS of the library */
    /* BVHAccel::nodes, reference layout (src/accelerators/bvh.cpp:105-115), n_nodes * 32 bytes */
    uint32_t n_nodes;
    const void *bvh_nodes;
    /* per BVH slot */
    uint32_t n_prims;
    const uint8_t  *prim_kind;      /* SPT_PRIM_* */
    const uint8_t  *prim_flags;     /* SPT_PF_* */
    const uint32_t *prim_id;        /* Primitive::primitiveId (src/core/primitive.cpp:32,155-169) */
    const uint32_t *prim_data;      /* triangle: triangle number ; quadric: row of quadrics[] */
    const int32_t  *prim_material;  /* row of materials[] */
    const int32_t  *prim_light;     /* row of lights[] or -1 (GeometricPrimitive::areaLight) */
    const int32_t  *prim_xform;     /* row of xforms[] (Shape::ObjectToWorld) */
    /* triangle geometry: world-space P (src/shapes/trianglemesh.cpp:61-63), object-space N,S */
    uint32_t n_tris;
    const int32_t *tri_vidx;        /* 3 per triangle, into the vertex arrays below */
    uint32_t n_verts;
    const float *P;                 /* 3 per vertex */
    const float *N;                 /* 3 per vertex (zeros where the mesh has none) */
    const float *UV;                /* 2 per vertex (unused where the mesh has none) */
    uint32_t n_quadrics;   const SptQuadric *quadrics;
    uint32_t n_xforms;     const SptXform *xforms;
    uint32_t n_materials;  const SptMaterial *materials;
    uint32_t n_lights;     const SptLight *lights;
    uint32_t n_light_shapes; const SptLightShape *light_shapes;
    SptSpectralTables tables;
    /* infinite light (at most one): MIPMap level 0 texels RGB (src/core/mipmap.h), and the
     * Distribution2D built by the reference (src/lights/infinite.cpp:80-96): func rows [h][w],
     * cdf rows [h][w+1], per-row integrals [h], marginal func [h], marginal cdf [h+1]. */
    int32_t env_w, env_h;
    const float *env_rgb;
    const float *env_func, *env_cdf, *env_func_int;
    const float *env_marg_func, *env_marg_cdf;
    float env_marg_int;
    /* image textures */
    uint32_t n_textures;   const SptTexture *textures;
    uint64_t n_texels;     const float *tex_texels;
    const float *ewa_weight_lut;    /* MIPMap::weightLut[128] as the reference computed it (mipmap.h:205-213), or NULL */
    /* measured BRDFs */
    uint32_t n_brdfs;      const SptBrdfTable *brdfs;
    uint32_t n_brdf_nodes; const SptKdNode *brdf_nodes;
    const float *brdf_spectra;      /* n_brdf_nodes rows of SPT_BAND_PITCH */
    uint64_t n_merl_floats; const float *merl_rgb;      /* half-angle tables, 3 floats per entry (SptBrdfTable::rgb_offset) */
} SptSceneDesc;

/* PerspectiveCamera (src/cameras/perspective.cpp:33-106, src/core/camera.cpp:84-103). */
typedef struct SptCameraDesc {
    float raster_to_camera[16];     /* ProjectiveCamera::RasterToCamera.m, row-major */
    float camera_to_world[16];      /* CameraToWorld.startTransform->m (static scenes) */
    float lens_radius, focal_distance;
    float shutter_open, shutter_close;
    float dx_camera[3], dy_camera[3];   /* PerspectiveCamera::dxCamera/dyCamera (perspective.cpp:45-46): the offset rays
                                           of GenerateRayDifferential (:73-106) that image textures are filtered with */
} SptCameraDesc;

/* SpectralImageFilm (src/film/spectralImage.cpp:40-75,176-194). */
typedef struct SptFilmDesc {
    int32_t x_resolution, y_resolution;
    int32_t x_pixel_start, y_pixel_start, x_pixel_count, y_pixel_count;
    float filter_xwidth, filter_ywidth, filter_inv_xwidth, filter_inv_ywidth;
    float filter_table[256];        /* 16x16, row = y (src/film/spectralImage.cpp:55-66) */
} SptFilmDesc;

/* Surface integrators.
 *   PATH        PathIntegrator (src/integrators/path.cpp:44-115)
 *   DIRECT_ALL  DirectLightingIntegrator with strategy "all" (src/integrators/directlighting.cpp:70-105):
 *               emitted light + UniformSampleAllLights (src/core/integrator.cpp:39-71) at the camera ray's hit.
 *               Scenes with specular materials are not lowered under it (its SpecularReflect/Transmit recursion). */
enum { SPT_INTEGRATOR_PATH = 0, SPT_INTEGRATOR_DIRECT_ALL = 1,
       SPT_INTEGRATOR_DIRECT_ONE = 2   /* DirectLightingIntegrator, strategy "one": emitted light + UniformSampleOneLight
                                          (src/core/integrator.cpp:72-107) at the camera ray's hit */
};

/* What SamplerRenderer + LDSampler + the surface integrator are configured with. */
typedef struct SptRenderParams {
    int32_t spp;                    /* LDSampler::nPixelSamples (power of two) */
    int32_t max_depth;              /* PathIntegrator::maxDepth (src/integrators/path.cpp:118-121) */
    int32_t x_start, x_end, y_start, y_end;   /* sample extent (Film::GetSampleExtent) */
    uint64_t seed;
    int32_t tile_rank, tile_nranks; /* image tile set of this GPU: tiles t with t % nranks == rank */
    int32_t tile_size;              /* tile edge in pixels (0 = default 32) */
    int32_t wave_pixels;            /* pixels per wavefront (0 = default) */
    int32_t skip_border;            /* 1: skip the sample-extent border column/row whose samples the
                                       box filter rejects */
    int32_t integrator;             /* SPT_INTEGRATOR_* */
} SptRenderParams;

/* kernel classes of the wavefront, for per-class device time (CUDA events on the launching stream).
 * SPT_K_TRACE_PATH times the ONE traversal launch of a bounce, which carries the bounce's path rays together with the
 * previous bounce's shadow and MIS rays (class_ms of SPT_K_TRACE_SHADOW / _MIS stay 0; their class_rays are counted);
 * SPT_K_ACCUMULATE = k_addlight (L += T (Le + Ld)), SPT_K_ADVANCE = k_advance (throughput, Russian roulette, next ray). */
enum { SPT_K_GEN = 0, SPT_K_TRACE_PATH = 1, SPT_K_SHADE = 2, SPT_K_TRACE_SHADOW = 3, SPT_K_TRACE_MIS = 4,
       SPT_K_ACCUMULATE = 5, SPT_K_FILM = 6, SPT_K_ADVANCE = 7, SPT_K_CLASSES = 8 };

typedef struct SptStats {
    uint64_t camera_samples;        /* samples traced (cumulative) */
    uint64_t closest_rays, any_rays;    /* rays traced (cumulative) */
    /* device counters, only while spt_scene_enable_counters(on): BVH nodes visited / primitives tested */
    uint64_t node_visits_closest, prim_tests_closest, node_visits_any, prim_tests_any;
    uint64_t kernel_launches;       /* cumulative */
    double   render_ms;             /* CUDA-event time of the last spt_render, first launch -> film final */
    double   trace_ms;              /* CUDA-event time of the last stand-alone spt_trace_* kernel */
    double   class_ms[SPT_K_CLASSES];        /* last spt_render: device time per kernel class */
    uint64_t class_launches[SPT_K_CLASSES];  /* last spt_render: launches per kernel class */
    uint64_t class_rays[SPT_K_CLASSES];      /* last spt_render: rays (or paths) processed per class */
    /* last spt_render: BSDF-sampled MIS rays of EstimateDirect (src/core/integrator.cpp:139-163) that were
     * NOT traced because they miss every shape of the sampled area light, so their contribution is exactly
     * zero; the reference traces them (they are counted in its rays/sample figure) */
    uint64_t mis_rays_elided;
    uint64_t first_vertices;        /* last spt_render: camera rays that found a surface (path vertices of bounce 0) */
    int32_t  lanes_used;            /* last spt_render: streams the waves were dealt to; with more than one, kernels of the
                                       lanes overlap and class_ms sums per-lane event deltas (can exceed render_ms) */
    int32_t  pad_;
    /* last spt_render: the FIRST launch of each class (the first wave's bounce 0: camera rays, first vertices) - one well-defined
     * launch per kernel for roofline figures: its device time and the units (rays / vertices / samples) it processed */
    double   first_launch_ms[SPT_K_CLASSES];
    uint64_t first_launch_units[SPT_K_CLASSES];
} SptStats;

typedef struct SptScene SptScene;
typedef struct SptFilm  SptFilm;

int         spt_nbands(void);
const char *spt_last_error(void);
int         spt_device_count(void);
int         spt_set_device(int ordinal);
/* Page-locked host memory for buffers that cross the boundary every frame (the film a host reads
 * back, scene tables it uploads): copies to and from it run at full PCIe rate. Plain malloc'ed
 * buffers are accepted everywhere as well. Returns NULL on failure. */
void       *spt_host_alloc(uint64_t bytes);
void        spt_host_free(void *p);
/* Releases the device memory the library caches between scenes (wave state of destroyed scenes). */
void        spt_trim(void);

/* Scene: uploads every table to HBM (replaces nothing in the reference; it is the hand-off). */
SptScene *spt_scene_create(const SptSceneDesc *desc);
void      spt_scene_destroy(SptScene *scene);
/* Waves of a frame are dealt to several streams ("lanes", at most `lanes` of them: 1..4, default 4; a frame that fits two waves uses two) so that the drain of one wave's
 * persistent trace kernel overlaps the other waves' work; lanes = 1 keeps everything on one stream (exact
 * per-kernel times). */
int       spt_scene_set_lanes(SptScene *scene, int lanes);
/* Traversal layout. EXACT (default): the reference's node order, slab and leaf arithmetic - first-hit ids
 * and distances bit-identical to BVHAccel::Intersect/IntersectP. FAST: a 4-wide BVH collapsed from the same flattened tree
 * (src/accelerators/bvh.cpp:354-372), children entered nearest first: the same primitive and the bit-identical distance
 * wherever the closest hit is unique; ties between primitives at exactly equal distance and rays grazing a box edge may
 * resolve differently. Built on the first switch to FAST (host side, from the nodes already in HBM). */
enum { SPT_TRAVERSAL_EXACT = 0, SPT_TRAVERSAL_FAST = 1 };
int       spt_scene_set_traversal(SptScene *scene, int mode);
int       spt_scene_enable_counters(SptScene *scene, int on);
int       spt_get_stats(SptScene *scene, SptStats *out);
/* SptStats::render_ms of the last spt_render alone (spt_get_stats also folds ~80 per-launch event deltas into class_ms). */
double    spt_last_render_ms(SptScene *scene);

/* K1: PerspectiveCamera::GenerateRayDifferential (src/cameras/perspective.cpp:73-106).
 * samples: n x 5 floats {imageX, imageY, lensU, lensV, time}; out_rays: n x 8 {o, d, mint, maxt}. */
int spt_camera_rays(const SptCameraDesc *cam, const float *samples, uint64_t n, float *out_rays);

/* K2: BVHAccel::Intersect (src/accelerators/bvh.cpp:380-432). rays: n x 8 {o,d,mint,maxt}.
 * out_slot: BVH slot or 0xffffffff on miss; out_prim_id: Primitive::primitiveId or 0; out_t: ray.maxt
 * after the call. Any output pointer may be NULL. */
int spt_trace_closest(SptScene *scene, const float *rays, uint64_t n,
                      uint32_t *out_slot, uint32_t *out_prim_id, float *out_t);
/* K3: BVHAccel::IntersectP (src/accelerators/bvh.cpp:435-481). out_hit: 0/1 per ray. */
int spt_trace_any(SptScene *scene, const float *rays, uint64_t n, uint8_t *out_hit);
/* Same kernels on buffers already resident in HBM (device pointers), for the timed kernel-only runs. */
int spt_trace_closest_dev(SptScene *scene, const float *rays_dev, uint64_t n,
                          uint32_t *out_slot_dev, float *out_t_dev);
int spt_trace_any_dev(SptScene *scene, const float *rays_dev, uint64_t n, uint8_t *out_hit_dev);

/* SamplerRenderer::Li + the surface integrator's Li for caller-supplied sample vectors
 * (src/renderers/samplerrenderer.cpp:225-247, src/integrators/path.cpp:44-115, directlighting.cpp:70-105).
 * spp: Sampler::samplesPerPixel - the camera ray differentials are scaled by 1/sqrt(spp)
 * (src/renderers/samplerrenderer.cpp:91, src/core/geometry.h:368-373) before image textures are filtered with them.
 * samples: n x 37 floats in the reference's Sample memory order {imageX,imageY,lensU,lensV,time,
 * oneD[0..13], twoD[0..8][2]} (src/core/sampler.cpp:88-117, src/integrators/path.cpp:33-41);
 * rng: n x n_rng floats consumed in order where the reference draws from RNG (bounces >= 3 and
 * Russian roulette, src/integrators/path.cpp:82,97; src/core/integrator.cpp:84-99).
 * integrator = SPT_INTEGRATOR_DIRECT_ALL: samples are n x (7 + 6 N) floats, N = sum of the lights' n_samples, in the
 * order DirectLightingIntegrator::RequestSamples leaves them (directlighting.cpp:46-60, src/core/light.cpp:56-60,
 * src/core/reflection.cpp:494-498): {5 camera floats, per light [light component x n_i, bsdf component x n_i],
 * 2 floats of the volume integrator, per light [light position x 2 n_i, bsdf direction x 2 n_i]}; rng is unused.
 * integrator = SPT_INTEGRATOR_DIRECT_ONE: n x 14 floats {5 camera floats, light component, light number, bsdf component,
 * 2 floats of the volume integrator, light position x 2, bsdf direction x 2} (directlighting.cpp:61-68).
 * out_L: n x SPT_NBANDS radiance BEFORE the NaN/negative/inf guards of
 * src/renderers/samplerrenderer.cpp:119-133. */
int spt_shade_samples(SptScene *scene, const SptCameraDesc *cam, int32_t integrator, int32_t max_depth, int32_t spp,
                      const float *samples, const float *rng, int32_t n_rng, uint64_t n, float *out_L);

/* Film: SpectralImageFilm pixels {c[SPT_NBANDS], weightSum} (src/film/spectralImage.h:74-84). */
SptFilm *spt_film_create(const SptFilmDesc *desc);
/* Same, accumulating into caller-provided device memory of
 * y_pixel_count * x_pixel_count * (SPT_NBANDS+1) floats (e.g. an NCCL buffer). */
SptFilm *spt_film_create_external(const SptFilmDesc *desc, float *pixels_dev);
void     spt_film_destroy(SptFilm *film);
int      spt_film_clear(SptFilm *film);
/* The same zeroing without waiting for the rest of the device: the CALLER guarantees that no render into this film is in flight
 * (its frame's end-of-frame barrier has completed); frames running into OTHER films go on. Returns when the film is zero. */
int      spt_film_clear_idle(SptFilm *film);
/* K7 on caller-supplied samples: SpectralImageFilm::AddSample (src/film/spectralImage.cpp:77-152)
 * preceded by the radiance guards of src/renderers/samplerrenderer.cpp:119-133.
 * image_xy: n x 2 {imageX, imageY}; L: n x SPT_NBANDS. */
int      spt_film_add_samples(SptFilm *film, const SptSpectralTables *tables,
                              const float *image_xy, const float *L, uint64_t n);
/* c: [y][x][SPT_NBANDS] un-normalised sums, weight: [y][x]. Either may be NULL. */
int      spt_film_download(SptFilm *film, float *c, float *weight);
float   *spt_film_device_ptr(SptFilm *film);   /* [y][x][SPT_NBANDS+1], band SPT_NBANDS = weightSum */
/* SpectralImageFilm::WriteImage's .dat (src/film/spectralImage.cpp:267-378): two text lines then
 * [band][x][y] float64. */
int      spt_film_write_dat(SptFilm *film, const char *path);

/* The whole job: SamplerRenderer::Render without the final WriteImage
 * (src/renderers/samplerrenderer.cpp:188-222). Accumulates into film. */
int spt_render(SptScene *scene, const SptCameraDesc *cam, SptFilm *film, const SptRenderParams *params);
/* The same job in two calls, for a host that renders frame after frame (the reference renders one image per process, so it has
 * no counterpart there): spt_render_begin enqueues the frame and returns; spt_render_end waits for the OLDEST frame begun and
 * finishes its statistics. Up to four frames may be in flight on a scene - begin(k + 1) before end(k) - so that the device goes
 * from one frame to the next without waiting for the host; each stream runs its waves in order, in its own wave buffers
 * (frames into different films, or into one film that is meant to accumulate both).
 * A frame of up to 2^23 paths begun this way runs as ONE wave (spt_render cuts it into two half-size waves on two streams) and
 * consecutive frames go to different streams: whole frames, not half frames, are what overlaps on the device.
 * Per-kernel-class times (SptStats::class_ms) are only collected for a frame begun while no other was in flight. Between begin
 * and end only these two calls, spt_get_stats and spt_last_render_ms may be used on the scene. */
int spt_render_begin(SptScene *scene, const SptCameraDesc *cam, SptFilm *film, const SptRenderParams *params);
int spt_render_end(SptScene *scene);

/* ---- several GPUs: the scene is replicated, the image is cut into square tiles dealt round-robin to the GPUs
 * (SptRenderParams::tile_rank / tile_nranks, all samples of a pixel on one GPU), and K7 of every GPU adds its samples STRAIGHT
 * INTO ONE film that lives on the first GPU - peer stores / reductions over NVLink issued by the film kernel itself, no
 * separate exchange step, no staging copy. What is left of the "gather" is a barrier at the end of the frame. This replaces
 * the reference's own partition of the image into independent tasks (src/renderers/samplerrenderer.cpp:203-214).
 *
 * One process, N devices: spt_multi_* below (one host thread per device inside the library).
 * One process per GPU (torchrun): the process that owns the film exports it (spt_film_ipc_export), the others open it
 * (spt_film_open_ipc) and render their tile sets into it with spt_render; the host barriers (e.g. over NCCL) at frame end. */
#define SPT_IPC_HANDLE_BYTES 64
int      spt_film_ipc_export(SptFilm *film, uint8_t handle[SPT_IPC_HANDLE_BYTES]);   /* library-allocated films only */
SptFilm *spt_film_open_ipc(const SptFilmDesc *desc, const uint8_t handle[SPT_IPC_HANDLE_BYTES]);

typedef struct SptMulti SptMulti;
/* devices: n_devices CUDA ordinals, or NULL for 0 .. n_devices-1; n_devices <= 0: every visible device. The film lives on
 * devices[0]; every other device must be able to reach it (cudaDeviceCanAccessPeer), else NULL / SPT_ERR_UNSUPP. */
SptMulti *spt_multi_create(const SptSceneDesc *scene, const SptFilmDesc *film, int n_devices, const int *devices);
void      spt_multi_destroy(SptMulti *m);
int       spt_multi_device_count(SptMulti *m);
/* The whole job on all devices: tile_rank / tile_nranks of params are set per device (a tile_size of 0 means 32).
 * Accumulates into the shared film; returns when every device has finished. */
int       spt_multi_render(SptMulti *m, const SptCameraDesc *cam, const SptRenderParams *params);
SptFilm  *spt_multi_film(SptMulti *m);                        /* the complete film (download / clear / write_dat as usual) */
int       spt_multi_get_stats(SptMulti *m, int index, SptStats *out);      /* per device, index 0 .. count-1 */
double    spt_multi_last_render_ms(SptMulti *m);              /* slowest device's SptStats::render_ms of the last frame */

#ifdef __cplusplus
}
#endif
#endif /* SPT_H */
