// lowering.h — host side of the drop-in boundary: turns a built reference Scene into the flat
// buffers of include/spt.h. Compiled against the reference's own headers (it is the code a
// maintainer adds to the reference tree); nothing here runs on the GPU.
#ifndef SPT_HOST_LOWERING_H
#define SPT_HOST_LOWERING_H
#include <stdint.h>
#include <string>
#include <vector>
#include <fstream>
#include <string.h>
#include "spt.h"


// Container file shared by the lowered scene and the golden-vector dumps:
// "SPTSCN01", u32 count, then per array {u32 name_len, name, u32 dtype, u32 ndim, u64 dims[ndim],
// u64 nbytes, data, zero padding so that name+data end on an 8-byte boundary}.
// dtype: 0 u8, 1 i32, 2 u32, 3 f32, 4 u64.
struct SptContainerWriter {
    std::ofstream f;
    uint32_t count;
    std::streampos countPos;
    bool begin(const std::string &path) {
        f.open(path.c_str(), std::ios::binary);
        if (!f) return false;
        f.write("SPTSCN01", 8);
        countPos = f.tellp();
        count = 0;
        f.write((const char *)&count, 4);
        return true;
    }
    void put(const char *name, uint32_t dtype, const void *data, uint64_t nbytes,
             uint64_t d0, uint64_t d1 = 0) {
        uint32_t nl = (uint32_t)strlen(name);
        f.write((const char *)&nl, 4);
        f.write(name, nl);
        f.write((const char *)&dtype, 4);
        uint32_t ndim = d1 ? 2 : 1;
        f.write((const char *)&ndim, 4);
        f.write((const char *)&d0, 8);
        if (d1) f.write((const char *)&d1, 8);
        f.write((const char *)&nbytes, 8);
        if (nbytes) f.write((const char *)data, nbytes);
        static const char zeros[8] = {0};
        uint64_t padn = (8 - (nbytes + nl) % 8) % 8;
        if (padn) f.write(zeros, padn);
        ++count;
    }
    template <typename T> void vec(const char *name, uint32_t dtype, const std::vector<T> &v,
                                   uint64_t inner = 0) {
        uint64_t n = v.size();
        put(name, dtype, v.empty() ? NULL : &v[0], n * sizeof(T), inner ? n / inner : n, inner);
    }
    template <typename T> void pod(const char *name, const std::vector<T> &v) {
        put(name, 0, v.empty() ? NULL : &v[0], v.size() * sizeof(T), v.size(), sizeof(T));
    }
    void end() {
        f.seekp(countPos);
        f.write((const char *)&count, 4);
        f.close();
    }
};

class Scene;
class Camera;
class Sampler;
class SurfaceIntegrator;

// Owns every array a SptSceneDesc points into.
struct LoweredScene {
    std::vector<uint8_t> bvh_nodes;          // n_nodes * 32 bytes, reference LinearBVHNode layout
    std::vector<uint8_t> prim_kind, prim_flags;
    std::vector<uint32_t> prim_id, prim_data;
    std::vector<int32_t> prim_material, prim_light, prim_xform;
    std::vector<int32_t> tri_vidx;
    std::vector<float> P, N, UV;
    std::vector<SptQuadric> quadrics;
    std::vector<SptXform> xforms;
    std::vector<SptMaterial> materials;
    std::vector<SptLight> lights;
    std::vector<SptLightShape> light_shapes;
    SptSpectralTables tables;
    int env_w, env_h;
    std::vector<float> env_rgb, env_func, env_cdf, env_func_int, env_marg_func, env_marg_cdf;
    float env_marg_int;
    std::vector<SptTexture> textures;
    std::vector<float> tex_texels;
    std::vector<float> ewa_weight_lut;
    std::vector<SptBrdfTable> brdfs;
    std::vector<SptKdNode> brdf_nodes;
    std::vector<float> brdf_spectra;
    std::vector<float> merl_rgb;

    SptCameraDesc camera;
    SptFilmDesc film;
    SptRenderParams params;
    std::string film_filename;

    LoweredScene();
    SptSceneDesc Desc() const;
    // Container file read by pbrt_v2_spectral_b200/scene_io.py (named little-endian arrays).
    bool Save(const std::string &path, std::string *err) const;
};

// Returns false (and a reason) for anything the GPU path does not implement: the caller must
// then fall back to the reference's SamplerRenderer, never approximate.
bool LowerScene(const Scene *scene, const Camera *camera, const Sampler *sampler,
                const SurfaceIntegrator *surf, LoweredScene *out, std::string *err);

#endif
