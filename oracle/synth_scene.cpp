// synth_scene.cpp — TEST / BENCH INFRASTRUCTURE (not part of the product path).
//
// BASELINE.json config 5 ("synthetic 10M-triangle random-mesh scene, diffuse + glossy materials", recipe: SURVEY.md 8d) built
// through the reference's OWN API (src/core/api.h: pbrtShape + ParamSet) instead of a gigabyte of .pbrt text: the reference's
// classes create the shapes, its BVHAccel builds and flattens the tree (src/accelerators/bvh.cpp:145-372), and the
// `Renderer "gpupath"` hook of oracle_dump.cpp lowers the built scene to <SPT_DUMP_PREFIX>.spt. Links the unmodified
// reference objects; runs wherever oracle/_ref was built (the GPU box included: it needs no source tree).
//
//   SPT_DUMP_PREFIX=out SPT_DUMP_PIXELS=1 SPT_DUMP_LI=0 synth_scene <ntris> <xres> <yres> <spp> [maxdepth [renderer [ncores]]]
// renderer "gpupath" (default) lowers the scene; "sampler" renders it with the reference's own SamplerRenderer on `ncores`
// host cores (bench.py's CPU leg for this workload).
//
// Recipe: triangle = random centre in [-1,1]^3 + two random edge vectors of length U(0.002, 0.02); chunks of <= 1 M triangles,
// even chunks matte (Kd random in [0.2,0.8]^3), odd chunks plastic (Ks .3, roughness U(0.01,0.3)); a sphere area light
// r = 0.2 at (0,0,3) with L = 50 and a constant infinite light L = 0.5; camera at (0,-4,0) looking at the origin, fov 40;
// LD sampler, box filter, path integrator. The random stream is this file's own (SplitMix64), seed 12345.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include "api.h"
#include "paramset.h"
#include "pbrt.h"

static uint64_t g_state = 12345;
static inline uint64_t next64() {
    uint64_t z = (g_state += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static inline double uni() { return (double)(next64() >> 11) * (1.0 / 9007199254740992.0); }
static inline double uni(double a, double b) { return a + (b - a) * uni(); }
static void edge(double e[3]) {                      // uniform direction (rejection in the unit ball), length U(0.002, 0.02)
    double x, y, z, l2;
    do { x = uni(-1, 1); y = uni(-1, 1); z = uni(-1, 1); l2 = x * x + y * y + z * z; } while (l2 > 1.0 || l2 < 1e-12);
    double s = uni(0.002, 0.02) / sqrt(l2);
    e[0] = x * s; e[1] = y * s; e[2] = z * s;
}
static void one(ParamSet &ps, const char *name, int v) { ps.AddInt(name, &v, 1); }
static void one(ParamSet &ps, const char *name, float v) { ps.AddFloat(name, &v, 1); }
static void rgb(ParamSet &ps, const char *name, float r, float g, float b) { float c[3] = { r, g, b }; ps.AddRGBSpectrum(name, c, 3); }

int main(int argc, char *argv[]) {
    if (argc < 5) { fprintf(stderr, "usage: synth_scene <ntris> <xres> <yres> <spp> [maxdepth]\n"); return 2; }
    long ntris = atol(argv[1]);
    int xres = atoi(argv[2]), yres = atoi(argv[3]), spp = atoi(argv[4]), maxdepth = argc > 5 ? atoi(argv[5]) : 5;
    std::string renderer = argc > 6 ? argv[6] : "gpupath";
    Options options;
    options.quiet = true;
    if (argc > 7) options.nCores = atoi(argv[7]);
    pbrtInit(options);
    pbrtLookAt(0, -4, 0, 0, 0, 0, 0, 0, 1);
    { ParamSet ps; one(ps, "fov", 40.f); pbrtCamera("perspective", ps); }
    { ParamSet ps; one(ps, "xresolution", xres); one(ps, "yresolution", yres); std::string fn = "synth.exr"; ps.AddString("filename", &fn, 1); pbrtFilm("image", ps); }
    { ParamSet ps; one(ps, "pixelsamples", spp); pbrtSampler("lowdiscrepancy", ps); }
    { ParamSet ps; pbrtPixelFilter("box", ps); }
    { ParamSet ps; one(ps, "maxdepth", maxdepth); pbrtSurfaceIntegrator("path", ps); }
    { ParamSet ps; pbrtRenderer(renderer, ps); }
    pbrtWorldBegin();
    pbrtAttributeBegin();
    { ParamSet ps; rgb(ps, "L", 50, 50, 50); pbrtAreaLightSource("diffuse", ps); }
    pbrtTranslate(0, 0, 3);
    { ParamSet ps; one(ps, "radius", .2f); pbrtShape("sphere", ps); }
    pbrtAttributeEnd();
    { ParamSet ps; rgb(ps, "L", .5f, .5f, .5f); pbrtLightSource("infinite", ps); }
    const long chunk = 1000000;
    int c = 0;
    for (long done = 0; done < ntris; done += chunk, ++c) {
        long n = ntris - done < chunk ? ntris - done : chunk;
        std::vector<Point> P((size_t)n * 3);
        std::vector<int> idx((size_t)n * 3);
        for (long t = 0; t < n; ++t) {
            double cx = uni(-1, 1), cy = uni(-1, 1), cz = uni(-1, 1), e1[3], e2[3];
            edge(e1); edge(e2);
            P[3 * t] = Point((float)cx, (float)cy, (float)cz);
            P[3 * t + 1] = Point((float)(cx + e1[0]), (float)(cy + e1[1]), (float)(cz + e1[2]));
            P[3 * t + 2] = Point((float)(cx + e2[0]), (float)(cy + e2[1]), (float)(cz + e2[2]));
            idx[3 * t] = (int)(3 * t); idx[3 * t + 1] = (int)(3 * t + 1); idx[3 * t + 2] = (int)(3 * t + 2);
        }
        ParamSet mat;
        rgb(mat, "Kd", (float)uni(.2, .8), (float)uni(.2, .8), (float)uni(.2, .8));
        if (c % 2 == 0) pbrtMaterial("matte", mat);
        else { rgb(mat, "Ks", .3f, .3f, .3f); one(mat, "roughness", (float)uni(.01, .3)); pbrtMaterial("plastic", mat); }
        ParamSet ps;
        ps.AddInt("indices", &idx[0], (int)idx.size());
        ps.AddPoint("P", &P[0], (int)P.size());
        pbrtShape("trianglemesh", ps);
    }
    pbrtWorldEnd();
    pbrtCleanup();
    return 0;
}
