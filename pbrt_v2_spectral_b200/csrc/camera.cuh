// camera.cuh — PerspectiveCamera::GenerateRayDifferential, compiled in the exact (no-FMA) translation unit:
// camera rays match the reference bit for bit.
#pragma once
#include "montecarlo.cuh"

// ---- K1: PerspectiveCamera::GenerateRayDifferential (src/cameras/perspective.cpp:73-106) --------
__device__ inline void camera_ray(const SptCameraDesc &cam, float imageX, float imageY, float lu, float lv, Ray *ray) {
    v3 Pcamera = xf_point(cam.raster_to_camera, V(imageX, imageY, 0.f));
    ray->o = V(0, 0, 0);
    ray->d = normalize(Pcamera);
    ray->mint = 0.f;
    ray->maxt = SPT_INF;
    if (cam.lens_radius > 0.f) {
        float lensU, lensV;
        concentric_sample_disk(lu, lv, &lensU, &lensV);
        lensU *= cam.lens_radius;
        lensV *= cam.lens_radius;
        float ft = cam.focal_distance / ray->d.z;
        v3 Pfocus = ray_at(*ray, ft);
        ray->o = V(lensU, lensV, 0.f);
        ray->d = normalize(vsub(Pfocus, ray->o));
    }
    ray->o = xf_point(cam.camera_to_world, ray->o);
    ray->d = xf_vector(cam.camera_to_world, ray->d);
}

