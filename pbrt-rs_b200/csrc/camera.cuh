// camera.cuh — PerspectiveCamera::generate_ray on the device (perspective.rs:90-112), shared by k_camera_rays (batched
// Camera::generate_ray) and k_raygen (the wavefront's ray generation): raster point -> camera space, thin lens
// (:101-107), camera-to-world with the origin's error bound (geometry.rs:865-881, 898-936).
#pragma once
#include "kernels.hpp"
#include "shade.cuh"

namespace pb2 {

PB2_D vec3 xf_point(const mat4& m, vec3 p) {                                  // transform.rs:351-369
    const float xp = ((m.m[0][0] * p.x + m.m[0][1] * p.y) + m.m[0][2] * p.z) + m.m[0][3];
    const float yp = ((m.m[1][0] * p.x + m.m[1][1] * p.y) + m.m[1][2] * p.z) + m.m[1][3];
    const float zp = ((m.m[2][0] * p.x + m.m[2][1] * p.y) + m.m[2][2] * p.z) + m.m[2][3];
    const float wp = ((m.m[3][0] * p.x + m.m[3][1] * p.y) + m.m[3][2] * p.z) + m.m[3][3];
    if (wp == 1.0f) return mk(xp, yp, zp);
    return mk(xp, yp, zp) / wp;
}
PB2_D vec3 xf_point_err(const mat4& m, vec3 p, vec3* err) {                   // geometry.rs:898-936
    const float xp = ((m.m[0][0] * p.x + m.m[0][1] * p.y) + m.m[0][2] * p.z) + m.m[0][3];
    const float yp = ((m.m[1][0] * p.x + m.m[1][1] * p.y) + m.m[1][2] * p.z) + m.m[1][3];
    const float zp = ((m.m[2][0] * p.x + m.m[2][1] * p.y) + m.m[2][2] * p.z) + m.m[2][3];
    const float wp = ((m.m[3][0] * p.x + m.m[3][1] * p.y) + m.m[3][2] * p.z) + m.m[3][3];
    const float xs = ((fabsf(m.m[0][0] * p.x) + fabsf(m.m[0][1] * p.y)) + fabsf(m.m[0][2] * p.z)) + fabsf(m.m[0][3]);
    const float ys = ((fabsf(m.m[1][0] * p.x) + fabsf(m.m[1][1] * p.y)) + fabsf(m.m[1][2] * p.z)) + fabsf(m.m[1][3]);
    const float zs = ((fabsf(m.m[2][0] * p.x) + fabsf(m.m[2][1] * p.y)) + fabsf(m.m[2][2] * p.z)) + fabsf(m.m[2][3]);
    *err = mk(xs, ys, zs) * gammaf_(3.0f);
    if (wp == 1.0f) return mk(xp, yp, zp);
    return mk(xp, yp, zp) / wp;
}
PB2_D vec3 xf_vector(const mat4& m, vec3 v) {                                 // transform.rs:371-386
    return mk((m.m[0][0] * v.x + m.m[0][1] * v.y) + m.m[0][2] * v.z,
              (m.m[1][0] * v.x + m.m[1][1] * v.y) + m.m[1][2] * v.z,
              (m.m[2][0] * v.x + m.m[2][1] * v.y) + m.m[2][2] * v.z);
}

PB2_D void camera_ray(const CameraView& cam, float fx, float fy, float lx, float ly, vec3* o_out, vec3* d_out, float* t_max_out) {
    const vec3 p_camera = xf_point(cam.raster_to_camera, mk(fx, fy, 0.0f));
    vec3 d_cam = unit(p_camera);
    vec3 o_cam = mk(0.0f, 0.0f, 0.0f);
    if (cam.lens_radius > 0.0f) {                                             // perspective.rs:101-107
        float px, py;
        concentric_disk(lx, ly, &px, &py);
        px = px * cam.lens_radius; py = py * cam.lens_radius;
        const float ft = cam.focal_distance / d_cam.z;
        const vec3 p_focus = o_cam + d_cam * ft;
        o_cam = mk(px, py, 0.0f);
        d_cam = unit(p_focus - o_cam);
    }
    vec3 o_err;
    vec3 o = xf_point_err(cam.camera_to_world, o_cam, &o_err);
    const vec3 d = xf_vector(cam.camera_to_world, d_cam);
    const float ls = len2(d);
    float t_max = __int_as_float(0x7f800000);
    if (ls > 0.0f) {
        const float dt = dot3(abs3(d), o_err) / ls;
        o = o + d * dt;
        t_max = t_max - dt;
    }
    *o_out = o;
    *d_out = d;
    *t_max_out = t_max;
}

}  // namespace pb2
