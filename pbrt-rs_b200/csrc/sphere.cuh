// sphere.cuh — the analytic Sphere shape of the reference (src/shapes/sphere.rs) for the traversal and shading kernels
// (host + device, sm_100a; compile with -fmad=false like everything that includes pb2_math.cuh).
//
// Replaces, per ray / per light sample: Sphere::{intersect_test :250-321, intersect :38-93, intersect_p :95-98, area :100-102,
// sample :104-125, sample2 :127-193, pdf2 :195-207} with the machinery they stand on — EFloat (src/core/efloat.rs), the
// error-carrying Transform applications (src/core/geometry.rs:898-1096, src/core/transform.rs:351-403), Shape::pdf2
// (src/core/shape.rs:54-69).  Where the Rust code cannot work it follows pbrt-v3, the code it is a port of (DESIGN.md defect
// ledger D64-D68): normals transform by the transposed inverse; `Transform * &SurfaceInteraction` (a `//TODO` that returns
// Default::default()) is pbrt-v3's Transform::operator()(SurfaceInteraction); the interaction is flipped by
// reverse_orientation ^ swaps_handedness; Sphere::sample's normal is normalize(o2w * Normal(obj)).
// acos / atan2 — f32::acos / f32::atan2 = platform libm in the reference — are fixed by the numerics contract like sin / cos:
// Cephes asinf / atanf polynomials, every operation one rounded f32 op (det_acos, det_atan2).
// object_to_world must be affine (last row 0 0 0 1): the homogeneous divide of transform.rs:363-367 then never happens
// (w' = 1 exactly) and is not evaluated; pb2_scene_add_spheres rejects anything else.
#pragma once
#include "pb2_math.cuh"

namespace pb2 {

PB2_HD float clamp_f(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }      // pbrt.rs:112-120
PB2_HD float det_asin(float x) {
    const float a = fabsf(x);
    const bool big = a > 0.5f;
    float z, w;
    if (big) { z = 0.5f * (1.0f - a); w = sqrtf(z); }
    else { w = a; z = a * a; }
    float p = ((((4.2163199048e-2f * z + 2.4181311049e-2f) * z + 4.5470025998e-2f) * z + 7.4953002686e-2f) * z + 1.6666752422e-1f) * z * w + w;
    if (big) p = 1.57079632679489661923f - (p + p);
    return x < 0.0f ? -p : p;
}
PB2_HD float det_acos(float x) {
    if (x > 0.5f) return 2.0f * det_asin(sqrtf(0.5f * (1.0f - x)));
    if (x < -0.5f) return PB2_PI - 2.0f * det_asin(sqrtf(0.5f * (1.0f + x)));
    return 1.57079632679489661923f - det_asin(x);
}
PB2_HD float det_atan_pos(float x) {
    float y = 0.0f;
    if (x > 2.414213562373095f) { y = 1.57079632679489661923f; x = -(1.0f / x); }
    else if (x > 0.4142135623730950f) { y = 0.785398163397448309616f; x = (x - 1.0f) / (x + 1.0f); }
    const float z = x * x;
    return y + ((((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z - 3.33329491539e-1f) * z * x + x);
}
PB2_HD float det_atan2(float y, float x) {
    if (x == 0.0f) return y > 0.0f ? 1.57079632679489661923f : (y < 0.0f ? -1.57079632679489661923f : 0.0f);
    const float q = y / x;
    const float a = q < 0.0f ? -det_atan_pos(-q) : det_atan_pos(q);
    if (x > 0.0f) return a;
    return y < 0.0f ? a - PB2_PI : a + PB2_PI;
}

// ---- src/core/efloat.rs: a value with a conservative interval [lo, hi] around it -------------------------------------------
struct efloat {
    float v, lo, hi;
};
PB2_HD efloat ef(float v, float err) {                                   // :15-25
    efloat r;
    r.v = v;
    if (err == 0.0f) { r.lo = v; r.hi = v; }
    else { r.lo = next_down(v - err); r.hi = next_up(v + err); }
    return r;
}
PB2_HD efloat operator+(efloat a, efloat b) { efloat r; r.v = a.v + b.v; r.lo = next_down(a.lo + b.lo); r.hi = next_up(a.hi + b.hi); return r; }
PB2_HD efloat operator-(efloat a, efloat b) { efloat r; r.v = a.v - b.v; r.lo = next_down(a.lo - b.hi); r.hi = next_up(a.hi - b.lo); return r; }
PB2_HD efloat operator*(efloat a, efloat b) {                            // :110-125
    efloat r;
    r.v = a.v * b.v;
    const float p0 = a.lo * b.lo, p1 = a.hi * b.lo, p2 = a.lo * b.hi, p3 = a.hi * b.hi;
    r.lo = next_down(fminf(fminf(fminf(p0, p1), p2), p3));
    r.hi = next_up(fmaxf(fmaxf(fmaxf(p0, p1), p2), p3));
    return r;
}
PB2_HD efloat operator/(efloat a, efloat b) {                            // :127-148
    efloat r;
    r.v = a.v / b.v;
    if (b.lo < 0.0f && b.hi > 0.0f) {
        r.lo = -u2f(0x7f800000u);
        r.hi = u2f(0x7f800000u);
        return r;
    }
    const float d0 = a.lo / b.lo, d1 = a.hi / b.lo, d2 = a.lo / b.hi, d3 = a.hi / b.hi;
    r.lo = next_down(fminf(fminf(fminf(d0, d1), d2), d3));
    r.hi = next_up(fmaxf(fmaxf(fmaxf(d0, d1), d2), d3));
    return r;
}
// :62-83: discriminant in binary64 from the values alone; the root carries MACHINE_EPSILON of error
PB2_HD bool ef_quadratic(efloat a, efloat b, efloat c, efloat* t0, efloat* t1) {
    const double discrim = (double)b.v * (double)b.v - 4.0 * (double)a.v * (double)c.v;
    if (discrim < 0.0) return false;
    const efloat root = ef((float)sqrt(discrim), PB2_MACHINE_EPS);
    const efloat mhalf = ef(-0.5f, 0.0f);
    const efloat q = b.v < 0.0f ? (b - root) * mhalf : (b + root) * mhalf;
    *t0 = q / a;
    *t1 = c / q;
    if (t0->v > t1->v) { const efloat s = *t0; *t0 = *t1; *t1 = s; }
    return true;
}

// ---- the sphere record: 128 bytes = 8 float4 ------------------------------------------------------------------------------
// m / mi: rows 0-2 of object_to_world / world_to_object (row-major, 4 floats per row).
struct DSphere {
    float m[12];
    float mi[12];
    float radius, z_min, z_max, theta_min;
    float theta_max, phi_max;
    uint32_t flags;         // bit 0 reverse_orientation, bit 1 transform swaps handedness
    uint32_t prim;          // the sphere's primitive id
};
static_assert(sizeof(DSphere) == 128, "DSphere must be 128 bytes");

PB2_HD vec3 xf_point(const float* t, vec3 p) {                           // transform.rs:351-369
    return mk(t[0] * p.x + t[1] * p.y + t[2] * p.z + t[3], t[4] * p.x + t[5] * p.y + t[6] * p.z + t[7], t[8] * p.x + t[9] * p.y + t[10] * p.z + t[11]);
}
PB2_HD vec3 xf_vector(const float* t, vec3 v) {                          // transform.rs:371-386
    return mk(t[0] * v.x + t[1] * v.y + t[2] * v.z, t[4] * v.x + t[5] * v.y + t[6] * v.z, t[8] * v.x + t[9] * v.y + t[10] * v.z);
}
PB2_HD vec3 xf_normal(const float* ti, vec3 n) {                         // transform.rs:388-403 with the inverse transposed (D64)
    return mk(ti[0] * n.x + ti[4] * n.y + ti[8] * n.z, ti[1] * n.x + ti[5] * n.y + ti[9] * n.z, ti[2] * n.x + ti[6] * n.y + ti[10] * n.z);
}
PB2_HD vec3 xf_point_err(const float* t, vec3 p, vec3* err) {            // geometry.rs:898-934
    const float xs = fabsf(t[0] * p.x) + fabsf(t[1] * p.y) + fabsf(t[2] * p.z) + fabsf(t[3]);
    const float ys = fabsf(t[4] * p.x) + fabsf(t[5] * p.y) + fabsf(t[6] * p.z) + fabsf(t[7]);
    const float zs = fabsf(t[8] * p.x) + fabsf(t[9] * p.y) + fabsf(t[10] * p.z) + fabsf(t[11]);
    *err = mk(xs, ys, zs) * gammaf_(3.0f);
    return xf_point(t, p);
}
PB2_HD vec3 xf_point_err2(const float* t, vec3 p, vec3 pe, vec3* err) {  // geometry.rs:936-1001
    const float g3 = gammaf_(3.0f);
    const float ex = (g3 + 1.0f) * (fabsf(t[0] * pe.x) + fabsf(t[1] * pe.y) + fabsf(t[2] * pe.z)) +
                     g3 * (fabsf(t[0] * p.x) + fabsf(t[1] * p.y) + fabsf(t[2] * p.z) + fabsf(t[3]));
    const float ey = (g3 + 1.0f) * (fabsf(t[4] * pe.x) + fabsf(t[5] * pe.y) + fabsf(t[6] * pe.z)) +
                     g3 * (fabsf(t[4] * p.x) + fabsf(t[5] * p.y) + fabsf(t[6] * p.z) + fabsf(t[7]));
    const float ez = (g3 + 1.0f) * (fabsf(t[8] * pe.x) + fabsf(t[9] * pe.y) + fabsf(t[10] * pe.z)) +
                     g3 * (fabsf(t[8] * p.x) + fabsf(t[9] * p.y) + fabsf(t[10] * p.z) + fabsf(t[11]));
    *err = mk(ex, ey, ez);
    return xf_point(t, p);
}
PB2_HD vec3 xf_vector_err(const float* t, vec3 v, vec3* err) {           // geometry.rs:1003-1024
    const float g3 = gammaf_(3.0f);
    *err = mk(g3 * (fabsf(t[0] * v.x) + fabsf(t[1] * v.y) + fabsf(t[2] * v.z)), g3 * (fabsf(t[4] * v.x) + fabsf(t[5] * v.y) + fabsf(t[6] * v.z)),
              g3 * (fabsf(t[8] * v.x) + fabsf(t[9] * v.y) + fabsf(t[10] * v.z)));
    return xf_vector(t, v);
}

PB2_HD float sphere_area(const DSphere& s) { return s.phi_max * s.radius * (s.z_max - s.z_min); }     // sphere.rs:100-102

// What Sphere::intersect leaves in the SurfaceInteraction, world space — the fields the path reads.
struct SphereVertex {
    vec3 p, err, n, wo, dpdu, sn;
    float u, v;
};

// sphere.rs:250-321: the object-space ray, the refined hit point, phi and t of the first root the ray accepts.
PB2_HD bool sphere_test(const DSphere& s, vec3 o, vec3 d, float t_max, vec3* p_hit_out, float* phi_out, vec3* od_out, float* t_out) {
    vec3 o_err, d_err;
    vec3 ro = xf_point_err(s.mi, o, &o_err);                             // geometry.rs:1077-1096
    const vec3 rd = xf_vector_err(s.mi, d, &d_err);
    const float l2 = len2(rd);
    if (l2 > 0.0f) {
        const float dt = dot3(abs3(rd), o_err) / l2;
        ro = ro + rd * dt;
    }
    const efloat ox = ef(ro.x, o_err.x), oy = ef(ro.y, o_err.y), oz = ef(ro.z, o_err.z);
    const efloat dx = ef(rd.x, d_err.x), dy = ef(rd.y, d_err.y), dz = ef(rd.z, d_err.z);
    const efloat a = dx * dx + dy * dy + dz * dz;
    const efloat b = (dx * ox + dy * oy + dz * oz) * ef(2.0f, 0.0f);
    const efloat c = ox * ox + oy * oy + oz * oz - ef(s.radius, 0.0f) * ef(s.radius, 0.0f);
    efloat t0, t1;
    if (!ef_quadratic(a, b, c, &t0, &t1)) return false;
    for (int k = 0; k < 2; ++k) {
        const efloat t = k == 0 ? t0 : t1;
        if (t.lo < 0.0f || t.hi > t_max) continue;
        vec3 p_hit = ro + rd * t.v;
        p_hit = p_hit * (s.radius / len(p_hit));
        if (p_hit.x == 0.0f && p_hit.y == 0.0f) p_hit.x = 1e-5f * s.radius;
        float phi = det_atan2(p_hit.y, p_hit.x);
        if (phi < 0.0f) phi += 2.0f * PB2_PI;
        if ((s.z_min > -s.radius && p_hit.z < s.z_min) || (s.z_max < s.radius && p_hit.z > s.z_max) || phi > s.phi_max) continue;
        *p_hit_out = p_hit;
        *phi_out = phi;
        *od_out = rd;
        *t_out = t.v;
        return true;
    }
    return false;
}
// sphere.rs:38-93 after intersect_test: the interaction in world space.
PB2_HD SphereVertex sphere_vertex(const DSphere& s, vec3 p_hit, float phi, vec3 obj_d) {
    SphereVertex r;
    r.u = phi / s.phi_max;
    const float theta = det_acos(clamp_f(p_hit.z / s.radius, -1.0f, 1.0f));
    r.v = (theta - s.theta_min) / (s.theta_max - s.theta_min);
    const float z_radius = sqrtf(p_hit.x * p_hit.x + p_hit.y * p_hit.y);
    const float inv_z_radius = 1.0f / z_radius;
    const float cos_phi = p_hit.x * inv_z_radius, sin_phi = p_hit.y * inv_z_radius;
    const vec3 dpdu = mk(-s.phi_max * p_hit.y, s.phi_max * p_hit.x, 0.0f);
    const vec3 dpdv = mk(p_hit.z * cos_phi, p_hit.z * sin_phi, -s.radius * det_sin(theta)) * (s.theta_max - s.theta_min);
    const vec3 p_error = abs3(p_hit) * gammaf_(5.0f);
    vec3 n = unit(cross3(dpdu, dpdv));                                   // interaction.rs:262-268, flipped as :285-290 would with a shape (D66)
    if (((s.flags & 1u) != 0u) != ((s.flags & 2u) != 0u)) n = -n;
    r.p = xf_point_err2(s.m, p_hit, p_error, &r.err);                    // pbrt-v3 Transform::operator()(SurfaceInteraction) (D65)
    r.n = unit(xf_normal(s.mi, n));
    r.wo = unit(xf_vector(s.m, -obj_d));
    r.dpdu = xf_vector(s.m, dpdu);
    r.sn = unit(xf_normal(s.mi, n));
    if (dot3(r.sn, r.n) < 0.0f) r.sn = -r.sn;
    return r;
}
PB2_HD bool sphere_intersect(const DSphere& s, vec3 o, vec3 d, float t_max, float* t_out, SphereVertex* v) {
    vec3 p_hit, od;
    float phi;
    if (!sphere_test(s, o, d, t_max, &p_hit, &phi, &od, t_out)) return false;
    *v = sphere_vertex(s, p_hit, phi, od);
    return true;
}
// The interaction of a hit already found by the traversal at distance t (the walk keeps {prim, t, u, v} only).
PB2_HD SphereVertex sphere_vertex_at(const DSphere& s, vec3 o, vec3 d, float t) {
    vec3 o_err, d_err;
    vec3 ro = xf_point_err(s.mi, o, &o_err);
    const vec3 rd = xf_vector_err(s.mi, d, &d_err);
    const float l2 = len2(rd);
    if (l2 > 0.0f) {
        const float dt = dot3(abs3(rd), o_err) / l2;
        ro = ro + rd * dt;
    }
    vec3 p_hit = ro + rd * t;
    p_hit = p_hit * (s.radius / len(p_hit));
    if (p_hit.x == 0.0f && p_hit.y == 0.0f) p_hit.x = 1e-5f * s.radius;
    float phi = det_atan2(p_hit.y, p_hit.x);
    if (phi < 0.0f) phi += 2.0f * PB2_PI;
    return sphere_vertex(s, p_hit, phi, rd);
}

// sphere.rs:104-125 Sphere::sample (D67)
PB2_HD void sphere_sample(const DSphere& s, float u0, float u1, vec3* p, vec3* p_err, vec3* n, float* pdf) {
    const float z = 1.0f - 2.0f * u0;                                    // sampling.rs:230-235
    const float r = sqrtf(fmaxf(1.0f - z * z, 0.0f));
    const float phi = 2.0f * PB2_PI * u1;
    float sn, cs;
    det_sincos(phi, &sn, &cs);
    vec3 obj = mk(0.f, 0.f, 0.f) + mk(r * cs, r * sn, z) * s.radius;
    vec3 nn = unit(xf_normal(s.mi, obj));
    if (s.flags & 1u) nn = nn * -1.0f;
    obj = obj * (s.radius / len(obj));
    const vec3 obj_err = abs3(obj) * gammaf_(5.0f);
    *p = xf_point_err2(s.m, obj, obj_err, p_err);
    *n = nn;
    *pdf = 1.0f / sphere_area(s);
}
// sphere.rs:127-193 Sphere::sample2
PB2_HD void sphere_sample2(const DSphere& s, vec3 rp, vec3 rerr, vec3 rn, float u0, float u1, vec3* p, vec3* p_err, vec3* n, float* pdf) {
    const vec3 p_center = xf_point(s.m, mk(0.f, 0.f, 0.f));
    const vec3 p_origin = offset_ray_origin(rp, rerr, rn, p_center - rp);
    if (len2(p_origin - p_center) <= s.radius * s.radius) {
        sphere_sample(s, u0, u1, p, p_err, n, pdf);
        vec3 wi = *p - rp;
        if (len2(wi) == 0.0f) *pdf = 0.0f;
        else {
            wi = unit(wi);
            *pdf *= len2(rp - *p) / fabsf(dot3(*n, -wi));
        }
        if (isinf(*pdf)) *pdf = 0.0f;
        return;
    }
    const float dc = len(rp - p_center);
    const float inv_dc = 1.0f / dc;
    const vec3 wc = (p_center - rp) * inv_dc;
    vec3 wc_x, wc_y;
    coord_system(wc, &wc_x, &wc_y);
    const float sin_theta_max = s.radius * inv_dc;
    const float sin_theta_max2 = sin_theta_max * sin_theta_max;
    const float inv_sin_theta_max = 1.0f / sin_theta_max;
    const float cos_theta_max = sqrtf(fmaxf(1.0f - sin_theta_max2, 0.0f));
    float cos_theta = (cos_theta_max - 1.0f) * u0 + 1.0f;
    float sin_theta2 = 1.0f - cos_theta * cos_theta;
    if (sin_theta_max2 < 0.00068523f) {
        sin_theta2 = sin_theta_max2 * u0;
        cos_theta = sqrtf(1.0f - sin_theta2);
    }
    const float cos_alpha = sin_theta2 * inv_sin_theta_max + cos_theta * sqrtf(fmaxf(1.0f - sin_theta2 * inv_sin_theta_max * inv_sin_theta_max, 0.0f));
    const float sin_alpha = sqrtf(fmaxf(1.0f - cos_alpha * cos_alpha, 0.0f));
    const float phi = u1 * 2.0f * PB2_PI;
    float sn, cs;
    det_sincos(phi, &sn, &cs);
    const vec3 n_world = ((-wc_x) * sin_alpha * cs + (-wc_y) * sin_alpha * sn) + (-wc) * cos_alpha;    // geometry.rs:1156-1165
    const vec3 p_world = p_center + n_world * s.radius;
    *p = p_world;
    *p_err = abs3(p_world) * gammaf_(5.0f);
    *n = (s.flags & 1u) ? n_world * -1.0f : n_world;
    *pdf = 1.0f / (2.0f * PB2_PI * (1.0f - cos_theta_max));
}
// sphere.rs:195-207 Sphere::pdf2, with Shape::pdf2 (shape.rs:54-69) for a reference point inside the sphere
PB2_HD float sphere_pdf2(const DSphere& s, vec3 rp, vec3 rerr, vec3 rn, vec3 wi) {
    const vec3 p_center = xf_point(s.m, mk(0.f, 0.f, 0.f));
    const vec3 p_origin = offset_ray_origin(rp, rerr, rn, p_center - rp);
    if (len2(p_origin - p_center) < s.radius * s.radius) {
        const vec3 o = offset_ray_origin(rp, rerr, rn, wi);
        float t;
        SphereVertex li;
        if (!sphere_intersect(s, o, wi, u2f(0x7f800000u), &t, &li)) return 0.0f;
        float pdf = len2(rp - li.p) / (fabsf(dot3(li.n, -wi)) * sphere_area(s));
        if (isinf(pdf)) pdf = 0.0f;
        return pdf;
    }
    const float sin_theta_max2 = s.radius * s.radius / len2(rp - p_center);
    const float cos_theta_max = sqrtf(fmaxf(1.0f - sin_theta_max2, 0.0f));
    return 1.0f / (2.0f * PB2_PI * (1.0f - cos_theta_max));
}

}  // namespace pb2
