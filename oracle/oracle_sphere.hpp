// ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.  (Included by oracle_core.hpp; see its header.)  PARITY UNPINNED.
//
// CPU restatement of the analytic Sphere (src/shapes/sphere.rs) with the pieces it stands on: EFloat running error
// intervals (src/core/efloat.rs), the error-carrying Transform applications (src/core/geometry.rs:898-1096) and Shape's
// default pdf2 (src/core/shape.rs:54-69).  KEEP/FIX ledger of this file (pbrt-v3 semantics where the Rust code cannot work):
//   D64 FIX  transform.rs:388-403   `Transform * Normal3Ref` multiplies by m_inv, not by its transpose            -> transpose
//   D65 FIX  transform.rs:620-627   `Transform * &SurfaceInteraction` is `//TODO Default::default()`                -> pbrt-v3 Transform::operator()(SurfaceInteraction)
//   D66 FIX  sphere.rs:89           the interaction is built with shape = None ("FIXME shape"), so reverse_orientation ^
//                                   swaps_handedness never flips n (interaction.rs:285-290)                         -> flip as with a shape
//   D67 FIX  sphere.rs:111          Sphere::sample transforms the normal as a point and does not normalise it       -> normalize(o2w * Normal(obj))
//   D68 FIX  efloat.rs:27-29        get_absolute_error's parentheses (`.max(..).abs()`)                             -> unused on this path
//   KEEP     sphere.rs:183-186      sample2 / pdf2 disagree on `<=` vs `<` for "inside" — as written (pbrt-v3 has the same)
//   KEEP     efloat.rs:62-83        quadratic: discriminant in f64 from the .v fields only, root error = MACHINE_EPSILON
// acos / atan2: the reference calls f32::acos / f32::atan2 (platform libm, last bit platform dependent); like sin / cos the
// numerics contract fixes one definition shared by oracle and kernels (Cephes asinf / atanf polynomials, every op
// separately rounded f32): acos_c, atan2_c below.
#pragma once

inline void sincos_contract(Float x, Float* s_out, Float* c_out);      // oracle_core.hpp, below BVHAccel

// ---------------------------------------------------------------- contract acos / atan2
inline Float asin_c(Float x) {
    const Float a = std::fabs(x);
    const bool big = a > 0.5f;
    Float z, w;
    if (big) { z = 0.5f * (1.0f - a); w = std::sqrt(z); }
    else { w = a; z = a * a; }
    Float p = ((((4.2163199048e-2f * z + 2.4181311049e-2f) * z + 4.5470025998e-2f) * z + 7.4953002686e-2f) * z + 1.6666752422e-1f) * z * w + w;
    if (big) p = 1.57079632679489661923f - (p + p);
    return x < 0.0f ? -p : p;
}
inline Float acos_c(Float x) {                                         // x in [-1, 1] (callers clamp)
    if (x > 0.5f) return 2.0f * asin_c(std::sqrt(0.5f * (1.0f - x)));
    if (x < -0.5f) return kPi - 2.0f * asin_c(std::sqrt(0.5f * (1.0f + x)));
    return 1.57079632679489661923f - asin_c(x);
}
inline Float atan_pos_c(Float x) {                                     // x >= 0
    Float y = 0.0f;
    if (x > 2.414213562373095f) { y = 1.57079632679489661923f; x = -(1.0f / x); }
    else if (x > 0.4142135623730950f) { y = 0.785398163397448309616f; x = (x - 1.0f) / (x + 1.0f); }
    const Float z = x * x;
    return y + ((((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z - 3.33329491539e-1f) * z * x + x);
}
inline Float atan2_c(Float y, Float x) {                               // result in (-pi, pi]
    if (x == 0.0f) return y > 0.0f ? 1.57079632679489661923f : (y < 0.0f ? -1.57079632679489661923f : 0.0f);
    const Float q = y / x;
    const Float a = q < 0.0f ? -atan_pos_c(-q) : atan_pos_c(q);
    if (x > 0.0f) return a;
    return y < 0.0f ? a - kPi : a + kPi;
}

// ---------------------------------------------------------------- src/core/efloat.rs
struct EFloat {
    Float v = 0, low = 0, high = 0;
    EFloat() = default;
    EFloat(Float v_, Float err) : v(v_) {                              // :15-25
        if (err == 0.0f) { low = v_; high = v_; }
        else { low = next_float_down(v_ - err); high = next_float_up(v_ + err); }
    }
};
inline EFloat operator+(EFloat a, EFloat b) {                          // :86-96
    EFloat r; r.v = a.v + b.v; r.low = next_float_down(a.low + b.low); r.high = next_float_up(a.high + b.high); return r;
}
inline EFloat operator-(EFloat a, EFloat b) {                          // :98-108
    EFloat r; r.v = a.v - b.v; r.low = next_float_down(a.low - b.high); r.high = next_float_up(a.high - b.low); return r;
}
inline EFloat operator*(EFloat a, EFloat b) {                          // :110-125 (f32::min / max chains, left to right)
    EFloat r; r.v = a.v * b.v;
    const Float p0 = a.low * b.low, p1 = a.high * b.low, p2 = a.low * b.high, p3 = a.high * b.high;
    r.low = next_float_down(fmin_(fmin_(fmin_(p0, p1), p2), p3));
    r.high = next_float_up(fmax_(fmax_(fmax_(p0, p1), p2), p3));
    return r;
}
inline EFloat operator/(EFloat a, EFloat b) {                          // :127-148
    EFloat r; r.v = a.v / b.v;
    if (b.low < 0.0f && b.high > 0.0f) { r.low = -kInfinity; r.high = kInfinity; return r; }
    const Float d0 = a.low / b.low, d1 = a.high / b.low, d2 = a.low / b.high, d3 = a.high / b.high;
    r.low = next_float_down(fmin_(fmin_(fmin_(d0, d1), d2), d3));
    r.high = next_float_up(fmax_(fmax_(fmax_(d0, d1), d2), d3));
    return r;
}
inline EFloat operator*(EFloat a, Float s) { return a * EFloat(s, 0.0f); }   // :185-191
// :62-83
inline bool efloat_quadratic(EFloat a, EFloat b, EFloat c, EFloat* t0, EFloat* t1) {
    const double discrim = (double)b.v * (double)b.v - 4.0 * (double)a.v * (double)c.v;
    if (discrim < 0.0) return false;
    const double root_discrim = std::sqrt(discrim);
    const EFloat float_root_discrim((Float)root_discrim, kMachineEpsilon);
    const EFloat q = b.v < 0.0f ? (b - float_root_discrim) * -0.5f : (b + float_root_discrim) * -0.5f;
    *t0 = q / a;
    *t1 = c / q;
    if (t0->v > t1->v) std::swap(*t0, *t1);
    return true;
}

// ---------------------------------------------------------------- Transform applications (row-major 4x4, affine: w' == 1)
struct Mat4 { Float m[4][4]; };
inline V3 xf_point(const Mat4& t, V3 p) {                              // transform.rs:351-369
    return {t.m[0][0] * p.x + t.m[0][1] * p.y + t.m[0][2] * p.z + t.m[0][3], t.m[1][0] * p.x + t.m[1][1] * p.y + t.m[1][2] * p.z + t.m[1][3],
            t.m[2][0] * p.x + t.m[2][1] * p.y + t.m[2][2] * p.z + t.m[2][3]};
}
inline V3 xf_vector(const Mat4& t, V3 v) {                             // transform.rs:371-386
    return {t.m[0][0] * v.x + t.m[0][1] * v.y + t.m[0][2] * v.z, t.m[1][0] * v.x + t.m[1][1] * v.y + t.m[1][2] * v.z,
            t.m[2][0] * v.x + t.m[2][1] * v.y + t.m[2][2] * v.z};
}
inline V3 xf_normal(const Mat4& t_inv, V3 n) {                         // transform.rs:388-403, D64 FIX: the inverse, transposed
    return {t_inv.m[0][0] * n.x + t_inv.m[1][0] * n.y + t_inv.m[2][0] * n.z, t_inv.m[0][1] * n.x + t_inv.m[1][1] * n.y + t_inv.m[2][1] * n.z,
            t_inv.m[0][2] * n.x + t_inv.m[1][2] * n.y + t_inv.m[2][2] * n.z};
}
inline V3 xf_point_err(const Mat4& t, V3 p, V3* err) {                 // geometry.rs:898-934
    const Float xs = std::fabs(t.m[0][0] * p.x) + std::fabs(t.m[0][1] * p.y) + std::fabs(t.m[0][2] * p.z) + std::fabs(t.m[0][3]);
    const Float ys = std::fabs(t.m[1][0] * p.x) + std::fabs(t.m[1][1] * p.y) + std::fabs(t.m[1][2] * p.z) + std::fabs(t.m[1][3]);
    const Float zs = std::fabs(t.m[2][0] * p.x) + std::fabs(t.m[2][1] * p.y) + std::fabs(t.m[2][2] * p.z) + std::fabs(t.m[2][3]);
    *err = V3{xs, ys, zs} * gamma(3.0f);
    return xf_point(t, p);
}
inline V3 xf_point_err2(const Mat4& t, V3 p, V3 pe, V3* err) {         // geometry.rs:936-1001 (a point that carries an error already)
    const Float g3 = gamma(3.0f);
    Float e[3];
    const Float c[3] = {p.x, p.y, p.z};
    for (int i = 0; i < 3; ++i)
        e[i] = (g3 + 1.0f) * (std::fabs(t.m[i][0] * pe.x) + std::fabs(t.m[i][1] * pe.y) + std::fabs(t.m[i][2] * pe.z)) +
               g3 * (std::fabs(t.m[i][0] * c[0]) + std::fabs(t.m[i][1] * c[1]) + std::fabs(t.m[i][2] * c[2]) + std::fabs(t.m[i][3]));
    *err = V3{e[0], e[1], e[2]};
    return xf_point(t, p);
}
inline V3 xf_vector_err(const Mat4& t, V3 v, V3* err) {                // geometry.rs:1003-1024
    const Float g3 = gamma(3.0f);
    *err = V3{g3 * (std::fabs(t.m[0][0] * v.x) + std::fabs(t.m[0][1] * v.y) + std::fabs(t.m[0][2] * v.z)),
              g3 * (std::fabs(t.m[1][0] * v.x) + std::fabs(t.m[1][1] * v.y) + std::fabs(t.m[1][2] * v.z)),
              g3 * (std::fabs(t.m[2][0] * v.x) + std::fabs(t.m[2][1] * v.y) + std::fabs(t.m[2][2] * v.z))};
    return xf_vector(t, v);
}
// geometry.rs:1077-1096: Ray::from((transform, ray, &mut o_err, &mut d_err))
inline Ray xf_ray_err(const Mat4& t, const Ray& r, V3* o_err, V3* d_err) {
    V3 o = xf_point_err(t, r.o, o_err);
    const V3 d = xf_vector_err(t, r.d, d_err);
    const Float l2 = length_squared(d);
    if (l2 > 0.0f) {
        const Float dt = dot(vabs(d), *o_err) / l2;
        o = o + d * dt;
    }
    return Ray{o, r.t_max, d, r.time};
}

// ---------------------------------------------------------------- src/shapes/sphere.rs
struct SphereDesc {                 // mirrors pb2_sphere (include/pbrt_b200.h)
    Float object_to_world[16];      // row-major, affine (last row 0 0 0 1)
    Float radius, z_min, z_max, phi_max;   // Sphere::new's arguments (phi_max in degrees)
    int32_t reverse_orientation;
    uint32_t material;
};
struct SphereSI {                   // what Sphere::intersect leaves in the SurfaceInteraction (world space) — the subset the path reads
    V3 p, error, n, wo, dpdu, sn, sdpdu;
    Float u, v;
};
struct Sphere {
    Mat4 o2w, w2o;
    Float radius, z_min, z_max, theta_min, theta_max, phi_max;
    bool reverse_orientation, swaps_handedness;
    uint32_t material;

    void init(const SphereDesc& d, const Mat4& inverse) {              // sphere.rs:229-248
        std::memcpy(o2w.m, d.object_to_world, sizeof(o2w.m));
        w2o = inverse;
        radius = d.radius;
        z_min = clampf_(fmin_(d.z_min, d.z_max), -radius, radius);
        z_max = clampf_(fmax_(d.z_max, d.z_min), -radius, radius);
        theta_min = acos_c(clampf_(fmin_(d.z_min, d.z_max) / radius, -1.0f, 1.0f));
        theta_max = acos_c(clampf_(fmax_(d.z_max, d.z_min) / radius, -1.0f, 1.0f));
        phi_max = kPi / 180.0f * clampf_(d.phi_max, 0.0f, 360.0f);     // pbrt.rs:133-135
        reverse_orientation = d.reverse_orientation != 0;
        const Float (*m)[4] = o2w.m;                                   // pbrt-v3 Transform::SwapsHandedness: det of the 3x3 block < 0
        const Float det = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
                          m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
        swaps_handedness = det < 0.0f;
        material = d.material;
    }
    static Float clampf_(Float v, Float lo, Float hi) { return v < lo ? lo : (v > hi ? hi : v); }   // pbrt.rs:112-120

    Bounds3 world_bound() const {                                      // shape.rs:18-20, sphere.rs:31-36, transform.rs:569-606
        const V3 lo{-radius, -radius, z_min}, hi{radius, radius, z_max};
        const V3 c0 = xf_point(o2w, V3{lo.x, lo.y, lo.z});
        Bounds3 b{c0, c0};
        b = bunion(b, xf_point(o2w, V3{hi.x, lo.y, lo.z}));
        b = bunion(b, xf_point(o2w, V3{lo.x, hi.y, lo.z}));
        b = bunion(b, xf_point(o2w, V3{lo.x, lo.y, hi.z}));
        b = bunion(b, xf_point(o2w, V3{lo.x, hi.y, hi.z}));
        b = bunion(b, xf_point(o2w, V3{hi.x, hi.y, lo.z}));
        b = bunion(b, xf_point(o2w, V3{hi.x, lo.y, hi.z}));
        b = bunion(b, xf_point(o2w, V3{hi.x, hi.y, hi.z}));
        return b;
    }
    Float area() const { return phi_max * radius * (z_max - z_min); }  // sphere.rs:100-102

    // sphere.rs:250-321 intersect_test: (hit, p_hit, phi, object-space ray, t)
    bool intersect_test(const Ray& r, V3* p_hit_out, Float* phi_out, Ray* ray_out, Float* t_out) const {
        V3 o_err, d_err;
        const Ray ray = xf_ray_err(w2o, r, &o_err, &d_err);
        const EFloat ox(ray.o.x, o_err.x), oy(ray.o.y, o_err.y), oz(ray.o.z, o_err.z);
        const EFloat dx(ray.d.x, d_err.x), dy(ray.d.y, d_err.y), dz(ray.d.z, d_err.z);
        const EFloat a = dx * dx + dy * dy + dz * dz;
        const EFloat b = (dx * ox + dy * oy + dz * oz) * 2.0f;
        const EFloat c = ox * ox + oy * oy + oz * oz - EFloat(radius, 0.0f) * EFloat(radius, 0.0f);
        EFloat t0, t1;
        if (!efloat_quadratic(a, b, c, &t0, &t1)) return false;
        const EFloat ts[2] = {t0, t1};
        for (const EFloat& t : ts) {
            if (t.low < 0.0f || t.high > ray.t_max) continue;
            V3 p_hit = ray.o + ray.d * t.v;                            // geometry.rs:794-796
            p_hit = p_hit * (radius / length(p_hit));
            if (p_hit.x == 0.0f && p_hit.y == 0.0f) p_hit.x = 1e-5f * radius;
            Float phi = atan2_c(p_hit.y, p_hit.x);
            if (phi < 0.0f) phi += 2.0f * kPi;
            if ((z_min > -radius && p_hit.z < z_min) || (z_max < radius && p_hit.z > z_max) || phi > phi_max) continue;
            *p_hit_out = p_hit; *phi_out = phi; *ray_out = ray; *t_out = t.v;
            return true;
        }
        return false;
    }
    bool intersect_p(const Ray& r) const {                             // sphere.rs:95-98
        V3 p; Float phi, t; Ray ray;
        return intersect_test(r, &p, &phi, &ray, &t);
    }
    // sphere.rs:38-93 + interaction.rs:248-293 + pbrt-v3 Transform::operator()(SurfaceInteraction) (D65, D66 FIX)
    bool intersect(const Ray& r, Float* t_hit, SphereSI* si) const {
        V3 p_hit; Float phi, t; Ray ray;
        if (!intersect_test(r, &p_hit, &phi, &ray, &t)) return false;
        const Float u = phi / phi_max;
        const Float theta = acos_c(clampf_(p_hit.z / radius, -1.0f, 1.0f));
        const Float v = (theta - theta_min) / (theta_max - theta_min);
        const Float z_radius = std::sqrt(p_hit.x * p_hit.x + p_hit.y * p_hit.y);
        const Float inv_z_radius = 1.0f / z_radius;
        const Float cos_phi = p_hit.x * inv_z_radius, sin_phi = p_hit.y * inv_z_radius;
        const V3 dpdu{-phi_max * p_hit.y, phi_max * p_hit.x, 0.0f};
        Float sin_theta, cos_theta_unused;
        sincos_contract(theta, &sin_theta, &cos_theta_unused);
        const V3 dpdv = V3{p_hit.z * cos_phi, p_hit.z * sin_phi, -radius * sin_theta} * (theta_max - theta_min);
        const V3 p_error = vabs(p_hit) * gamma(5.0f);
        // SurfaceInteraction::new: n = normalize(cross(dpdu, dpdv)), flipped by reverse_orientation ^ swaps_handedness; shading = geometric
        V3 n = normalize(cross(dpdu, dpdv));
        if (reverse_orientation != swaps_handedness) n = -n;
        const V3 wo = -ray.d;
        // object_to_world * si
        si->p = xf_point_err2(o2w, p_hit, p_error, &si->error);
        si->n = normalize(xf_normal(w2o, n));
        si->wo = normalize(xf_vector(o2w, wo));
        si->dpdu = xf_vector(o2w, dpdu);
        si->sdpdu = si->dpdu;
        si->sn = normalize(xf_normal(w2o, n));
        si->sn = dot(si->sn, si->n) < 0.0f ? -si->sn : si->sn;         // Faceforward(shading.n, n)
        si->u = u; si->v = v;
        *t_hit = t;
        return true;
    }
    // sphere.rs:104-125 Sphere::sample (D67 FIX)
    void sample(Float u0, Float u1, V3* p, V3* p_err, V3* n, Float* pdf) const {
        const Float z = 1.0f - 2.0f * u0;                              // sampling.rs:230-235
        const Float r = std::sqrt(fmax_(1.0f - z * z, 0.0f));
        const Float phi = 2.0f * kPi * u1;
        Float s, c;
        sincos_contract(phi, &s, &c);
        V3 obj = V3{0, 0, 0} + V3{r * c, r * s, z} * radius;
        V3 nn = normalize(xf_normal(w2o, obj));
        if (reverse_orientation) nn = nn * -1.0f;
        obj = obj * (radius / length(obj));
        const V3 obj_err = vabs(obj) * gamma(5.0f);
        *p = xf_point_err2(o2w, obj, obj_err, p_err);
        *n = nn;
        *pdf = 1.0f / area();
    }
    // sphere.rs:127-193 Sphere::sample2: sample a point on the sphere as seen from (rp, rerr, rn)
    void sample2(V3 rp, V3 rerr, V3 rn, Float u0, Float u1, V3* p, V3* p_err, V3* n, Float* pdf) const {
        const V3 p_center = xf_point(o2w, V3{0, 0, 0});
        const V3 p_origin = offset_ray_origin_(rp, rerr, rn, p_center - rp);
        if (length_squared(p_origin - p_center) <= radius * radius) {
            sample(u0, u1, p, p_err, n, pdf);
            V3 wi = *p - rp;
            if (length_squared(wi) == 0.0f) *pdf = 0.0f;
            else {
                wi = normalize(wi);
                *pdf *= length_squared(rp - *p) / std::fabs(dot(*n, -wi));
            }
            if (std::isinf(*pdf)) *pdf = 0.0f;
            return;
        }
        const Float dc = length(rp - p_center);
        const Float inv_dc = 1.0f / dc;
        const V3 wc = (p_center - rp) * inv_dc;
        V3 wc_x, wc_y;
        coordinate_system(wc, &wc_x, &wc_y);
        const Float sin_theta_max = radius * inv_dc;
        const Float sin_theta_max2 = sin_theta_max * sin_theta_max;
        const Float inv_sin_theta_max = 1.0f / sin_theta_max;
        const Float cos_theta_max = std::sqrt(fmax_(1.0f - sin_theta_max2, 0.0f));
        Float cos_theta = (cos_theta_max - 1.0f) * u0 + 1.0f;
        Float sin_theta2 = 1.0f - cos_theta * cos_theta;
        if (sin_theta_max2 < 0.00068523f) {                            // sin^2(1.5 deg)
            sin_theta2 = sin_theta_max2 * u0;
            cos_theta = std::sqrt(1.0f - sin_theta2);
        }
        const Float cos_alpha = sin_theta2 * inv_sin_theta_max +
                                cos_theta * std::sqrt(fmax_(1.0f - sin_theta2 * inv_sin_theta_max * inv_sin_theta_max, 0.0f));
        const Float sin_alpha = std::sqrt(fmax_(1.0f - cos_alpha * cos_alpha, 0.0f));
        const Float phi = u1 * 2.0f * kPi;
        Float s, c;
        sincos_contract(phi, &s, &c);
        // geometry.rs:1156-1165 spherical_direction(sin_alpha, cos_alpha, phi, -wc_x, -wc_y, -wc)
        const V3 n_world = ((-wc_x) * sin_alpha * c + (-wc_y) * sin_alpha * s) + (-wc) * cos_alpha;
        const V3 p_world = p_center + n_world * radius;
        *p = p_world;
        *p_err = vabs(p_world) * gamma(5.0f);
        *n = reverse_orientation ? n_world * -1.0f : n_world;
        *pdf = 1.0f / (2.0f * kPi * (1.0f - cos_theta_max));
    }
    // sphere.rs:195-207 Sphere::pdf2 with shape.rs:54-69 as the inside branch
    Float pdf2(V3 rp, V3 rerr, V3 rn, V3 wi) const {
        const V3 p_center = xf_point(o2w, V3{0, 0, 0});
        const V3 p_origin = offset_ray_origin_(rp, rerr, rn, p_center - rp);
        if (length_squared(p_origin - p_center) < radius * radius) {
            const V3 o = offset_ray_origin_(rp, rerr, rn, wi);         // it.spawn_ray(wi), interaction.rs:132-135
            const Ray ray{o, kInfinity, wi, 0.0f};
            Float t;
            SphereSI li;
            if (!intersect(ray, &t, &li)) return 0.0f;
            Float pdf = length_squared(rp - li.p) / (std::fabs(dot(li.n, -wi)) * area());
            if (std::isinf(pdf)) pdf = 0.0f;
            return pdf;
        }
        const Float sin_theta_max2 = radius * radius / length_squared(rp - p_center);
        const Float cos_theta_max = std::sqrt(fmax_(1.0f - sin_theta_max2, 0.0f));
        return 1.0f / (2.0f * kPi * (1.0f - cos_theta_max));           // sampling.rs:248-250
    }
    // geometry.rs:1139-1154 (same text as oracle_core.hpp's offset_ray_origin, which is defined further down that file)
    static V3 offset_ray_origin_(V3 p, V3 p_error, V3 n, V3 w) {
        const Float d = dot(vabs(n), p_error);
        V3 offset = n * d;
        if (dot(w, n) < 0.0f) offset = -offset;
        V3 po = p + offset;
        for (int i = 0; i < 3; ++i) {
            if (offset[i] > 0.0f) po[i] = next_float_up(po[i]);
            else if (offset[i] < 0.0f) po[i] = next_float_down(po[i]);
        }
        return po;
    }
};
