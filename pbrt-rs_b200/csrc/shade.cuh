// shade.cuh — device-side scattering, light sampling and film math of the wavefront PathIntegrator (sm_100a).
//
// Replaces, per path vertex: BSDF::{f, pdf, sample_f} (src/core/reflection.rs:206-449) with the three lobes the
// materials of the configs need — LambertianReflection (:821-855 + BxDF defaults :459-507), MicrofacetReflection with
// TrowbridgeReitzDistribution (:977-1056, src/core/microfacet.rs:145-248,336-406) and FresnelSpecular (:733-819) —
// fr_dielectric (:19-40), the shading-space helpers (:71-156), power_heuristic (src/core/sampling.rs:306-313),
// Distribution1D::sample_discrete (:130-149), RGB<->XYZ / y_value (src/core/spectrum.rs:96-107,679-682) and the
// FilmTile::add_sample footprint (src/core/film.rs:252-295).  matte / plastic / glass follow pbrt-v3 (the reference's
// material files are empty).  Arithmetic order is the reference's, one rounded f32 op each
// (-fmad=false); sin/cos are det_sincos (pb2_math.cuh).  Appendix-A fixes applied: D6 D22 D23 D28 D30 D33-D36 D40.
#pragma once
#include "pb2_math.cuh"

namespace pb2 {

struct rgb3 {
    float r, g, b;
};
PB2_HD rgb3 mkc(float r, float g, float b) { rgb3 c; c.r = r; c.g = g; c.b = b; return c; }
PB2_HD rgb3 gray(float v) { return mkc(v, v, v); }
PB2_HD rgb3 operator+(rgb3 a, rgb3 b) { return mkc(a.r + b.r, a.g + b.g, a.b + b.b); }
PB2_HD rgb3 operator*(rgb3 a, rgb3 b) { return mkc(a.r * b.r, a.g * b.g, a.b * b.b); }
PB2_HD rgb3 operator*(rgb3 a, float s) { return mkc(a.r * s, a.g * s, a.b * s); }
PB2_HD rgb3 operator/(rgb3 a, float s) { return mkc(a.r / s, a.g / s, a.b / s); }
PB2_HD bool black(rgb3 a) { return a.r == 0.0f && a.g == 0.0f && a.b == 0.0f; }
PB2_HD float luminance(rgb3 a) { return (0.212671f * a.r + 0.715160f * a.g) + 0.072169f * a.b; }
PB2_HD float max_channel(rgb3 a) {          // spectrum.rs:161-165: fold from f32::MIN with `if max > v {max} else {v}`
    float m = -3.402823466e+38f;
    m = (m > a.r) ? m : a.r;
    m = (m > a.g) ? m : a.g;
    m = (m > a.b) ? m : a.b;
    return m;
}
PB2_HD bool any_nan(rgb3 a) { return isnan(a.r) || isnan(a.g) || isnan(a.b); }
PB2_HD void to_xyz(rgb3 c, float* x, float* y, float* z) {
    *x = (0.412453f * c.r + 0.357580f * c.g) + 0.180423f * c.b;
    *y = (0.212671f * c.r + 0.715160f * c.g) + 0.072169f * c.b;
    *z = (0.019334f * c.r + 0.119193f * c.g) + 0.950227f * c.b;
}
PB2_HD rgb3 from_xyz(float x, float y, float z) {
    return mkc((3.240479f * x - 1.537150f * y) - 0.498535f * z, (-0.969256f * x + 1.875991f * y) + 0.041556f * z,
               (0.055648f * x - 0.204043f * y) + 1.057311f * z);
}
PB2_HD float clamp01s(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

enum : unsigned { kReflection = 1u, kTransmission = 2u, kDiffuse = 4u, kGlossy = 8u, kSpecular = 16u, kAllLobes = 31u };

// ---- shading-space helpers (reflection.rs:71-156) --------------------------------------------------------------
PB2_HD float cos_t(vec3 w) { return w.z; }
PB2_HD float cos2_t(vec3 w) { return w.z * w.z; }
PB2_HD float abs_cos_t(vec3 w) { return fabsf(w.z); }
PB2_HD float sin2_t(vec3 w) { return fmaxf(1.0f - cos2_t(w), 0.0f); }
PB2_HD float sin_t(vec3 w) { return sqrtf(sin2_t(w)); }
PB2_HD float tan_t(vec3 w) { return sin_t(w) / cos_t(w); }
PB2_HD float tan2_t(vec3 w) { return sin2_t(w) / cos2_t(w); }
PB2_HD float cos_p(vec3 w) { const float s = sin_t(w); return s == 0.0f ? 1.0f : clamp01s(w.x / s, -1.0f, 1.0f); }
PB2_HD float sin_p(vec3 w) { const float s = sin_t(w); return s == 0.0f ? 0.0f : clamp01s(w.y / s, -1.0f, 1.0f); }
PB2_HD float cos2_p(vec3 w) { return cos_p(w) * cos_p(w); }
PB2_HD float sin2_p(vec3 w) { return sin_p(w) * sin_p(w); }
PB2_HD bool same_side(vec3 a, vec3 b) { return a.z * b.z > 0.0f; }
PB2_HD vec3 mirror(vec3 wo, vec3 n) { return -wo + n * (2.0f * dot3(wo, n)); }
PB2_HD vec3 face_toward(vec3 n, vec3 v) { return dot3(n, v) < 0.0f ? -n : n; }       // pbrt Faceforward(n, v)
PB2_HD bool refract_dir(vec3 wi, vec3 n, float eta, vec3* wt) {
    const float ci = dot3(n, wi);
    const float s2i = fmaxf(1.0f - ci * ci, 0.0f);
    const float s2t = eta * eta * s2i;
    if (s2t >= 1.0f) return false;
    const float ct = sqrtf(1.0f - s2t);
    *wt = -wi * eta + n * (eta * ci - ct);
    return true;
}
PB2_HD float fresnel_dielectric(float ci, float eta_i, float eta_t) {
    ci = clamp01s(ci, -1.0f, 1.0f);
    if (!(ci > 0.0f)) {
        const float t = eta_i; eta_i = eta_t; eta_t = t;
        ci = fabsf(ci);
    }
    const float si = sqrtf(fmaxf(1.0f - ci * ci, 0.0f));
    const float st = eta_i / eta_t * si;
    if (st >= 1.0f) return 1.0f;
    const float ct = sqrtf(fmaxf(1.0f - st * st, 0.0f));
    const float r_parl = (eta_t * ci - eta_i * ct) / (eta_t * ci + eta_i * ct);
    const float r_perp = (eta_i * ci - eta_t * ct) / (eta_i * ci + eta_t * ct);
    return (r_parl * r_parl + r_perp * r_perp) / 2.0f;
}

// ---- sampling.rs --------------------------------------------------------------------------------------------------
PB2_HD void concentric_disk(float u0, float u1, float* x, float* y) {      // concentric_sample_disk, :258-273
    const float ox = u0 * 2.0f - 1.0f, oy = u1 * 2.0f - 1.0f;
    float dx = 0.0f, dy = 0.0f;
    if (!(ox == 0.0f && oy == 0.0f)) {
        float r, theta;
        if (fabsf(ox) > fabsf(oy)) { r = ox; theta = (PB2_PI / 4.0f) * (oy / ox); }
        else { r = oy; theta = (PB2_PI / 2.0f) - (PB2_PI / 4.0f) * (ox / oy); }
        float s, c;
        det_sincos(theta, &s, &c);
        dx = c * r;
        dy = s * r;
    }
    *x = dx;
    *y = dy;
}
PB2_HD vec3 cosine_hemisphere(float u0, float u1) {          // :289-294 (D30 FIX)
    float dx, dy;
    concentric_disk(u0, u1, &dx, &dy);
    return mk(dx, dy, sqrtf(fmaxf(0.0f, (1.0f - dx * dx) - dy * dy)));
}
PB2_HD float power_heuristic(float f_pdf, float g_pdf) {
    const float f = 1.0f * f_pdf, g = 1.0f * g_pdf;
    return (f * f) / (f * f + g * g);
}

// ---- Trowbridge-Reitz, isotropic alpha, visible-area sampling (microfacet.rs) ------------------------------------
PB2_HD float tr_d(float a, vec3 wh) {
    const float t2 = tan2_t(wh);
    if (isinf(t2)) return 0.0f;
    const float cos4 = cos2_t(wh) * cos2_t(wh);
    const float e = (cos2_p(wh) / (a * a) + sin2_p(wh) / (a * a)) * t2;
    return 1.0f / (PB2_PI * a * a * cos4 * (1.0f + e) * (1.0f + e));
}
PB2_HD float tr_lambda(float a, vec3 w) {
    const float at = fabsf(tan_t(w));
    if (isinf(at)) return 0.0f;
    const float alpha = sqrtf(cos2_p(w) * a * a + sin2_p(w) * a * a);
    const float a2t2 = (alpha * at) * (alpha * at);
    return (-1.0f + sqrtf(1.0f + a2t2)) / 2.0f;
}
PB2_HD float tr_g1(float a, vec3 w) { return 1.0f / (1.0f + tr_lambda(a, w)); }
PB2_HD float tr_g(float a, vec3 wo, vec3 wi) { return 1.0f / ((1.0f + tr_lambda(a, wo)) + tr_lambda(a, wi)); }
PB2_HD float tr_pdf(float a, vec3 wo, vec3 wh) { return tr_d(a, wh) * tr_g1(a, wo) * fabsf(dot3(wo, wh)) / abs_cos_t(wo); }
PB2_HD void tr_sample11(float cos_theta, float u1, float u2, float* slope_x, float* slope_y) {
    if (cos_theta > 0.9999f) {
        const float r = sqrtf(u1 / (1.0f - u1));
        const float phi = 6.28318530718f * u2;
        float s, c;
        det_sincos(phi, &s, &c);
        *slope_x = r * c;
        *slope_y = r * s;
        return;
    }
    const float sin_theta = sqrtf(fmaxf(1.0f - cos_theta * cos_theta, 0.0f));
    const float tan_theta = sin_theta / cos_theta;
    float a = 1.0f / tan_theta;
    const float g1 = 2.0f / (1.0f + sqrtf(1.0f + 1.0f / (a * a)));
    a = 2.0f * u1 / g1 - 1.0f;
    float tmp = 1.0f / (a * a - 1.0f);
    if (tmp > 1e10f) tmp = 1e10f;
    const float b = tan_theta;
    const float d = sqrtf(fmaxf(b * b * tmp * tmp - (a * a - b * b), 0.0f));
    const float sx1 = b * tmp - d, sx2 = b * tmp + d;
    *slope_x = (a < 0.0f || sx2 > 1.0f / tan_theta) ? sx1 : sx2;
    float s;
    if (u2 > 0.5f) { s = 1.0f; u2 = 2.0f * (u2 - 0.5f); }
    else { s = -1.0f; u2 = 2.0f * (0.5f - u2); }
    const float z = (u2 * (u2 * (u2 * 0.27385f - 0.73369f) + 0.46341f)) /
                    (u2 * (u2 * (u2 * 0.093073f + 0.309420f) - 1.00000f) + 0.5979999f);
    *slope_y = s * z * sqrtf(1.0f + *slope_x * *slope_x);
}
PB2_HD vec3 tr_sample_wh(float a, vec3 wo, float u0, float u1) {
    const bool flip = wo.z < 0.0f;
    const vec3 wi = flip ? -wo : wo;
    const vec3 ws = unit(mk(a * wi.x, a * wi.y, wi.z));
    float sx, sy;
    tr_sample11(cos_t(ws), u0, u1, &sx, &sy);
    const float tmp = cos_p(ws) * sx - sin_p(ws) * sy;
    sy = sin_p(ws) * sx + cos_p(ws) * sy;
    sx = tmp;
    sx = a * sx;                         // D40 FIX
    sy = a * sy;
    const vec3 wh = unit(mk(-sx, -sy, 1.0f));
    return flip ? -wh : wh;
}

// ---- materials as device PODs -------------------------------------------------------------------------------------
struct DMaterial {
    int type;               // PB2_MAT_*
    int cls;                // shading class = material queue / k_shade instantiation: 0 Lambertian matte, 1 general
                            // (plastic, metal, Oren-Nayar matte), 2 specular (glass, mirror)
    float kd[3], ks[3], kr[3], kt[3];
    float alpha;            // plastic, metal: Trowbridge-Reitz alpha (roughness_to_alpha applied on the host when remapping)
    float eta;              // glass
    float on_a, on_b;       // matte with sigma: OrenNayar::new's A and B (reflection.rs:925-937)
    float metal_eta[3], metal_k[3];
};

struct DLight {
    int type;               // PB2_LIGHT_*
    float p[3];             // point
    float l[3];             // I or L_emit
    unsigned prim;          // area: emissive triangle
    int two_sided;
    float area;
    float p0[3], p1[3], p2[3];
    float axis[3];          // spot: row 2 of world_to_light (spot.rs:52-53); distant: w_light (distant.rs:31)
    float cos_total_width, cos_falloff_start;   // spot.rs:38-39
    float world_radius;     // distant.rs:73-77
    int has_n, has_uv;      // area: the mesh carries vertex normals / UVs
    float n0[3], n1[3], n2[3];   // vertex normals of the emissive triangle (Triangle::sample, triangle.rs:338-341)
    float uv[6];            // its UVs (Shape::pdf2 -> Triangle::intersect frame check)
    int sphere;             // area light over an analytic sphere: index into SceneView::spheres, else -1
    // what estimate_direct derives from the emissive triangle's vertices alone, computed once when the scene is uploaded
    // (api_path.cu, same functions): Triangle::sample's normal normalize(cross(p1 - p0, p2 - p0)) (triangle.rs:336), the
    // interaction normal normalize(cross(dp02, dp12)) of a hit on it (triangle.rs:234-236), 1 / area (shape.rs:43) and whether
    // Triangle::intersect accepts the triangle's frame at all (triangle.rs:193-215)
    float ns_sample[3], n_hit[3];
    float inv_area;
    int frame_ok;
};

// HomogeneousMedium (media/homogeneous.rs:12-29)
struct DMedium {
    float sigma_a[3], sigma_s[3], sigma_t[3];
    float g;
};

enum LobeKind : unsigned { kLambert = 0u, kMicrofacet = 1u, kFresnelSpecular = 2u, kOrenNayar = 3u, kSpecularReflection = 4u, kMicrofacetConductor = 5u,
                           kMicrofacetTransmission = 6u, kFresnelBlend = 7u };
struct Lobe {
    unsigned kind, type;
    rgb3 r, t;              // conductor: t = eta
    float alpha, eta_a, eta_b;   // Oren-Nayar: eta_a, eta_b = A, B
    rgb3 k;                 // conductor absorption
};

// fr_conductor (reflection.rs:42-69), one channel
PB2_HD float fresnel_conductor1(float ci, float eta_i, float eta_t, float k) {
    ci = clamp01s(ci, -1.0f, 1.0f);
    const float eta = eta_t / eta_i, eta_k = k / eta_i;
    const float cos2 = ci * ci, sin2 = 1.0f - cos2;
    const float eta2 = eta * eta, eta_k2 = eta_k * eta_k;
    const float t0 = (eta2 - eta_k2) - sin2;
    const float a2_plus_b2 = sqrtf(t0 * t0 + (eta2 * eta_k2) * 4.0f);
    const float t1 = a2_plus_b2 + cos2;
    const float a = sqrtf((a2_plus_b2 + t0) * 0.5f);
    const float t2 = a * (2.0f * ci);
    const float rs = (t1 - t2) / (t1 + t2);
    const float t3 = a2_plus_b2 * cos2 + sin2 * sin2;
    const float t4 = t2 * sin2;
    const float rp = (rs * (t3 - t4)) / (t3 + t4);
    return (rp + rs) * 0.5f;
}

PB2_HD bool lobe_matches(const Lobe& l, unsigned flags) { return (l.type & flags) == l.type; }       // D33 FIX

// Lobe kinds a shading class can hold (make_bsdf<CLS>): the branches of the other kinds drop out of that class's kernel.
// CLS 3 is not a queue of its own: it is class 1 compiled for scenes whose class-1 materials are all PlasticMaterial with
// non-black Kd and Ks — lobes[0] = Lambertian, lobes[1] = dielectric microfacet, both kinds compile-time constants.  The general
// class-1 kernel carries six lobe kinds through fourteen inlined lobe evaluations: 43,800 SASS instructions, 66 % of its stall
// samples waiting for the instruction cache (profiles/r02_tuning.md).
template <int CLS>
PB2_HD constexpr bool may_be(unsigned kind) {
    return CLS < 0 || (CLS == 0 && kind == kLambert) ||
           (CLS == 1 && (kind == kLambert || kind == kMicrofacet || kind == kOrenNayar || kind == kMicrofacetConductor || kind == kMicrofacetTransmission || kind == kFresnelBlend)) ||
           (CLS == 2 && (kind == kFresnelSpecular || kind == kSpecularReflection)) ||
           (CLS == 3 && (kind == kLambert || kind == kMicrofacet));
}
template <int CLS>
PB2_HD rgb3 lobe_f(const Lobe& l, vec3 wo, vec3 wi) {
    if (may_be<CLS>(kLambert) && l.kind == kLambert) return l.r * (1.0f / PB2_PI);
    if (may_be<CLS>(kOrenNayar) && l.kind == kOrenNayar) {                                  // reflection.rs:943-971 (pbrt-v3 semantics, D61)
        const float sti = sin_t(wi), sto = sin_t(wo);
        float max_cos = 0.0f;
        if (sti > 1e-4f && sto > 1e-4f) {
            const float d_cos = cos_p(wi) * cos_p(wo) + sin_p(wi) * sin_p(wo);
            max_cos = fmaxf(0.0f, d_cos);
        }
        float sin_alpha, tan_beta;
        if (abs_cos_t(wi) > abs_cos_t(wo)) { sin_alpha = sto; tan_beta = sti / abs_cos_t(wi); }
        else { sin_alpha = sti; tan_beta = sto / abs_cos_t(wo); }
        return l.r * (1.0f / PB2_PI) * (l.eta_a + ((l.eta_b * max_cos) * sin_alpha) * tan_beta);
    }
    if (may_be<CLS>(kMicrofacet) && (l.kind == kMicrofacet || (may_be<CLS>(kMicrofacetConductor) && l.kind == kMicrofacetConductor))) {
        const float co = abs_cos_t(wo), ci = abs_cos_t(wi);
        vec3 wh = wi + wo;
        if (ci == 0.0f || co == 0.0f) return gray(0.0f);
        if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return gray(0.0f);
        wh = unit(wh);
        const float c = dot3(wi, face_toward(wh, mk(0.0f, 0.0f, 1.0f)));                                            // D6 FIX
        rgb3 fr;
        if (may_be<CLS>(kMicrofacetConductor) && l.kind == kMicrofacetConductor) {                    // FresnelConductor::evaluate(|cos|), reflection.rs:583-587
            const float ac = fabsf(c);
            fr = mkc(fresnel_conductor1(ac, 1.0f, l.t.r, l.k.r), fresnel_conductor1(ac, 1.0f, l.t.g, l.k.g), fresnel_conductor1(ac, 1.0f, l.t.b, l.k.b));
        } else fr = gray(fresnel_dielectric(c, l.eta_a, l.eta_b));
        return l.r * tr_d(l.alpha, wh) * tr_g(l.alpha, wo, wi) * fr / (4.0f * ci * co);
    }
    if (may_be<CLS>(kFresnelBlend) && l.kind == kFresnelBlend) {                          // reflection.rs:1224-1240 (r = Rd, t = Rs)
        const float ai = 1.0f - 0.5f * abs_cos_t(wi), ao = 1.0f - 0.5f * abs_cos_t(wo);
        const rgb3 diffuse = gray(28.0f / (23.0f * PB2_PI)) * l.r * (gray(1.0f) + l.t * -1.0f) * (1.0f - (ai * ai) * (ai * ai) * ai) *
                             (1.0f - (ao * ao) * (ao * ao) * ao);
        vec3 wh = wi + wo;
        if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return gray(0.0f);
        wh = unit(wh);
        const float c = 1.0f - dot3(wi, wh);
        const rgb3 schlick = l.t + (gray(1.0f) + l.t * -1.0f) * ((c * c) * (c * c) * c);  // schlick_fresnel :1212-1215
        const rgb3 specular = schlick * (tr_d(l.alpha, wh) / ((4.0f * fabsf(dot3(wi, wh))) * fmaxf(abs_cos_t(wi), abs_cos_t(wo))));
        return diffuse + specular;
    }
    if (may_be<CLS>(kMicrofacetTransmission) && l.kind == kMicrofacetTransmission) {      // reflection.rs:1093-1136, TransportMode::Radiance
        if (same_side(wo, wi)) return gray(0.0f);
        const float cto = cos_t(wo), cti = cos_t(wi);
        if (cti == 0.0f || cto == 0.0f) return gray(0.0f);
        const float eta = cto > 0.0f ? l.eta_b / l.eta_a : l.eta_a / l.eta_b;
        vec3 wh = unit(wo + wi * eta);
        if (wh.z < 0.0f) wh = -wh;
        if (dot3(wo, wh) * dot3(wi, wh) > 0.0f) return gray(0.0f);
        const float fr = fresnel_dielectric(dot3(wo, wh), l.eta_a, l.eta_b);
        const float sqrt_denom = dot3(wo, wh) + eta * dot3(wi, wh);
        const float factor = 1.0f / eta;
        const float num = ((((((tr_d(l.alpha, wh) * tr_g(l.alpha, wo, wi)) * eta) * eta) * fabsf(dot3(wi, wh))) * fabsf(dot3(wo, wh))) * factor) * factor;
        const float den = ((cti * cto) * sqrt_denom) * sqrt_denom;
        return (gray(1.0f) + gray(fr) * -1.0f) * l.t * fabsf(num / den);
    }
    return gray(0.0f);
}
template <int CLS>
PB2_HD float lobe_pdf(const Lobe& l, vec3 wo, vec3 wi) {
    if (may_be<CLS>(kLambert) && (l.kind == kLambert || (may_be<CLS>(kOrenNayar) && l.kind == kOrenNayar))) return same_side(wo, wi) ? abs_cos_t(wi) * (1.0f / PB2_PI) : 0.0f;
    if (may_be<CLS>(kMicrofacet) && (l.kind == kMicrofacet || (may_be<CLS>(kMicrofacetConductor) && l.kind == kMicrofacetConductor))) {
        if (!same_side(wo, wi)) return 0.0f;
        const vec3 wh = unit(wo + wi);
        return tr_pdf(l.alpha, wo, wh) / (4.0f * dot3(wo, wh));
    }
    if (may_be<CLS>(kFresnelBlend) && l.kind == kFresnelBlend) {                          // reflection.rs:1267-1275
        if (!same_side(wo, wi)) return 0.0f;
        const vec3 wh = unit(wo + wi);
        const float pdf_wh = tr_pdf(l.alpha, wo, wh);
        return 0.5f * (abs_cos_t(wi) * (1.0f / PB2_PI) + pdf_wh / (4.0f * dot3(wo, wh)));
    }
    if (may_be<CLS>(kMicrofacetTransmission) && l.kind == kMicrofacetTransmission) {      // reflection.rs:1170-1187, D62 FIX
        if (same_side(wo, wi)) return 0.0f;
        const float eta = cos_t(wo) > 0.0f ? l.eta_b / l.eta_a : l.eta_a / l.eta_b;
        const vec3 wh = unit(wo + wi * eta);
        if (dot3(wo, wh) * dot3(wi, wh) > 0.0f) return 0.0f;
        const float sqrt_denom = dot3(wo, wh) + eta * dot3(wi, wh);
        const float dwh_dwi = fabsf(((eta * eta) * dot3(wi, wh)) / (sqrt_denom * sqrt_denom));
        return tr_pdf(l.alpha, wo, wh) * dwh_dwi;
    }
    return 0.0f;
}
template <int CLS>
PB2_HD rgb3 lobe_sample_f(const Lobe& l, vec3 wo, vec3* wi, float u0, float u1, float* pdf, unsigned* sampled) {
    if (may_be<CLS>(kSpecularReflection) && l.kind == kSpecularReflection) {   // reflection.rs:640-651 with FresnelNoOp (:606-611)
        *wi = mk(-wo.x, -wo.y, wo.z);
        *pdf = 1.0f;
        return l.r * gray(1.0f) / abs_cos_t(*wi);
    }
    if (may_be<CLS>(kLambert) && (l.kind == kLambert || (may_be<CLS>(kOrenNayar) && l.kind == kOrenNayar))) {
        *wi = cosine_hemisphere(u0, u1);
        if (wo.z < 0.0f) wi->z = wi->z * -1.0f;
        *pdf = lobe_pdf<CLS>(l, wo, *wi);
        return lobe_f<CLS>(l, wo, *wi);
    }
    if (may_be<CLS>(kMicrofacet) && (l.kind == kMicrofacet || (may_be<CLS>(kMicrofacetConductor) && l.kind == kMicrofacetConductor))) {
        if (wo.z == 0.0f) return gray(0.0f);
        const vec3 wh = tr_sample_wh(l.alpha, wo, u0, u1);
        if (dot3(wo, wh) < 0.0f) return gray(0.0f);
        *wi = mirror(wo, wh);                                   // D36 FIX
        if (!same_side(wo, *wi)) return gray(0.0f);
        *pdf = tr_pdf(l.alpha, wo, wh) / (4.0f * dot3(wo, wh));
        return lobe_f<CLS>(l, wo, *wi);
    }
    if (may_be<CLS>(kFresnelBlend) && l.kind == kFresnelBlend) {                          // reflection.rs:1242-1265
        if (u0 < 0.5f) {
            *wi = cosine_hemisphere(fminf(PB2_ONE_MINUS_EPS, 2.0f * u0), u1);
            if (wo.z < 0.0f) wi->z = wi->z * -1.0f;
        } else {
            const vec3 wh = tr_sample_wh(l.alpha, wo, fminf(PB2_ONE_MINUS_EPS, 2.0f * (u0 - 0.5f)), u1);
            *wi = mirror(wo, wh);
            if (!same_side(wo, *wi)) return gray(0.0f);
        }
        *pdf = lobe_pdf<CLS>(l, wo, *wi);
        return lobe_f<CLS>(l, wo, *wi);
    }
    if (may_be<CLS>(kMicrofacetTransmission) && l.kind == kMicrofacetTransmission) {      // reflection.rs:1138-1168
        if (wo.z == 0.0f) return gray(0.0f);
        const vec3 wh = tr_sample_wh(l.alpha, wo, u0, u1);
        if (dot3(wo, wh) < 0.0f) return gray(0.0f);
        const float eta = cos_t(wo) > 0.0f ? l.eta_a / l.eta_b : l.eta_b / l.eta_a;
        if (!refract_dir(wo, wh, eta, wi)) return gray(0.0f);
        *pdf = lobe_pdf<CLS>(l, wo, *wi);
        return lobe_f<CLS>(l, wo, *wi);
    }
    if (!may_be<CLS>(kFresnelSpecular)) return gray(0.0f);
    const float fr = fresnel_dielectric(cos_t(wo), l.eta_a, l.eta_b);
    if (u0 < fr) {
        *wi = mk(-wo.x, -wo.y, wo.z);
        *sampled = kSpecular | kReflection;
        *pdf = fr;
        return l.r * fr / abs_cos_t(*wi);
    }
    const bool entering = cos_t(wo) > 0.0f;
    const float eta_i = entering ? l.eta_a : l.eta_b, eta_t = entering ? l.eta_b : l.eta_a;
    if (!refract_dir(wo, face_toward(mk(0.0f, 0.0f, 1.0f), wo), eta_i / eta_t, wi)) return gray(0.0f);
    rgb3 ft = l.t * (1.0f - fr);
    ft = ft * ((eta_i * eta_i) / (eta_t * eta_t));             // TransportMode::Radiance
    *sampled = kSpecular | kTransmission;
    *pdf = 1.0f - fr;
    return ft / abs_cos_t(*wi);
}

// NL = the most lobes the material can have (1 for matte and glass, 2 for plastic): with NL = 1 every lobe index is the
// constant 0, the lobe kind set by make_bsdf<MAT> is a compile-time constant and the lobe array lives in registers.
template <int NL, int CLS = -1>
struct BsdfT {
    float eta;
    vec3 ns, ng, ss, ts;
    int n;
    Lobe lobes[NL];
};
using Bsdf = BsdfT<2, -1>;

template <int NL, int CLS>
PB2_HD int bsdf_count(const BsdfT<NL, CLS>& b, unsigned flags) {
    int c = 0;
#pragma unroll
    for (int i = 0; i < NL; ++i) c += (i < b.n && lobe_matches(b.lobes[i], flags)) ? 1 : 0;
    return c;
}
template <int NL, int CLS>
PB2_HD vec3 to_local(const BsdfT<NL, CLS>& b, vec3 v) { return mk(dot3(v, b.ss), dot3(v, b.ts), dot3(v, b.ns)); }
template <int NL, int CLS>
PB2_HD vec3 to_world(const BsdfT<NL, CLS>& b, vec3 v) {                  // D34 FIX
    return mk((b.ss.x * v.x + b.ts.x * v.y) + b.ns.x * v.z, (b.ss.y * v.x + b.ts.y * v.y) + b.ns.y * v.z,
              (b.ss.z * v.x + b.ts.z * v.y) + b.ns.z * v.z);
}
template <int NL, int CLS>
PB2_HD rgb3 bsdf_sum_f(const BsdfT<NL, CLS>& b, vec3 wo, vec3 wi, bool refl, unsigned flags) {
    rgb3 sum = gray(0.0f);
#pragma unroll
    for (int i = 0; i < NL; ++i) {
        if (i >= b.n) break;
        const Lobe& l = b.lobes[i];
        if (lobe_matches(l, flags) && ((refl && (l.type & kReflection)) || (!refl && (l.type & kTransmission)))) sum = sum + lobe_f<CLS>(l, wo, wi);
    }
    return sum;
}
template <int NL, int CLS>
PB2_HD rgb3 bsdf_f(const BsdfT<NL, CLS>& b, vec3 wo_w, vec3 wi_w, unsigned flags) {
    const vec3 wi = to_local(b, wi_w), wo = to_local(b, wo_w);
    if (wo.z == 0.0f) return gray(0.0f);
    const bool refl = dot3(wi_w, b.ng) * dot3(wo_w, b.ng) > 0.0f;
    return bsdf_sum_f(b, wo, wi, refl, flags);
}
template <int NL, int CLS>
PB2_HD float bsdf_pdf(const BsdfT<NL, CLS>& b, vec3 wo_w, vec3 wi_w, unsigned flags) {
    if (b.n == 0) return 0.0f;
    const vec3 wo = to_local(b, wo_w), wi = to_local(b, wi_w);
    if (wo.z == 0.0f) return 0.0f;
    float p = 0.0f;
    int matching = 0;
#pragma unroll
    for (int i = 0; i < NL; ++i)
        if (i < b.n && lobe_matches(b.lobes[i], flags)) { ++matching; p += lobe_pdf<CLS>(b.lobes[i], wo, wi); }
    return matching > 0 ? p / (float)matching : 0.0f;
}
// BSDF::sample_f (:286-381).  *pdf must be pre-set by the caller (it is left untouched on the early exits, as in the reference).
template <int NL, int CLS>
PB2_HD rgb3 bsdf_sample_f(const BsdfT<NL, CLS>& b, vec3 wo_w, vec3* wi_w, float u0, float u1, float* pdf, unsigned flags, unsigned* sampled) {
    const int matching = bsdf_count(b, flags);
    if (matching == 0) { *pdf = 0.0f; *sampled = 0u; return gray(0.0f); }
    int comp = (int)floorf(u0 * (float)matching);
    if (comp > matching - 1) comp = matching - 1;
    int count = comp, chosen = 0;
#pragma unroll
    for (int i = 0; i < NL; ++i)
        if (i < b.n && lobe_matches(b.lobes[i], flags)) {
            if (count == 0) { chosen = i; break; }
            --count;
        }
    // (a select, not lobes[chosen]: a run-time index would push the lobe array into local memory)
    const Lobe bx = (NL > 1 && chosen == 1) ? b.lobes[NL - 1] : b.lobes[0];
    const float u0r = fminf(PB2_ONE_MINUS_EPS, u0 * (float)matching - (float)comp);
    vec3 wi = mk(0.0f, 0.0f, 0.0f);
    const vec3 wo = to_local(b, wo_w);
    if (wo.z == 0.0f) return gray(0.0f);
    *pdf = 0.0f;
    *sampled = bx.type;
    rgb3 fv = lobe_sample_f<CLS>(bx, wo, &wi, u0r, u1, pdf, sampled);
    if (*pdf == 0.0f) { *sampled = 0u; return gray(0.0f); }
    *wi_w = to_world(b, wi);
    if (!(bx.type & kSpecular) && matching > 1) {
#pragma unroll
        for (int i = 0; i < NL; ++i)
            if (i < b.n && i != chosen && lobe_matches(b.lobes[i], flags)) *pdf += lobe_pdf<CLS>(b.lobes[i], wo, wi);
    }
    if (matching > 1) *pdf = *pdf / (float)matching;
    if (!(bx.type & kSpecular)) {
        const bool refl = dot3(*wi_w, b.ng) * dot3(wo_w, b.ng) > 0.0f;
        fv = bsdf_sum_f(b, wo, wi, refl, flags);
    }
    return fv;
}

// Material::compute_scattering_functions (pbrt-v3 materials over the reference's BxDF blocks): matte
// (Lambertian, or Oren-Nayar when sigma != 0), plastic, glass, mirror, metal.  CLS is the shading class of `m` (DMaterial::cls)
// known at compile time — the wavefront shades one class per launch: 0 = Lambertian matte (one lobe whose kind is a
// compile-time constant: the kernel that carries C2 / C4), 1 = general (plastic, metal, Oren-Nayar matte; up to two lobes),
// 2 = specular (glass, mirror; one lobe); CLS < 0 reads everything at run time.
template <int CLS = -1>
PB2_HD BsdfT<(CLS == 0 || CLS == 2) ? 1 : 2, CLS> make_bsdf(const DMaterial& m, vec3 ng, vec3 ns, vec3 ss, vec3 ts) {
    BsdfT<(CLS == 0 || CLS == 2) ? 1 : 2, CLS> b;
    const int type = CLS == 0 ? 0 : (CLS == 3 ? 1 : m.type);
    b.eta = type == 2 ? m.eta : 1.0f;
    b.ns = ns;
    b.ng = ng;
    b.ss = ss;                                                 // reflection.rs:220-234: ss = normalize(shading.dpdu), ts = cross(ns, ss)
    b.ts = ts;                                                 // (rebuild_vertex: per primitive for a plain mesh)
    b.n = 0;
    const rgb3 kd = mkc(m.kd[0], m.kd[1], m.kd[2]), ks = mkc(m.ks[0], m.ks[1], m.ks[2]);
    const rgb3 kr = mkc(m.kr[0], m.kr[1], m.kr[2]), kt = mkc(m.kt[0], m.kt[1], m.kt[2]);
    const rgb3 zero = gray(0.0f);
    // (the first lobe of every material is written to lobes[0] by constant index: a `lobes[b.n++]` whose index the compiler
    // cannot fold moves the whole lobe array to local memory)
    if (CLS == 3) {                                            // PlasticMaterial, Kd and Ks non-black (checked when the scene is uploaded)
        b.n = 2;
        Lobe& l0 = b.lobes[0];
        l0.kind = kLambert; l0.type = kReflection | kDiffuse; l0.r = kd; l0.t = zero; l0.alpha = 0.0f; l0.eta_a = 1.0f; l0.eta_b = 1.0f; l0.k = zero;
        Lobe& l1 = b.lobes[1];
        l1.kind = kMicrofacet; l1.type = kReflection | kGlossy; l1.r = ks; l1.t = zero; l1.alpha = m.alpha; l1.eta_a = 1.5f; l1.eta_b = 1.0f; l1.k = zero;
    } else if (type == 0 || type == 1) {
        if (!black(kd)) {
            Lobe& l = b.lobes[0];
            b.n = 1;
            const bool oren_nayar = CLS != 0 && type == 0;     // class 0 is Lambertian by construction
            l.kind = oren_nayar ? kOrenNayar : kLambert; l.type = kReflection | kDiffuse; l.r = kd; l.t = zero; l.alpha = 0.0f;
            l.eta_a = oren_nayar ? m.on_a : 1.0f; l.eta_b = oren_nayar ? m.on_b : 1.0f; l.k = zero;
        }
        if (type == 1 && !black(ks)) {
            Lobe& l = b.lobes[b.n++];
            l.kind = kMicrofacet; l.type = kReflection | kGlossy; l.r = ks; l.t = zero; l.alpha = m.alpha; l.eta_a = 1.5f; l.eta_b = 1.0f; l.k = zero;
        }
    } else if (type == 2 && (CLS == 1 || (CLS < 0 && m.cls == 1))) {
        // GlassMaterial with roughness: MicrofacetReflection(Kr, TrowbridgeReitz, FresnelDielectric(1, eta)) + MicrofacetTransmission(Kt, ..., 1, eta)
        if (!black(kr)) {
            Lobe& l = b.lobes[0];
            b.n = 1;
            l.kind = kMicrofacet; l.type = kReflection | kGlossy; l.r = kr; l.t = zero; l.alpha = m.alpha; l.eta_a = 1.0f; l.eta_b = m.eta; l.k = zero;
        }
        if (!black(kt)) {
            Lobe& l = b.lobes[b.n++];
            l.kind = kMicrofacetTransmission; l.type = kTransmission | kGlossy; l.r = zero; l.t = kt; l.alpha = m.alpha; l.eta_a = 1.0f; l.eta_b = m.eta; l.k = zero;
        }
    } else if (type == 5) {                                    // SubstrateMaterial: FresnelBlend(Kd, Ks, TrowbridgeReitz)
        if (!black(kd) || !black(ks)) {
            Lobe& l = b.lobes[0];
            b.n = 1;
            l.kind = kFresnelBlend; l.type = kReflection | kGlossy; l.r = kd; l.t = ks; l.alpha = m.alpha; l.eta_a = 1.0f; l.eta_b = 1.0f; l.k = zero;
        }
    } else if (type == 4) {                                    // MetalMaterial: MicrofacetReflection(1, TrowbridgeReitz, FresnelConductor(1, eta, k))
        Lobe& l = b.lobes[0];
        b.n = 1;
        l.kind = kMicrofacetConductor; l.type = kReflection | kGlossy; l.r = gray(1.0f); l.t = mkc(m.metal_eta[0], m.metal_eta[1], m.metal_eta[2]);
        l.alpha = m.alpha; l.eta_a = 1.0f; l.eta_b = 1.0f; l.k = mkc(m.metal_k[0], m.metal_k[1], m.metal_k[2]);
    } else {
        // type 3 = MirrorMaterial: SpecularReflection(Kr, FresnelNoOp); type 2 = GlassMaterial: FresnelSpecular(Kr, Kt, 1, eta).
        // One write site with selects, so the lobe stays in registers whichever of the two it is.
        const bool mir = type == 3;
        if (mir ? !black(kr) : !(black(kr) && black(kt))) {
            Lobe& l = b.lobes[0];
            b.n = 1;
            l.kind = mir ? kSpecularReflection : kFresnelSpecular;
            l.type = mir ? (kReflection | kSpecular) : (kReflection | kTransmission | kSpecular);
            l.r = kr; l.t = mir ? zero : kt; l.alpha = 0.0f; l.eta_a = 1.0f; l.eta_b = mir ? 1.0f : m.eta; l.k = zero;
        }
    }
    return b;
}

// ---- HenyeyGreenstein (medium.rs:34-87) behind the interface estimate_direct uses for a BSDF ---------------------------------
PB2_HD float phase_hg(float cos_theta, float g) {
    const float denom = 1.0f + g * g + 2.0f * g * cos_theta;
    return (1.0f / PB2_PI / 4.0f) * (1.0f - g * g) / (denom * sqrtf(denom));       // INV_4_PI = INV_PI / 4 (pbrt.rs:19-21)
}
struct PhaseHG {
    float g;
    vec3 ns;                // unused (a medium interaction has no normal)
};
PB2_HD float hg_sample_p(float g, vec3 wo, vec3* wi, float u0, float u1) {
    float cos_theta;
    if (fabsf(g) < 1e-3f) cos_theta = 1.0f - 2.0f * u0;
    else {
        const float sqr_term = (1.0f - g * g) / (1.0f + g - 2.0f * g * u0);
        cos_theta = -(1.0f + g * g - sqr_term * sqr_term) / (2.0f * g);
    }
    const float sin_theta = sqrtf(fmaxf(1.0f - cos_theta * cos_theta, 0.0f));
    const float phi = 2.0f * PB2_PI * u1;
    vec3 v1, v2;
    coord_system(wo, &v1, &v2);
    float sp, cp;
    det_sincos(phi, &sp, &cp);
    *wi = (v1 * sin_theta * cp + v2 * sin_theta * sp) + wo * cos_theta;              // geometry.rs:1156-1165
    return phase_hg(cos_theta, g);
}
// estimate_direct's medium-interaction branches (integrator.rs:165-170, 217-226): f = Spectrum(p), pdf = p, no cosine factor
PB2_HD rgb3 bsdf_f(const PhaseHG& ph, vec3 wo, vec3 wi, unsigned) { return gray(phase_hg(dot3(wo, wi), ph.g)); }
PB2_HD float bsdf_pdf(const PhaseHG& ph, vec3 wo, vec3 wi, unsigned) { return phase_hg(dot3(wo, wi), ph.g); }
PB2_HD rgb3 bsdf_sample_f(const PhaseHG& ph, vec3 wo, vec3* wi, float u0, float u1, float* pdf, unsigned, unsigned* sampled) {
    const float p = hg_sample_p(ph.g, wo, wi, u0, u1);
    *pdf = p;
    *sampled = 0u;
    return gray(p);
}
PB2_HD float cos_factor(const PhaseHG&, vec3) { return 1.0f; }                       // (x * 1.0f is exact)
template <int NL, int CLS>
PB2_HD float cos_factor(const BsdfT<NL, CLS>& b, vec3 wi) { return fabsf(dot3(wi, b.ns)); }   // wi.abs_dot(&isect.shading.n)

// Distribution1D::sample_discrete (sampling.rs:130-149) over cdf[0..n], func[0..n) — D57 KEEP (cdf < u), D58 signed clamp.
PB2_HD int sample_discrete(const float* cdf, const float* func, int n, float func_int, float u, float* pdf) {
    int first = 0, len = n + 1;
    while (len > 0) {
        const int half = len >> 1, middle = first + half;
        if (cdf[middle] < u) { first = middle + 1; len -= half + 1; }
        else len = half;
    }
    int off = first - 1;
    const int hi = n - 1;
    if (off < 0) off = 0; else if (off > hi) off = hi;
    *pdf = func_int > 0.0f ? func[off] / (func_int * (float)n) : 0.0f;
    return off;
}

}  // namespace pb2
