// ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.  PARITY UNPINNED (see oracle_core.hpp).
//
// CPU restatement of the PathIntegrator side of the hot path: PathIntegrator::li (src/integrators/path.rs:65-213),
// uniform_sample_one_light / estimate_direct (src/core/integrator.rs:92-266), SamplerIntegrator::render (:399-480),
// BSDF + Lambertian / MicrofacetReflection(Trowbridge-Reitz) / FresnelSpecular (src/core/reflection.rs,
// src/core/microfacet.rs), DiffuseAreaLight / PointLight (src/lights/), Distribution1D (src/core/sampling.rs),
// Film / FilmTile / Box+Gaussian filters (src/core/film.rs, src/filters/), RGB<->XYZ (src/core/spectrum.rs).
// Defect decisions follow SURVEY.md Appendix A; matte / plastic / glass are restated from pbrt-v3 (Appendix B) because
// src/materials/{matte,plastic,glass}.rs are empty files in the reference.
#pragma once
#include "oracle_core.hpp"

#include <atomic>
#include <chrono>
#include <memory>
#include <mutex>
#include <thread>
#include <unordered_map>

namespace orc {

// ---------------------------------------------------------------- spectrum.rs (RGBSpectrum)
struct RGB {
    Float r, g, b;
};
inline RGB rgb(Float v) { return {v, v, v}; }
inline RGB operator+(RGB a, RGB b) { return {a.r + b.r, a.g + b.g, a.b + b.b}; }
inline RGB operator*(RGB a, RGB b) { return {a.r * b.r, a.g * b.g, a.b * b.b}; }
inline RGB operator*(RGB a, Float s) { return {a.r * s, a.g * s, a.b * s}; }
inline RGB operator/(RGB a, Float s) { return {a.r / s, a.g / s, a.b / s}; }
inline bool is_black(RGB a) { return a.r == 0.0f && a.g == 0.0f && a.b == 0.0f; }        // spectrum.rs:176-183, D28 FIX
inline Float y_value(RGB a) { return (0.212671f * a.r + 0.715160f * a.g) + 0.072169f * a.b; }   // :679-682
inline Float max_component_value(RGB a) {                                                  // :161-165 (fold from f32::MIN)
    Float m = -std::numeric_limits<Float>::max();
    m = (m > a.r) ? m : a.r;
    m = (m > a.g) ? m : a.g;
    m = (m > a.b) ? m : a.b;
    return m;
}
inline bool has_nans(RGB a) { return std::isnan(a.r) || std::isnan(a.g) || std::isnan(a.b); }
inline void rgb_to_xyz(RGB c, Float xyz[3]) {                                              // :103-107
    xyz[0] = (0.412453f * c.r + 0.357580f * c.g) + 0.180423f * c.b;
    xyz[1] = (0.212671f * c.r + 0.715160f * c.g) + 0.072169f * c.b;
    xyz[2] = (0.019334f * c.r + 0.119193f * c.g) + 0.950227f * c.b;
}
inline void xyz_to_rgb(const Float xyz[3], Float out[3]) {                                 // :96-100
    out[0] = (3.240479f * xyz[0] - 1.537150f * xyz[1]) - 0.498535f * xyz[2];
    out[1] = (-0.969256f * xyz[0] + 1.875991f * xyz[1]) + 0.041556f * xyz[2];
    out[2] = (0.055648f * xyz[0] - 0.204043f * xyz[1]) + 1.057311f * xyz[2];
}
inline Float clampf(Float v, Float lo, Float hi) { return v < lo ? lo : (v > hi ? hi : v); }   // pbrt.rs:112-120

// ---------------------------------------------------------------- scene description (same POD layout as the C ABI)
enum { MAT_MATTE = 0, MAT_PLASTIC = 1, MAT_GLASS = 2, MAT_MIRROR = 3, MAT_METAL = 4, MAT_SUBSTRATE = 5 };
struct MaterialDesc {
    int32_t type;
    Float kd[3], ks[3];
    Float roughness;
    int32_t remap_roughness;
    Float kr[3], kt[3];
    Float eta;
    Float sigma;                     // matte: Oren-Nayar roughness in degrees (reflection.rs:917-937); 0 = Lambertian
    Float metal_eta[3], metal_k[3];  // metal: conductor index of refraction and absorption (reflection.rs:42-69)
};
enum { LIGHT_POINT = 0, LIGHT_AREA = 1, LIGHT_SPOT = 2, LIGHT_DISTANT = 3 };
struct LightDesc {
    int32_t type;
    Float p[3];
    Float i[3];
    uint32_t prim_id;
    int32_t two_sided;
    Float axis[3];          // spot: row 2 of world_to_light (spot.rs:52-53); distant: w (distant.rs:31)
    Float total_width;      // spot, degrees (spot.rs:38)
    Float falloff_start;    // spot, degrees (spot.rs:39)
};
struct CameraDesc {
    Float pos[3], look[3], up[3];
    Float fov;
    int32_t res_x, res_y;
    Float lens_radius, focal_distance;       // perspective.rs:26-27; lens_radius <= 0: pinhole
};
enum { FILTER_BOX = 0, FILTER_GAUSSIAN = 1, FILTER_TRIANGLE = 2, FILTER_MITCHELL = 3, FILTER_SINC = 4 };
struct FilmDesc {
    int32_t res_x, res_y;        // full_resolution (film.rs:19)
    int32_t filter;
    Float radius_x, radius_y;
    Float gaussian_alpha;
    Float mitchell_b, mitchell_c;
    Float sinc_tau;
    Float crop_window[4];        // {min.x, min.y, max.x, max.y} (film.rs:33,41-50); all zero = {0, 0, 1, 1}
    Float max_sample_luminance;  // film.rs:27; <= 0 = infinity
};
enum { LIGHTS_UNIFORM = 0, LIGHTS_POWER = 1, LIGHTS_SPATIAL = 2 };
struct PathDesc {
    int32_t max_depth;
    Float rr_threshold;
    int32_t light_strategy;
    int32_t spp;
    int32_t sample_begin, sample_end;
    int32_t sampler;          // 0 = RandomSampler (samplers/random.rs), 1 = HaltonSampler (samplers/halton.rs),
                              // 2 = StratifiedSampler (samplers/stratified.rs), 3 = ZeroTwoSequenceSampler (samplers/zerotwosequence.rs)
    int32_t n_sampled_dimensions;   // PixelSampler::new (sampler.rs:268-284): tabulated 1D and 2D dimensions (kinds 2, 3)
    int32_t x_samples, y_samples;   // StratifiedSampler::new (stratified.rs:23-39); spp == x_samples * y_samples
    int32_t jitter;
    int32_t integrator;       // 0 = PathIntegrator (integrators/path.rs), 1 = VolPathIntegrator (integrators/volpath.rs)
};

// src/media/homogeneous.rs HomogeneousMedium::new(sigma_a, sigma_s, g)
struct MediumDesc {
    Float sigma_a[3], sigma_s[3];
    Float g;
};
static constexpr uint32_t kNoMaterial = 0xFFFFFFFFu;   // GeometricPrimitive { material: None }: a surface that only separates media

// ---------------------------------------------------------------- sampling.rs:68-154 Distribution1D (D29 FIX, D57 KEEP, D58)
struct Distribution1D {
    std::vector<Float> func, cdf;
    Float func_int = 0;
    void init(const std::vector<Float>& f) {
        size_t n = f.size();
        func = f;
        cdf.assign(n + 1, 0.0f);
        for (size_t i = 1; i < n + 1; ++i) cdf[i] = cdf[i - 1] + func[i - 1] / (Float)n;
        func_int = cdf[n];
        if (func_int == 0.0f) for (size_t i = 1; i < n + 1; ++i) cdf[i] = (Float)i / (Float)n;
        else for (size_t i = 1; i < n + 1; ++i) cdf[i] /= func_int;
    }
    size_t count() const { return func.size(); }
    // pbrt.rs:229-243 find_interval with pred = cdf[i] < u (D57), signed clamp (D58)
    size_t sample_discrete(Float u, Float* pdf) const {
        long first = 0, len = (long)cdf.size();
        while (len > 0) {
            long half = len >> 1, middle = first + half;
            if (cdf[middle] < u) { first = middle + 1; len -= half + 1; }
            else len = half;
        }
        long off = first - 1;
        long hi = (long)cdf.size() - 2;
        if (off < 0) off = 0; else if (off > hi) off = hi;
        *pdf = func_int > 0.0f ? func[off] / (func_int * (Float)count()) : 0.0f;
        return (size_t)off;
    }
};

// lowdiscrepancy.rs:322-331 radical_inverse for the first five bases (2, 3, 5, 7, 11), as SpatialLightDistribution uses it
// (lightdistrib.rs:128-142); the digit loop follows pbrt-v3 where the port's is broken (S1 below).
inline Float radical_inverse_small(int base_index, uint64_t a) {
    static const uint64_t kBases[5] = {2, 3, 5, 7, 11};
    if (base_index == 0) {
        uint64_t r = 0;
        for (int i = 0; i < 64; ++i) r |= ((a >> i) & 1ull) << (63 - i);       // reverse_bits64, :364-369
        return (Float)r * 5.4210108624275222e-20f;
    }
    const uint64_t base = kBases[base_index];
    const Float inv_base = 1.0f / (Float)base;
    uint64_t reversed = 0;
    Float inv_base_n = 1.0f;
    while (a != 0) {
        const uint64_t next = a / base, digit = a - next * base;
        reversed = reversed * base + digit;
        inv_base_n *= inv_base;
        a = next;
    }
    return fmin_((Float)reversed * inv_base_n, kOneMinusEpsilon);
}

// ---------------------------------------------------------------- reflection.rs
enum : uint8_t { BSDF_REFLECTION = 1, BSDF_TRANSMISSION = 2, BSDF_DIFFUSE = 4, BSDF_GLOSSY = 8, BSDF_SPECULAR = 16, BSDF_ALL = 31 };

inline Float fr_dielectric(Float cos_theta_i, Float eta_i, Float eta_t) {                  // :19-40
    cos_theta_i = clampf(cos_theta_i, -1.0f, 1.0f);
    bool entering = cos_theta_i > 0.0f;
    if (!entering) { std::swap(eta_i, eta_t); cos_theta_i = std::fabs(cos_theta_i); }
    Float sin_theta_i = std::sqrt(fmax_(1.0f - cos_theta_i * cos_theta_i, 0.0f));
    Float sin_theta_t = eta_i / eta_t * sin_theta_i;
    if (sin_theta_t >= 1.0f) return 1.0f;
    Float cos_theta_t = std::sqrt(fmax_(1.0f - sin_theta_t * sin_theta_t, 0.0f));
    Float r_parl = (eta_t * cos_theta_i - eta_i * cos_theta_t) / (eta_t * cos_theta_i + eta_i * cos_theta_t);
    Float r_perp = (eta_i * cos_theta_i - eta_t * cos_theta_t) / (eta_i * cos_theta_i + eta_t * cos_theta_t);
    return (r_parl * r_parl + r_perp * r_perp) / 2.0f;
}
inline Float cos_theta(V3 w) { return w.z; }                                               // :71-135
inline Float cos2_theta(V3 w) { return w.z * w.z; }
inline Float abs_cos_theta(V3 w) { return std::fabs(w.z); }
inline Float sin2_theta(V3 w) { return fmax_(1.0f - cos2_theta(w), 0.0f); }
inline Float sin_theta(V3 w) { return std::sqrt(sin2_theta(w)); }
inline Float tan_theta(V3 w) { return sin_theta(w) / cos_theta(w); }
inline Float tan2_theta(V3 w) { return sin2_theta(w) / cos2_theta(w); }
inline Float cos_phi(V3 w) { Float s = sin_theta(w); return s == 0.0f ? 1.0f : clampf(w.x / s, -1.0f, 1.0f); }
inline Float sin_phi(V3 w) { Float s = sin_theta(w); return s == 0.0f ? 0.0f : clampf(w.y / s, -1.0f, 1.0f); }
inline Float cos2_phi(V3 w) { return cos_phi(w) * cos_phi(w); }
inline Float sin2_phi(V3 w) { return sin_phi(w) * sin_phi(w); }
inline V3 reflect(V3 wo, V3 n) { return -wo + n * (2.0f * dot(wo, n)); }                   // :140-142
inline bool refract(V3 wi, V3 n, Float eta, V3* wt) {                                      // :145-156, D35 FIX
    Float cos_theta_i = dot(n, wi);
    Float sin2_theta_i = fmax_(1.0f - cos_theta_i * cos_theta_i, 0.0f);
    Float sin2_theta_t = eta * eta * sin2_theta_i;
    if (sin2_theta_t >= 1.0f) return false;
    Float cos_theta_t = std::sqrt(1.0f - sin2_theta_t);
    *wt = -wi * eta + n * (eta * cos_theta_i - cos_theta_t);
    return true;
}
inline bool same_hemisphere(V3 w, V3 wp) { return w.z * wp.z > 0.0f; }
inline V3 faceforward(V3 n, V3 v) { return dot(n, v) < 0.0f ? -n : n; }                    // pbrt Faceforward(n, v) (D6 FIX)

// ---------------------------------------------------------------- microfacet.rs:145-248, 336-406 (Trowbridge-Reitz)
inline Float roughness_to_alpha(Float roughness) {                                         // :160-168
    roughness = fmax_(roughness, 1e-3f);
    Float x = std::log(roughness);
    return (((1.62142f + 0.819955f * x) + 0.1734f * x * x) + 0.0171201f * x * x * x) + 0.000640711f * x * x * x * x;
}
struct TrowbridgeReitz {
    Float ax, ay;
    Float d(V3 wh) const {                                                                  // :176-186
        Float t2 = tan2_theta(wh);
        if (std::isinf(t2)) return 0.0f;
        Float cos4 = cos2_theta(wh) * cos2_theta(wh);
        Float e = (cos2_phi(wh) / (ax * ax) + sin2_phi(wh) / (ay * ay)) * t2;
        return 1.0f / (kPi * ax * ay * cos4 * (1.0f + e) * (1.0f + e));
    }
    Float lambda(V3 w) const {                                                              // :188-199
        Float abs_tan = std::fabs(tan_theta(w));
        if (std::isinf(abs_tan)) return 0.0f;
        Float alpha = std::sqrt(cos2_phi(w) * ax * ax + sin2_phi(w) * ay * ay);
        Float a2t2 = (alpha * abs_tan) * (alpha * abs_tan);
        return (-1.0f + std::sqrt(1.0f + a2t2)) / 2.0f;
    }
    Float g1(V3 w) const { return 1.0f / (1.0f + lambda(w)); }                              // :15-17
    Float g(V3 wo, V3 wi) const { return 1.0f / ((1.0f + lambda(wo)) + lambda(wi)); }       // :18-20
    Float pdf(V3 wo, V3 wh) const { return d(wh) * g1(wo) * std::fabs(dot(wo, wh)) / abs_cos_theta(wo); }   // :23-29 (visible area)
    static void sample11(Float cos_t, Float u1, Float u2, Float* slope_x, Float* slope_y) {  // :336-384
        if (cos_t > 0.9999f) {
            Float r = std::sqrt(u1 / (1.0f - u1));
            Float phi = 6.28318530718f * u2;
            *slope_x = r * cos_c(phi);
            *slope_y = r * sin_c(phi);
            return;
        }
        Float sin_t = std::sqrt(fmax_(1.0f - cos_t * cos_t, 0.0f));
        Float tan_t = sin_t / cos_t;
        Float a = 1.0f / tan_t;
        Float g1 = 2.0f / (1.0f + std::sqrt(1.0f + 1.0f / (a * a)));
        a = 2.0f * u1 / g1 - 1.0f;
        Float tmp = 1.0f / (a * a - 1.0f);
        if (tmp > 1e10f) tmp = 1e10f;
        Float b = tan_t;
        Float dd = std::sqrt(fmax_(b * b * tmp * tmp - (a * a - b * b), 0.0f));
        Float sx1 = b * tmp - dd, sx2 = b * tmp + dd;
        *slope_x = (a < 0.0f || sx2 > 1.0f / tan_t) ? sx1 : sx2;
        Float s;
        if (u2 > 0.5f) { s = 1.0f; u2 = 2.0f * (u2 - 0.5f); }
        else { s = -1.0f; u2 = 2.0f * (0.5f - u2); }
        Float z = (u2 * (u2 * (u2 * 0.27385f - 0.73369f) + 0.46341f)) /
                  (u2 * (u2 * (u2 * 0.093073f + 0.309420f) - 1.00000f) + 0.5979999f);
        *slope_y = s * z * std::sqrt(1.0f + *slope_x * *slope_x);
    }
    static V3 sample(V3 wi, Float ax, Float ay, Float u1, Float u2) {                        // :386-406, D40 FIX
        V3 ws = normalize(V3{ax * wi.x, ay * wi.y, wi.z});
        Float sx, sy;
        sample11(cos_theta(ws), u1, u2, &sx, &sy);
        Float tmp = cos_phi(ws) * sx - sin_phi(ws) * sy;
        sy = sin_phi(ws) * sx + cos_phi(ws) * sy;
        sx = tmp;
        sx = ax * sx;
        sy = ay * sy;
        return normalize(V3{-sx, -sy, 1.0f});
    }
    V3 sample_wh(V3 wo, Float u0, Float u1) const {                                         // :201-243 (sample_visible_area)
        bool flip = wo.z < 0.0f;
        V3 wh = sample(flip ? -wo : wo, ax, ay, u0, u1);
        return flip ? -wh : wh;
    }
};

// fr_conductor (reflection.rs:42-69), one channel; eta_i = 1
inline Float fr_conductor1(Float cos_theta_i, Float eta_i, Float eta_t, Float k) {
    cos_theta_i = clampf(cos_theta_i, -1.0f, 1.0f);
    const Float eta = eta_t / eta_i, eta_k = k / eta_i;
    const Float cos2 = cos_theta_i * cos_theta_i, sin2 = 1.0f - cos2;
    const Float eta2 = eta * eta, eta_k2 = eta_k * eta_k;
    const Float t0 = (eta2 - eta_k2) - sin2;
    const Float a2_plus_b2 = std::sqrt(t0 * t0 + (eta2 * eta_k2) * 4.0f);
    const Float t1 = a2_plus_b2 + cos2;
    const Float a = std::sqrt((a2_plus_b2 + t0) * 0.5f);
    const Float t2 = a * (2.0f * cos_theta_i);
    const Float rs = (t1 - t2) / (t1 + t2);
    const Float t3 = a2_plus_b2 * cos2 + sin2 * sin2;
    const Float t4 = t2 * sin2;
    const Float rp = (rs * (t3 - t4)) / (t3 + t4);
    return (rp + rs) * 0.5f;
}

// ---------------------------------------------------------------- BxDFs as tagged PODs
// The Rust OrenNayar (reflection.rs:917-971) does not convert sigma to radians (`signma` is unused), takes sin_phi_o from
// sin_theta(wo) and builds d_cos from sin_theta instead of sin_phi; pbrt-v3 semantics here (D61 FIX).
// MicrofacetTransmission::pdf (reflection.rs:1170-1187) divides by sqrt_denom and then MULTIPLIES by it (`/ sqrt_denom * sqrt_denom`);
// pbrt-v3's dwh_dwi = |eta^2 (wi . wh) / sqrt_denom^2| is followed (D62 FIX).
enum LobeKind : uint8_t { LOBE_LAMBERT, LOBE_MICROFACET, LOBE_FRESNEL_SPECULAR, LOBE_OREN_NAYAR, LOBE_SPECULAR_REFLECTION, LOBE_MICROFACET_CONDUCTOR,
                          LOBE_MICROFACET_TRANSMISSION, LOBE_FRESNEL_BLEND };
struct Lobe {
    LobeKind kind;
    uint8_t type;       // BxDFType bits
    RGB r, t;           // reflectance / transmittance (conductor: t = eta)
    Float alpha;        // microfacet
    Float eta_a, eta_b; // fresnel specular / dielectric; Oren-Nayar: A, B
    RGB k{0, 0, 0};     // conductor absorption
    bool matches(uint8_t flags) const { return (type & flags) == type; }                    // :455-457, D33 FIX
    RGB f(V3 wo, V3 wi) const {
        switch (kind) {
            case LOBE_LAMBERT: return r * (1.0f / kPi);                                     // :840-842 (r * INV_PI)
            case LOBE_OREN_NAYAR: {                                                          // :943-971 (pbrt-v3)
                const Float sin_theta_i = sin_theta(wi), sin_theta_o = sin_theta(wo);
                Float max_cos = 0.0f;
                if (sin_theta_i > 1e-4f && sin_theta_o > 1e-4f) {
                    const Float sin_phi_i = sin_phi(wi), cos_phi_i = cos_phi(wi), sin_phi_o = sin_phi(wo), cos_phi_o = cos_phi(wo);
                    const Float d_cos = cos_phi_i * cos_phi_o + sin_phi_i * sin_phi_o;
                    max_cos = fmax_(0.0f, d_cos);
                }
                Float sin_alpha, tan_beta;
                if (abs_cos_theta(wi) > abs_cos_theta(wo)) { sin_alpha = sin_theta_o; tan_beta = sin_theta_i / abs_cos_theta(wi); }
                else { sin_alpha = sin_theta_i; tan_beta = sin_theta_o / abs_cos_theta(wo); }
                return r * (1.0f / kPi) * (eta_a + ((eta_b * max_cos) * sin_alpha) * tan_beta);
            }
            case LOBE_FRESNEL_BLEND: {                                                       // :1224-1240 (r = Rd, t = Rs)
                auto pow5 = [](Float v) { return (v * v) * (v * v) * v; };
                const RGB diffuse = rgb(28.0f / (23.0f * kPi)) * r * (rgb(1.0f) + t * -1.0f) * (1.0f - pow5(1.0f - 0.5f * abs_cos_theta(wi))) *
                                    (1.0f - pow5(1.0f - 0.5f * abs_cos_theta(wo)));
                V3 wh = wi + wo;
                if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return rgb(0);
                wh = normalize(wh);
                TrowbridgeReitz tr{alpha, alpha};
                const RGB schlick = t + (rgb(1.0f) + t * -1.0f) * pow5(1.0f - dot(wi, wh));    // schlick_fresnel :1212-1215
                const RGB specular = schlick * (tr.d(wh) / ((4.0f * std::fabs(dot(wi, wh))) * fmax_(abs_cos_theta(wi), abs_cos_theta(wo))));
                return diffuse + specular;
            }
            case LOBE_MICROFACET_TRANSMISSION: {                                             // :1093-1136 (TransportMode::Radiance)
                if (same_hemisphere(wo, wi)) return rgb(0);
                const Float cos_theta_o = cos_theta(wo), cos_theta_i = cos_theta(wi);
                if (cos_theta_i == 0.0f || cos_theta_o == 0.0f) return rgb(0);
                const Float eta = cos_theta(wo) > 0.0f ? eta_b / eta_a : eta_a / eta_b;
                V3 wh = normalize(wo + wi * eta);
                if (wh.z < 0.0f) wh = -wh;
                if (dot(wo, wh) * dot(wi, wh) > 0.0f) return rgb(0);
                TrowbridgeReitz tr{alpha, alpha};
                const Float fr = fr_dielectric(dot(wo, wh), eta_a, eta_b);
                const Float sqrt_denom = dot(wo, wh) + eta * dot(wi, wh);
                const Float factor = 1.0f / eta;
                const Float num = ((((((tr.d(wh) * tr.g(wo, wi)) * eta) * eta) * std::fabs(dot(wi, wh))) * std::fabs(dot(wo, wh))) * factor) * factor;
                const Float den = ((cos_theta_i * cos_theta_o) * sqrt_denom) * sqrt_denom;
                return (rgb(1.0f) + rgb(fr) * -1.0f) * t * std::fabs(num / den);
            }
            case LOBE_MICROFACET_CONDUCTOR: {                                                // :998-1017 with FresnelConductor (:583-587)
                Float co = abs_cos_theta(wo), ci = abs_cos_theta(wi);
                V3 wh = wi + wo;
                if (ci == 0.0f || co == 0.0f) return rgb(0);
                if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return rgb(0);
                wh = normalize(wh);
                TrowbridgeReitz tr{alpha, alpha};
                const Float c = std::fabs(dot(wi, faceforward(wh, V3{0, 0, 1})));
                const RGB fr{fr_conductor1(c, 1.0f, t.r, k.r), fr_conductor1(c, 1.0f, t.g, k.g), fr_conductor1(c, 1.0f, t.b, k.b)};
                return r * tr.d(wh) * tr.g(wo, wi) * fr / (4.0f * ci * co);
            }
            case LOBE_MICROFACET: {                                                          // :998-1017
                Float co = abs_cos_theta(wo), ci = abs_cos_theta(wi);
                V3 wh = wi + wo;
                if (ci == 0.0f || co == 0.0f) return rgb(0);
                if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return rgb(0);
                wh = normalize(wh);
                TrowbridgeReitz tr{alpha, alpha};
                Float fr = fr_dielectric(dot(wi, faceforward(wh, V3{0, 0, 1})), eta_a, eta_b);   // D6 FIX
                return r * tr.d(wh) * tr.g(wo, wi) * rgb(fr) / (4.0f * ci * co);
            }
            default: return rgb(0);                                                          // FresnelSpecular :761-763
        }
    }
    Float pdf(V3 wo, V3 wi) const {
        switch (kind) {
            case LOBE_LAMBERT:
            case LOBE_OREN_NAYAR: return same_hemisphere(wo, wi) ? abs_cos_theta(wi) * (1.0f / kPi) : 0.0f;   // :501-507
            case LOBE_FRESNEL_BLEND: {                                                       // :1267-1275
                if (!same_hemisphere(wo, wi)) return 0.0f;
                const V3 wh = normalize(wo + wi);
                TrowbridgeReitz tr{alpha, alpha};
                const Float pdf_wh = tr.pdf(wo, wh);
                return 0.5f * (abs_cos_theta(wi) * (1.0f / kPi) + pdf_wh / (4.0f * dot(wo, wh)));
            }
            case LOBE_MICROFACET_TRANSMISSION: {                                             // :1170-1187, D62 FIX
                if (same_hemisphere(wo, wi)) return 0.0f;
                const Float eta = cos_theta(wo) > 0.0f ? eta_b / eta_a : eta_a / eta_b;
                const V3 wh = normalize(wo + wi * eta);
                if (dot(wo, wh) * dot(wi, wh) > 0.0f) return 0.0f;
                const Float sqrt_denom = dot(wo, wh) + eta * dot(wi, wh);
                const Float dwh_dwi = std::fabs(((eta * eta) * dot(wi, wh)) / (sqrt_denom * sqrt_denom));
                TrowbridgeReitz tr{alpha, alpha};
                return tr.pdf(wo, wh) * dwh_dwi;
            }
            case LOBE_MICROFACET_CONDUCTOR:
            case LOBE_MICROFACET: {                                                          // :1042-1048
                if (!same_hemisphere(wo, wi)) return 0.0f;
                V3 wh = normalize(wo + wi);
                TrowbridgeReitz tr{alpha, alpha};
                return tr.pdf(wo, wh) / (4.0f * dot(wo, wh));
            }
            default: return 0.0f;
        }
    }
    RGB sample_f(V3 wo, V3* wi, Float u0, Float u1, Float* pdf_out, uint8_t* sampled_type) const {
        switch (kind) {
            case LOBE_SPECULAR_REFLECTION: {                                                 // :640-651, FresnelNoOp (:606-611)
                *wi = V3{-wo.x, -wo.y, wo.z};
                *pdf_out = 1.0f;
                return r * rgb(1.0f) / abs_cos_theta(*wi);
            }
            case LOBE_LAMBERT:
            case LOBE_OREN_NAYAR: {                                                          // :459-472 BxDF default
                *wi = cosine_sample_hemisphere(u0, u1);
                if (wo.z < 0.0f) wi->z *= -1.0f;
                *pdf_out = pdf(wo, *wi);
                return f(wo, *wi);
            }
            case LOBE_FRESNEL_BLEND: {                                                       // :1242-1265
                if (u0 < 0.5f) {
                    *wi = cosine_sample_hemisphere(fmin_(kOneMinusEpsilon, 2.0f * u0), u1);
                    if (wo.z < 0.0f) wi->z *= -1.0f;
                } else {
                    TrowbridgeReitz tr{alpha, alpha};
                    const V3 wh = tr.sample_wh(wo, fmin_(kOneMinusEpsilon, 2.0f * (u0 - 0.5f)), u1);
                    *wi = reflect(wo, wh);
                    if (!same_hemisphere(wo, *wi)) return rgb(0);
                }
                *pdf_out = pdf(wo, *wi);
                return f(wo, *wi);
            }
            case LOBE_MICROFACET_TRANSMISSION: {                                             // :1138-1168
                if (wo.z == 0.0f) return rgb(0);
                TrowbridgeReitz tr{alpha, alpha};
                const V3 wh = tr.sample_wh(wo, u0, u1);
                if (dot(wo, wh) < 0.0f) return rgb(0);
                const Float eta = cos_theta(wo) > 0.0f ? eta_a / eta_b : eta_b / eta_a;
                if (!refract(wo, wh, eta, wi)) return rgb(0);
                *pdf_out = pdf(wo, *wi);
                return f(wo, *wi);
            }
            case LOBE_MICROFACET_CONDUCTOR:
            case LOBE_MICROFACET: {                                                          // :1019-1040, D36 FIX
                if (wo.z == 0.0f) return rgb(0);
                TrowbridgeReitz tr{alpha, alpha};
                V3 wh = tr.sample_wh(wo, u0, u1);
                if (dot(wo, wh) < 0.0f) return rgb(0);
                *wi = reflect(wo, wh);
                if (!same_hemisphere(wo, *wi)) return rgb(0);
                *pdf_out = tr.pdf(wo, wh) / (4.0f * dot(wo, wh));
                return f(wo, *wi);
            }
            default: {                                                                       // FresnelSpecular :765-811
                Float fr = fr_dielectric(cos_theta(wo), eta_a, eta_b);
                if (u0 < fr) {
                    *wi = V3{-wo.x, -wo.y, wo.z};
                    *sampled_type = BSDF_SPECULAR | BSDF_REFLECTION;
                    *pdf_out = fr;
                    return r * fr / abs_cos_theta(*wi);
                }
                bool entering = cos_theta(wo) > 0.0f;
                Float eta_i = entering ? eta_a : eta_b, eta_t = entering ? eta_b : eta_a;
                if (!refract(wo, faceforward(V3{0, 0, 1}, wo), eta_i / eta_t, wi)) return rgb(0);   // D6 FIX
                RGB ft = t * (1.0f - fr);
                ft = ft * ((eta_i * eta_i) / (eta_t * eta_t));                                // TransportMode::Radiance
                *sampled_type = BSDF_SPECULAR | BSDF_TRANSMISSION;
                *pdf_out = 1.0f - fr;
                return ft / abs_cos_theta(*wi);
            }
        }
    }
};

struct BSDF {                                                                                // :206-449
    Float eta = 1.0f;
    V3 ns, ng, ss, ts;
    int n = 0;
    Lobe lobes[2];
    int num_components(uint8_t flags) const { int c = 0; for (int i = 0; i < n; ++i) c += lobes[i].matches(flags); return c; }
    V3 world_to_local(V3 v) const { return {dot(v, ss), dot(v, ts), dot(v, ns)}; }
    V3 local_to_world(V3 v) const {                                                          // :256-262, D34 FIX
        return {(ss.x * v.x + ts.x * v.y) + ns.x * v.z, (ss.y * v.x + ts.y * v.y) + ns.y * v.z, (ss.z * v.x + ts.z * v.y) + ns.z * v.z};
    }
    RGB f(V3 wo_w, V3 wi_w, uint8_t flags) const {                                           // :265-284
        V3 wi = world_to_local(wi_w), wo = world_to_local(wo_w);
        if (wo.z == 0.0f) return rgb(0);
        bool refl = dot(wi_w, ng) * dot(wo_w, ng) > 0.0f;
        RGB sum = rgb(0);
        for (int i = 0; i < n; ++i)
            if (lobes[i].matches(flags) && ((refl && (lobes[i].type & BSDF_REFLECTION)) || (!refl && (lobes[i].type & BSDF_TRANSMISSION))))
                sum = sum + lobes[i].f(wo, wi);
        return sum;
    }
    Float pdf(V3 wo_w, V3 wi_w, uint8_t flags) const {                                       // :420-448
        if (n == 0) return 0.0f;
        V3 wo = world_to_local(wo_w), wi = world_to_local(wi_w);
        if (wo.z == 0.0f) return 0.0f;
        Float p = 0.0f;
        int matching = 0;
        for (int i = 0; i < n; ++i)
            if (lobes[i].matches(flags)) { ++matching; p += lobes[i].pdf(wo, wi); }
        return matching > 0 ? p / (Float)matching : 0.0f;
    }
    RGB sample_f(V3 wo_w, V3* wi_w, Float u0, Float u1, Float* pdf_out, uint8_t flags, uint8_t* sampled) const {   // :286-381
        int matching = num_components(flags);
        if (matching == 0) { *pdf_out = 0.0f; *sampled = 0; return rgb(0); }
        int comp = std::min((int)std::floor(u0 * (Float)matching), matching - 1);
        int count = comp, chosen = -1;
        for (int i = 0; i < n; ++i)
            if (lobes[i].matches(flags) && count-- == 0) { chosen = i; break; }
        const Lobe& bx = lobes[chosen];
        Float u0r = fmin_(kOneMinusEpsilon, u0 * (Float)matching - (Float)comp);
        V3 wi{0, 0, 0};
        V3 wo = world_to_local(wo_w);
        if (wo.z == 0.0f) return rgb(0);
        *pdf_out = 0.0f;
        *sampled = bx.type;
        RGB fv = bx.sample_f(wo, &wi, u0r, u1, pdf_out, sampled);
        if (*pdf_out == 0.0f) { *sampled = 0; return rgb(0); }
        *wi_w = local_to_world(wi);
        if (!(bx.type & BSDF_SPECULAR) && matching > 1)
            for (int i = 0; i < n; ++i)
                if (i != chosen && lobes[i].matches(flags)) *pdf_out += lobes[i].pdf(wo, wi);
        if (matching > 1) *pdf_out /= (Float)matching;
        if (!(bx.type & BSDF_SPECULAR)) {
            bool refl = dot(*wi_w, ng) * dot(wo_w, ng) > 0.0f;
            fv = rgb(0);
            for (int i = 0; i < n; ++i)
                if (lobes[i].matches(flags) && ((refl && (lobes[i].type & BSDF_REFLECTION)) || (!refl && (lobes[i].type & BSDF_TRANSMISSION))))
                    fv = fv + lobes[i].f(wo, wi);
        }
        return fv;
    }
};

inline Float power_heuristic(Float f_pdf, Float g_pdf) {                                     // sampling.rs:306-313 (nf = ng = 1)
    Float f = 1.0f * f_pdf, g = 1.0f * g_pdf;
    return (f * f) / (f * f + g * g);
}

// Instrumentation for the path-tracing roofline (B(sample) = sum of B(r) over the rays of the path + path state per
// vertex + 16 B film): rays, boxes slab-tested and triangles tested in reference order per ray kind (0 = path / camera rays,
// 1 = shadow rays of estimate_direct, 2 = its BSDF-sampled MIS rays) and the number of path vertices shaded.  Counting is on
// only while a thread points tl_path_counters at its own block (orc_render_counted); results never depend on it.
struct PathCounters {
    uint64_t camera_samples = 0, vertices = 0;
    uint64_t rays[3] = {0, 0, 0};
    TraversalCounters trav[3];
};
inline thread_local PathCounters* tl_path_counters = nullptr;
inline thread_local int tl_ray_kind = 0;

// ---------------------------------------------------------------- interaction.rs SurfaceInteraction (subset)
struct SurfaceInteraction {
    V3 p, error, n, wo, dpdu;
    V3 sn, sdpdu;         // shading.n, shading.dpdu (interaction.rs:305-316; D59: the geometric values without mesh normals / tangents)
    uint32_t prim;
};

struct LightRt {
    LightDesc d;
    V3 p0, p1, p2;      // area: the emissive triangle
    bool has_n = false, has_uv = false;
    V3 n0, n1, n2;      // its vertex normals (TriangleMesh::n), when the mesh has them
    Float uv[6];        // its UVs (TriangleMesh::uv)
    Float area;
    const Sphere* sphere = nullptr;               // area light on an analytic sphere (DiffuseAreaLight over shapes/sphere.rs)
    Float cos_total_width, cos_falloff_start;     // spot.rs:38-39
    V3 w_light;                                   // distant.rs:31
    Float world_radius;                           // distant.rs:73-77 pre_process
    RGB l() const { return {d.i[0], d.i[1], d.i[2]}; }
    bool is_delta() const { return d.type != LIGHT_AREA; }                                   // light.rs:28-31, D24 FIX
    Float falloff(V3 w) const {                                                              // spot.rs:51-63
        const Float cos_theta = (d.axis[0] * w.x + d.axis[1] * w.y) + d.axis[2] * w.z;       // (world_to_light * w).z, transform.rs:374-386
        if (cos_theta < cos_total_width) return 0.0f;
        if (cos_theta >= cos_falloff_start) return 1.0f;
        const Float delta = (cos_theta - cos_total_width) / (cos_falloff_start - cos_total_width);
        return (delta * delta) * (delta * delta);
    }
};

// Light::sample_li without its VisibilityTester (point.rs:47-66, spot.rs:71-85, distant.rs:50-67, diffuse.rs:60-81 +
// shape.rs:38-53 + triangle.rs:330-348): incident radiance and pdf at a bare point `p` — what SpatialLightDistribution
// integrates per voxel (lightdistrib.rs:143-157).  Same operations, in the same order, as the head of estimate_direct.
inline RGB light_sample_li(const LightRt& light, V3 p, Float ul0, Float ul1, Float* pdf_out) {
    if (light.is_delta()) {
        *pdf_out = 1.0f;
        if (light.d.type == LIGHT_DISTANT) return light.l();
        const V3 pl{light.d.p[0], light.d.p[1], light.d.p[2]};
        const V3 wi = normalize(pl - p);
        if (light.d.type == LIGHT_SPOT) return light.l() * light.falloff(-wi) / length_squared(pl - p);
        return light.l() / length_squared(pl - p);
    }
    if (light.sphere) {                                                                        // diffuse.rs:60-81 + sphere.rs:127-193, bare reference point
        V3 ps, pe, ns;
        Float pdf;
        light.sphere->sample2(p, V3{0, 0, 0}, V3{0, 0, 0}, ul0, ul1, &ps, &pe, &ns, &pdf);
        *pdf_out = pdf;
        if (pdf == 0.0f || length_squared(ps - p) == 0.0f) { *pdf_out = 0.0f; return rgb(0); }
        const V3 wi = normalize(ps - p);
        return (light.d.two_sided || dot(ns, -wi) > 0.0f) ? light.l() : rgb(0);
    }
    Float su0 = std::sqrt(ul0);
    Float b0 = 1.0f - su0, b1 = ul1 * su0;
    V3 ps = (light.p0 * b0 + light.p1 * b1) + light.p2 * ((1.0f - b0) - b1);
    V3 ns = normalize(cross(light.p1 - light.p0, light.p2 - light.p0));
    if (light.has_n) ns = faceforward(ns, (light.n0 * b0 + light.n1 * b1) + light.n2 * ((1.0f - b0) - b1));
    Float pdf = 1.0f / light.area;
    V3 w = ps - p;
    if (length_squared(w) == 0.0f) pdf = 0.0f;
    else {
        w = normalize(w);
        pdf *= length_squared(p - ps) / std::fabs(dot(ns, -w));
        if (std::isinf(pdf)) pdf = 0.0f;
    }
    *pdf_out = pdf;
    if (pdf == 0.0f || length_squared(ps - p) == 0.0f) { *pdf_out = 0.0f; return rgb(0); }
    const V3 wi = normalize(ps - p);
    return (light.d.two_sided || dot(ns, -wi) > 0.0f) ? light.l() : rgb(0);                   // D55 FIX
}

struct MaterialRt {
    MaterialDesc d;
    Float alpha;        // plastic, metal: roughness (remapped) — host-side, microfacet.rs:160-168
    Float on_a, on_b;   // matte with sigma: OrenNayar::new's A and B (reflection.rs:925-937, sigma in radians: D61 FIX)
};

class Scene {
public:
    BVHAccel bvh;
    std::vector<uint32_t> tri_material;
    std::vector<MaterialRt> materials;
    std::vector<LightRt> lights;
    std::vector<int32_t> tri_light;
    Distribution1D light_distrib;
    std::vector<V3> vn, vs;                  // TriangleMesh::n / ::s (triangle.rs:19-20), one per vertex, empty = absent
    // participating media (media/homogeneous.rs) and each primitive's MediumInterface {inside, outside} as indices (-1 = none);
    // camera_medium = the medium camera rays start in (perspective.rs:109)
    std::vector<MediumDesc> media;
    std::vector<int32_t> prim_inside, prim_outside;
    int32_t camera_medium = -1;
    void set_media(const MediumDesc* m, uint32_t n, const int32_t* inside, const int32_t* outside, int32_t cam_medium) {
        media.assign(m, m + n);
        const size_t np = tri_material.size();
        prim_inside.assign(np, -1);
        prim_outside.assign(np, -1);
        if (inside) prim_inside.assign(inside, inside + np);
        if (outside) prim_outside.assign(outside, outside + np);
        camera_medium = cam_medium;
    }

    // TriangleMesh's optional per-vertex normals, tangents and UVs (world space, as given); call after init()
    void set_shading_geometry(const Float* normals, const Float* tangents, const Float* uv) {
        const size_t nv = bvh.verts.size();
        vn.clear(); vs.clear(); bvh.uvs.clear();
        if (normals) { vn.resize(nv); for (size_t i = 0; i < nv; ++i) vn[i] = {normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]}; }
        if (tangents) { vs.resize(nv); for (size_t i = 0; i < nv; ++i) vs[i] = {tangents[3 * i], tangents[3 * i + 1], tangents[3 * i + 2]}; }
        if (uv) bvh.uvs.assign(uv, uv + 2 * nv);
        for (LightRt& l : lights) {
            if (l.d.type != LIGHT_AREA || l.sphere) continue;
            const uint32_t* ix = &bvh.indices[3 * (size_t)l.d.prim_id];
            l.has_n = !vn.empty();
            if (l.has_n) { l.n0 = vn[ix[0]]; l.n1 = vn[ix[1]]; l.n2 = vn[ix[2]]; }
            l.has_uv = bvh.tri_uv(l.d.prim_id, l.uv) != nullptr;
        }
    }

    // spheres: analytic Sphere primitives appended to the primitive list (ids nt .. nt + n_spheres - 1); an area light whose prim_id
    // is such an id is a DiffuseAreaLight over that sphere
    void init(const Float* verts, uint64_t nv, const uint32_t* idx, uint64_t nt, const uint32_t* tri_mat, const MaterialDesc* mats,
              uint32_t n_mats, const LightDesc* lts, uint32_t n_lights, int max_prims, const SphereDesc* sph = nullptr, uint32_t n_spheres = 0) {
        bvh.spheres.resize(n_spheres);
        for (uint32_t i = 0; i < n_spheres; ++i) {
            M4 o2w;
            std::memcpy(o2w.m, sph[i].object_to_world, sizeof(o2w.m));
            const M4 inv = m4_inverse(o2w);                                                    // Transform::new (transform.rs:198-206)
            Mat4 w2o;
            std::memcpy(w2o.m, inv.m, sizeof(w2o.m));
            bvh.spheres[i].init(sph[i], w2o);
        }
        bvh.build(verts, nv, idx, nt, max_prims);
        tri_material.assign(tri_mat, tri_mat + nt);
        for (uint32_t i = 0; i < n_spheres; ++i) tri_material.push_back(sph[i].material);
        const uint64_t n_prims = nt + n_spheres;
        materials.resize(n_mats);
        prim_inside.assign(n_prims, -1);
        prim_outside.assign(n_prims, -1);
        for (uint32_t i = 0; i < n_mats; ++i) {
            materials[i].d = mats[i];
            materials[i].alpha = mats[i].remap_roughness ? roughness_to_alpha(mats[i].roughness) : mats[i].roughness;
            {
                const Float sig = kPi / 180.0f * clampf(mats[i].sigma, 0.0f, 90.0f), sigma2 = sig * sig;     // MatteMaterial clamps to [0, 90]
                materials[i].on_a = 1.0f - (sigma2 / (2.0f * (sigma2 + 0.33f)));
                materials[i].on_b = 0.45f * sigma2 / (sigma2 + 0.09f);
            }
        }
        tri_light.assign(n_prims, -1);
        lights.resize(n_lights);
        for (uint32_t i = 0; i < n_lights; ++i) {
            lights[i].d = lts[i];
            lights[i].area = 0;
            lights[i].cos_total_width = lights[i].cos_falloff_start = lights[i].world_radius = 0;
            lights[i].w_light = V3{0, 0, 0};
            if (lts[i].type == LIGHT_SPOT) {                                                   // spot.rs:38-39, pbrt.rs:133-135
                lights[i].cos_total_width = std::cos(kPi / 180.0f * lts[i].total_width);
                lights[i].cos_falloff_start = std::cos(kPi / 180.0f * lts[i].falloff_start);
            }
            if (lts[i].type == LIGHT_DISTANT) {
                lights[i].w_light = normalize(V3{lts[i].axis[0], lts[i].axis[1], lts[i].axis[2]});   // distant.rs:31
                // Light::pre_process (distant.rs:73-77) + Bounds3::bounding_sphere (geometry.rs:473-480)
                const Bounds3 wb = bvh.world_bound();
                const V3 c = (wb.mn + wb.mx) / 2.0f;
                const bool inside = c.x >= wb.mn.x && c.x <= wb.mx.x && c.y >= wb.mn.y && c.y <= wb.mx.y && c.z >= wb.mn.z && c.z <= wb.mx.z;
                lights[i].world_radius = inside ? length(c - wb.mx) : 0.0f;
            }
            if (lts[i].type == LIGHT_AREA && lts[i].prim_id >= nt) {                            // sphere.rs:100-102
                lights[i].sphere = &bvh.spheres[lts[i].prim_id - nt];
                lights[i].p0 = lights[i].p1 = lights[i].p2 = V3{0, 0, 0};
                lights[i].area = lights[i].sphere->area();
                tri_light[lts[i].prim_id] = (int32_t)i;
            } else if (lts[i].type == LIGHT_AREA) {
                bvh.tri(lts[i].prim_id, &lights[i].p0, &lights[i].p1, &lights[i].p2);
                lights[i].area = length(cross(lights[i].p1 - lights[i].p0, lights[i].p2 - lights[i].p0)) * 0.5f;   // triangle.rs:323-328
                tri_light[lts[i].prim_id] = (int32_t)i;
            }
        }
    }
    // lightdistrib.rs:222-232 + integrator.rs:268-277
    void set_light_strategy(int strategy) {
        light_strategy = strategy;
        spatial_on = strategy == LIGHTS_SPATIAL && lights.size() != 1;                          // :223: one light -> uniform
        if (spatial_on) init_spatial(64);
        std::vector<Float> f(lights.size(), 1.0f);
        if (strategy == LIGHTS_POWER && lights.size() != 1)
            for (size_t i = 0; i < lights.size(); ++i) {
                const LightRt& l = lights[i];
                RGB power;
                if (l.d.type == LIGHT_POINT) power = l.l() * (4.0f * kPi);                                      // point.rs:68-70
                else if (l.d.type == LIGHT_SPOT) power = l.l() * (2.0f * kPi * (1.0f - 0.5f * (l.cos_falloff_start + l.cos_total_width)));   // spot.rs:87-89
                else if (l.d.type == LIGHT_DISTANT) power = l.l() * (kPi * l.world_radius * l.world_radius);    // distant.rs:69-71
                else power = l.l() * ((l.d.two_sided ? 2.0f : 1.0f) * l.area * kPi);                            // diffuse.rs:83-85
                f[i] = y_value(power);
            }
        light_distrib.init(f);
    }

    // ---- SpatialLightDistribution (lightdistrib.rs:71-220), selected by "spatial" when the scene has more than one light --------
    //   D63 FIX  :115-119 `pi as Float + 1.0 / n_voxel` -> (pi + 1) / n_voxel (pbrt-v3: the voxel's far corner)
    // The reference fills a lock-free hash table lazily (:166-219); which entry a voxel lands in cannot change a result, so the
    // restatement keeps a mutex-protected map from the packed voxel position to its distribution.
    int light_strategy = LIGHTS_UNIFORM;
    bool spatial_on = false;
    int n_voxel[3] = {1, 1, 1};
    mutable std::mutex spatial_mu;
    mutable std::unordered_map<uint64_t, std::unique_ptr<Distribution1D>> spatial_map;
    static int float_as_usize(Float r) { return std::isnan(r) || r <= 0.0f ? 0 : (r >= 1.0e9f ? 1000000000 : (int)r); }   // Rust `as usize`: saturating
    void init_spatial(int max_voxels) {                                                        // :83-105
        const Bounds3 b = bvh.world_bound();
        const V3 diag = b.mx - b.mn;
        const Float b_max = diag[max_dimension(diag)];
        for (int i = 0; i < 3; ++i) n_voxel[i] = std::max(1, float_as_usize(std::round(diag[i] / b_max * (Float)max_voxels)));
        spatial_map.clear();
    }
    static V3 bounds_lerp(const Bounds3& b, V3 t) {                                            // geometry.rs:454-458 + pbrt.rs:224-226
        return {(1.0f - t.x) * b.mn.x + t.x * b.mx.x, (1.0f - t.y) * b.mn.y + t.y * b.mx.y, (1.0f - t.z) * b.mn.z + t.z * b.mx.z};
    }
    void spatial_contrib(const int pi[3], std::vector<Float>* contrib) const {                 // :107-158
        const Bounds3 wb = bvh.world_bound();
        const V3 p0{(Float)pi[0] / (Float)n_voxel[0], (Float)pi[1] / (Float)n_voxel[1], (Float)pi[2] / (Float)n_voxel[2]};
        const V3 p1{(Float)(pi[0] + 1) / (Float)n_voxel[0], (Float)(pi[1] + 1) / (Float)n_voxel[1], (Float)(pi[2] + 1) / (Float)n_voxel[2]};   // D63 FIX
        const V3 c0 = bounds_lerp(wb, p0), c1 = bounds_lerp(wb, p1);
        const Bounds3 vb{vmin(c0, c1), vmax(c0, c1)};                                          // geometry.rs:549-559
        const int n_samples = 128;
        contrib->assign(lights.size(), 0.0f);
        for (int i = 0; i < n_samples; ++i) {
            const V3 po = bounds_lerp(vb, V3{radical_inverse_small(0, i), radical_inverse_small(1, i), radical_inverse_small(2, i)});
            const Float u0 = radical_inverse_small(3, i), u1 = radical_inverse_small(4, i);
            for (size_t j = 0; j < lights.size(); ++j) {
                Float pdf = 0.0f;
                const RGB li = light_sample_li(lights[j], po, u0, u1, &pdf);
                if (pdf > 0.0f) (*contrib)[j] += y_value(li) / pdf;
            }
        }
        Float sum_contrib = 0.0f;
        for (Float c : *contrib) sum_contrib += c;
        const Float avg_contrib = sum_contrib / (Float)((size_t)n_samples * contrib->size());
        const Float min_contrib = avg_contrib > 0.0f ? 0.001f * avg_contrib : 1.0f;
        for (Float& c : *contrib) c = fmax_(min_contrib, c);
    }
    void voxel_of(V3 p, int pi[3]) const {                                                     // :165-175 + geometry.rs:460-467
        const Bounds3 wb = bvh.world_bound();
        V3 o = p - wb.mn;
        if (wb.mx.x > wb.mn.x) o.x /= wb.mx.x - wb.mn.x;
        if (wb.mx.y > wb.mn.y) o.y /= wb.mx.y - wb.mn.y;
        if (wb.mx.z > wb.mn.z) o.z /= wb.mx.z - wb.mn.z;
        for (int i = 0; i < 3; ++i) {
            const Float f = o[i] * (Float)n_voxel[i];
            const int v = std::isnan(f) ? 0 : (f >= 2147483648.0f ? 2147483647 : (f <= -2147483648.0f ? -2147483647 - 1 : (int)f));   // Rust `as i32`
            pi[i] = std::min(std::max(v, 0), n_voxel[i] - 1);
        }
    }
    // LightDistribution::lookup (lightdistrib.rs:22-24, :41-43, :63-65, :160-219)
    const Distribution1D& lookup(V3 p) const {
        if (!spatial_on) return light_distrib;
        int pi[3];
        voxel_of(p, pi);
        const uint64_t packed = ((uint64_t)pi[0] << 40) | ((uint64_t)pi[1] << 20) | (uint64_t)pi[2];
        std::lock_guard<std::mutex> lock(spatial_mu);
        std::unique_ptr<Distribution1D>& e = spatial_map[packed];
        if (!e) {
            std::vector<Float> contrib;
            spatial_contrib(pi, &contrib);
            e.reset(new Distribution1D());
            e->init(contrib);
        }
        return *e;
    }

    // Scene::intersect -> SurfaceInteraction (triangle.rs:182-250; D59: shading = geometric)
    bool intersect(Ray& ray, SurfaceInteraction* si) const {
        Hit h;
        Float b0;
        PathCounters* pc = tl_path_counters;
        if (pc) pc->rays[tl_ray_kind]++;
        SphereSI ssi;
        if (!bvh.intersect(ray, &h, &b0, pc ? &pc->trav[tl_ray_kind] : nullptr, &ssi)) return false;
        if (bvh.is_sphere(h.prim_id)) {                                                         // sphere.rs:38-93
            si->p = ssi.p; si->error = ssi.error; si->n = ssi.n; si->wo = ssi.wo;
            si->dpdu = ssi.dpdu; si->sn = ssi.sn; si->sdpdu = ssi.sdpdu;
            si->prim = h.prim_id;
            return true;
        }
        V3 p0, p1, p2;
        bvh.tri(h.prim_id, &p0, &p1, &p2);
        Interaction it = triangle_interaction(p0, p1, p2, b0, h.b1, h.b2);
        si->p = it.p; si->error = it.error; si->n = it.n;
        si->wo = -ray.d;
        V3 du, dv;
        Float uvb[6];
        triangle_frame(p0, p1, p2, &du, &dv, bvh.tri_uv(h.prim_id, uvb));
        si->dpdu = du;
        si->sn = si->n;                                                                         // D59 FIX
        si->sdpdu = du;
        si->prim = h.prim_id;
        if (!vn.empty() || !vs.empty()) {                                                       // triangle.rs:251-311
            const uint32_t* ix = &bvh.indices[3 * (size_t)h.prim_id];
            V3 ns = si->n;
            if (!vn.empty()) {
                ns = (vn[ix[0]] * b0 + vn[ix[1]] * h.b1) + vn[ix[2]] * h.b2;
                ns = length_squared(ns) > 0.0f ? normalize(ns) : si->n;
            }
            V3 ss = normalize(si->dpdu);
            if (!vs.empty()) {
                const V3 st = (vs[ix[0]] * b0 + vs[ix[1]] * h.b1) + vs[ix[2]] * h.b2;
                if (length_squared(st) > 0.0f) ss = normalize(st);
            }
            V3 ts = cross(ss, ns);
            if (length_squared(ts) > 0.0f) { ts = normalize(ts); ss = cross(ts, ns); }
            else coordinate_system(ns, &ss, &ts);
            // set_shading_geometry(ss, ts, .., orientation_is_authoritative = true) (interaction.rs:297-316)
            si->sn = normalize(cross(ss, ts));
            si->n = faceforward(si->n, si->sn);                                                 // D6 FIX
            si->sdpdu = ss;
        }
        return true;
    }
    RGB le(const SurfaceInteraction& si, V3 w) const {                                          // interaction.rs:387-395 + diffuse.rs:150-156
        int32_t li = tri_light[si.prim];
        if (li < 0) return rgb(0);
        const LightRt& l = lights[li];
        return (l.d.two_sided || dot(si.n, w) > 0.0f) ? l.l() : rgb(0);
    }
    // Appendix B: matte / plastic / glass -> BSDF
    BSDF make_bsdf(const SurfaceInteraction& si) const {
        const MaterialRt& m = materials[tri_material[si.prim]];
        BSDF b;
        b.eta = m.d.type == MAT_GLASS ? m.d.eta : 1.0f;
        b.ns = si.sn; b.ng = si.n;                                                              // reflection.rs:220-234
        b.ss = normalize(si.sdpdu);
        b.ts = cross(b.ns, b.ss);
        RGB kd{m.d.kd[0], m.d.kd[1], m.d.kd[2]}, ks{m.d.ks[0], m.d.ks[1], m.d.ks[2]};
        RGB kr{m.d.kr[0], m.d.kr[1], m.d.kr[2]}, kt{m.d.kt[0], m.d.kt[1], m.d.kt[2]};
        if (m.d.type == MAT_MIRROR) {                                                           // pbrt-v3 MirrorMaterial
            if (!is_black(kr)) b.lobes[b.n++] = Lobe{LOBE_SPECULAR_REFLECTION, BSDF_REFLECTION | BSDF_SPECULAR, kr, rgb(0), 0, 1, 1};
        } else if (m.d.type == MAT_SUBSTRATE) {                                                 // pbrt-v3 SubstrateMaterial (isotropic)
            if (!is_black(kd) || !is_black(ks)) b.lobes[b.n++] = Lobe{LOBE_FRESNEL_BLEND, BSDF_REFLECTION | BSDF_GLOSSY, kd, ks, m.alpha, 1, 1};
        } else if (m.d.type == MAT_METAL) {                                                     // pbrt-v3 MetalMaterial (isotropic)
            Lobe l{LOBE_MICROFACET_CONDUCTOR, BSDF_REFLECTION | BSDF_GLOSSY, rgb(1.0f), RGB{m.d.metal_eta[0], m.d.metal_eta[1], m.d.metal_eta[2]}, m.alpha, 1, 1};
            l.k = RGB{m.d.metal_k[0], m.d.metal_k[1], m.d.metal_k[2]};
            b.lobes[b.n++] = l;
        } else if (m.d.type == MAT_MATTE && m.d.sigma != 0.0f) {                                // pbrt-v3 MatteMaterial, sigma != 0
            if (!is_black(kd)) b.lobes[b.n++] = Lobe{LOBE_OREN_NAYAR, BSDF_REFLECTION | BSDF_DIFFUSE, kd, rgb(0), 0, m.on_a, m.on_b};
        } else if (m.d.type == MAT_MATTE || m.d.type == MAT_PLASTIC) {
            if (!is_black(kd)) b.lobes[b.n++] = Lobe{LOBE_LAMBERT, BSDF_REFLECTION | BSDF_DIFFUSE, kd, rgb(0), 0, 1, 1};
            if (m.d.type == MAT_PLASTIC && !is_black(ks))
                b.lobes[b.n++] = Lobe{LOBE_MICROFACET, BSDF_REFLECTION | BSDF_GLOSSY, ks, rgb(0), m.alpha, 1.5f, 1.0f};
        } else if (m.d.type == MAT_GLASS && m.d.roughness != 0.0f) {                            // pbrt-v3 GlassMaterial, rough: two microfacet lobes
            if (!is_black(kr)) b.lobes[b.n++] = Lobe{LOBE_MICROFACET, BSDF_REFLECTION | BSDF_GLOSSY, kr, rgb(0), m.alpha, 1.0f, m.d.eta};
            if (!is_black(kt)) b.lobes[b.n++] = Lobe{LOBE_MICROFACET_TRANSMISSION, BSDF_TRANSMISSION | BSDF_GLOSSY, rgb(0), kt, m.alpha, 1.0f, m.d.eta};
        } else if (!(is_black(kr) && is_black(kt))) {
            b.lobes[b.n++] = Lobe{LOBE_FRESNEL_SPECULAR, BSDF_REFLECTION | BSDF_TRANSMISSION | BSDF_SPECULAR, kr, kt, 0, 1.0f, m.d.eta};
        }
        return b;
    }
};

// ---------------------------------------------------------------- src/core/lowdiscrepancy.rs + src/samplers/halton.rs
// HaltonSampler.  The Rust port does not produce Halton points as written; where it is broken this
// follows pbrt-v3 (core/lowdiscrepancy.h/.cpp, samplers/halton.cpp), the code it is a port of:
//   S1 lowdiscrepancy.rs:293-305 radical_inverse_specialized never accumulates the reversed digits (`let _reversed_digits`)
//      and squares inv_base instead of scaling inv_base_n                                        -> FIX
//   S2 :307-320 scramble_radical_inverse_specialized scales by inv_base instead of inv_base_n   -> FIX
//   S3 sampler.rs:344-396 GlobalSampler::get_1d/get_2d run on the base struct, which cannot reach the Halton
//      overrides of get_index_for_sample / sample_dimension (no virtual dispatch through Deref)  -> FIX (dispatch)
//   S4 halton.rs:122-125 `current_pixel % MAX_RESOLUTION` is negative for negative pixels         -> FIX (pbrt-v3 Mod)
//   S5 the table of the first 1000 primes (lowdiscrepancy.rs:12-76) is generated by a sieve here, not copied
struct HaltonTables {
    static constexpr int kPrimeTableSize = 1000;              // lowdiscrepancy.rs:11
    static constexpr int kMaxResolution = 128;                // halton.rs:39
    std::vector<uint32_t> primes, prime_sums;
    std::vector<uint16_t> perms;                              // RADICAL_INVERSE_PERMUTATIONS (halton.rs:17-22)
    int base_scales[2], base_exponents[2];
    uint64_t sample_stride, mult_inverse[2];

    static void extended_gcd(uint64_t a, uint64_t b, int64_t* x, int64_t* y) {       // halton.rs:52-62
        if (b == 0) { *x = 1; *y = 0; return; }
        int64_t d = (int64_t)(a / b), xp, yp;
        extended_gcd(b, a % b, &xp, &yp);
        *x = yp;
        *y = xp - d * yp;
    }
    static uint64_t multiplicative_inverse(int64_t a, int64_t n) {                   // halton.rs:41-50
        int64_t x, y;
        extended_gcd((uint64_t)a, (uint64_t)n, &x, &y);
        int64_t r = x - (x / n) * n;
        return (uint64_t)(r < 0 ? r + n : r);
    }
    void init(int res_x, int res_y) {                         // HaltonSampler::new, halton.rs:64-103
        primes.clear();
        for (uint32_t c = 2; (int)primes.size() < kPrimeTableSize; ++c) {
            bool is_prime = true;
            for (uint32_t p : primes) { if (p * p > c) break; if (c % p == 0) { is_prime = false; break; } }
            if (is_prime) primes.push_back(c);
        }
        prime_sums.assign(kPrimeTableSize, 0);
        for (int i = 1; i < kPrimeTableSize; ++i) prime_sums[i] = prime_sums[i - 1] + primes[i - 1];
        // compute_radical_inverse_permutations (lowdiscrepancy.rs:333-349) with RNG::default(); shuffle = sampling.rs:280-287
        perms.resize(prime_sums.back() + primes.back());
        RNG rng;
        size_t off = 0;
        for (int i = 0; i < kPrimeTableSize; ++i) {
            const uint32_t n = primes[i];
            for (uint32_t j = 0; j < n; ++j) perms[off + j] = (uint16_t)j;
            for (uint32_t j = 0; j < n; ++j) {
                const uint32_t other = j + uniform_u32_bounded(rng, n - j);
                std::swap(perms[off + j], perms[off + other]);
            }
            off += n;
        }
        const int res[2] = {res_x, res_y};
        for (int i = 0; i < 2; ++i) {
            const int base = i == 0 ? 2 : 3;
            int scale = 1, exp = 0;
            while (scale < std::min(kMaxResolution, res[i])) { scale *= base; ++exp; }
            base_scales[i] = scale;
            base_exponents[i] = exp;
        }
        sample_stride = (uint64_t)base_scales[0] * (uint64_t)base_scales[1];
        mult_inverse[0] = multiplicative_inverse(base_scales[1], base_scales[0]);
        mult_inverse[1] = multiplicative_inverse(base_scales[0], base_scales[1]);
    }
    static uint32_t uniform_u32_bounded(RNG& rng, uint32_t b) {                      // rng.rs:36-44
        const uint32_t threshold = (~b + 1u) % b;
        for (;;) {
            const uint32_t r = rng.uniform_u32();
            if (r >= threshold) return r % b;
        }
    }
    static uint32_t reverse_bits32(uint32_t n) {                                     // lowdiscrepancy.rs:371-379
        n = (n << 16) | (n >> 16);
        n = ((n & 0x00ff00ffu) << 8) | ((n & 0xff00ff00u) >> 8);
        n = ((n & 0x0f0f0f0fu) << 4) | ((n & 0xf0f0f0f0u) >> 4);
        n = ((n & 0x33333333u) << 2) | ((n & 0xccccccccu) >> 2);
        n = ((n & 0x55555555u) << 1) | ((n & 0xaaaaaaaau) >> 1);
        return n;
    }
    static uint64_t reverse_bits64(uint64_t n) {                                     // :364-369
        return ((uint64_t)reverse_bits32((uint32_t)n) << 32) | (uint64_t)reverse_bits32((uint32_t)(n >> 32));
    }
    static uint64_t inverse_radical_inverse(uint64_t base, uint64_t inverse, int n_digits) {      // :381-390
        uint64_t index = 0;
        for (int i = 0; i < n_digits; ++i) {
            const uint64_t digit = inverse % base;
            inverse /= base;
            index = index * base + digit;
        }
        return index;
    }
    Float radical_inverse(int base_index, uint64_t a) const {                       // :322-331, S1
        if (base_index == 0) return (Float)reverse_bits64(a) * 5.4210108624275222e-20f;
        const uint64_t base = primes[base_index];
        const Float inv_base = 1.0f / (Float)base;
        uint64_t reversed = 0;
        Float inv_base_n = 1.0f;
        while (a != 0) {
            const uint64_t next = a / base, digit = a - next * base;
            reversed = reversed * base + digit;
            inv_base_n *= inv_base;
            a = next;
        }
        return fmin_((Float)reversed * inv_base_n, kOneMinusEpsilon);
    }
    Float scrambled_radical_inverse(int base_index, uint64_t a) const {             // :307-320,:351-362, S2
        const uint64_t base = primes[base_index];
        const uint16_t* perm = perms.data() + prime_sums[base_index];                // halton.rs:105-113
        const Float inv_base = 1.0f / (Float)base;
        uint64_t reversed = 0;
        Float inv_base_n = 1.0f;
        while (a != 0) {
            const uint64_t next = a / base, digit = a - next * base;
            reversed = reversed * base + perm[digit];
            inv_base_n *= inv_base;
            a = next;
        }
        return fmin_(inv_base_n * ((Float)reversed + inv_base * (Float)perm[0] / (1.0f - inv_base)), kOneMinusEpsilon);
    }
    // HaltonSampler::get_index_for_sample (halton.rs:117-141)
    int64_t index_for_sample(int px, int py, uint64_t sample_num) const {
        int64_t offset = 0;
        if (sample_stride > 1) {
            const int pm[2] = {((px % kMaxResolution) + kMaxResolution) % kMaxResolution, ((py % kMaxResolution) + kMaxResolution) % kMaxResolution};   // S4
            for (int i = 0; i < 2; ++i) {
                const uint64_t dim_offset = inverse_radical_inverse(i == 0 ? 2 : 3, (uint64_t)pm[i], base_exponents[i]);
                offset += (int64_t)(dim_offset * (sample_stride / (uint64_t)base_scales[i]) * mult_inverse[i]);
            }
            offset %= (int64_t)sample_stride;
        }
        return offset + (int64_t)(sample_num * sample_stride);
    }
    // HaltonSampler::sample_dimension (halton.rs:143-155), sample_at_pixel_center = false
    Float sample_dimension(int64_t index, int dim) const {
        if (dim == 0) return radical_inverse(0, (uint64_t)index >> base_exponents[0]);
        if (dim == 1) return radical_inverse(1, (uint64_t)index / (uint64_t)base_scales[1]);
        return scrambled_radical_inverse(dim, (uint64_t)index);
    }
};

// ---------------------------------------------------------------- src/samplers/sobol.rs + lowdiscrepancy.rs:507-560
// SobolSampler, a GlobalSampler like Halton: sample index = sobol_interval_to_index(log2 resolution, sample
// number, pixel - sample_bounds.min), dimension d = the index's bits times the d-th 32x52 generator matrix.  The generator
// matrices are constant data (Joe & Kuo direction numbers): tools/make_sobol_tables.py converts the reference's
// src/core/sobolmatrices.rs into pbrt-rs_b200/data/sobol_tables.bin, embedded here (oracle_capi.cpp) and in the library.
// Where the port cannot run this follows pbrt-v3 (samplers/sobol.cpp, core/lowdiscrepancy.h):
//   Q1 lowdiscrepancy.rs:529-534 the second loop of sobol_interval_to_index never shifts `b` nor advances `c`: it cannot
//      terminate                                                                                   -> FIX (b >>= 1, c += 1)
//   Q2 samplers/sobol.rs:56-58 SobolSampler::sample_dimension is todo!()                          -> FIX: pbrt-v3 SampleDimension
//      (sobol_sample(index, dim, 0); dimensions 0 / 1 remapped to the pixel: s * resolution + sample_bounds.min, minus the
//      pixel, clamped to [0, 1 - eps])
//   Q3 the non-float64 branch of sobol_sample (:553-560) is used as written (f32 build)
struct SobolTables {
    uint32_t n_dims = 0, size = 0, n_vdc = 0, n_inv = 0;
    const uint32_t* m32 = nullptr;
    const uint64_t *vdc = nullptr, *vdc_inv = nullptr;
    int sb_min[2] = {0, 0};
    int resolution = 1, log2_resolution = 0;
    bool load(const unsigned char* blob, size_t bytes) {
        if (bytes < 32) return false;
        uint32_t h[8];
        std::memcpy(h, blob, 32);
        if (h[0] != 0x31424F53u) return false;
        n_dims = h[1]; size = h[2]; n_vdc = h[3]; n_inv = h[4];
        const size_t need = 32 + (size_t)n_dims * size * 4 + ((size_t)n_vdc + n_inv) * size * 8;
        if (bytes < need) return false;
        m32 = reinterpret_cast<const uint32_t*>(blob + 32);
        vdc = reinterpret_cast<const uint64_t*>(blob + 32 + (size_t)n_dims * size * 4);
        vdc_inv = vdc + (size_t)n_vdc * size;
        return true;
    }
    void init(int sb_x0, int sb_y0, int sb_w, int sb_h) {                               // SobolSampler::new, sobol.rs:20-37
        sb_min[0] = sb_x0; sb_min[1] = sb_y0;
        int v = std::max(sb_w, sb_h), r = 1, l = 0;                                      // round_up_pow2_i32 / log_2_int_i32
        while (r < v) { r <<= 1; ++l; }
        resolution = r;
        log2_resolution = l;
    }
    uint64_t interval_to_index(uint32_t m, uint64_t frame, int px, int py) const {      // lowdiscrepancy.rs:507-536, Q1
        if (m == 0) return 0;
        const uint32_t m2 = m << 1;
        uint64_t index = frame << m2;
        uint64_t delta = 0;
        for (int c = 0; frame != 0; frame >>= 1, ++c)
            if (frame & 1) delta ^= vdc[(size_t)(m - 1) * size + c];
        uint64_t b = (uint64_t)((((uint32_t)px) << m) | (uint32_t)py) ^ delta;
        for (int c = 0; b != 0; b >>= 1, ++c)
            if (b & 1) index ^= vdc_inv[(size_t)(m - 1) * size + c];
        return index;
    }
    Float sample(int64_t a, int dimension) const {                                      // sobol_sample :538-560, scramble = 0
        uint32_t v = 0;
        for (size_t i = (size_t)dimension * size; a != 0; a >>= 1, ++i)
            if (a & 1) v ^= m32[i];
        return fmin_(kOneMinusEpsilon, (Float)v * 2.3283064365386963e-10f);
    }
    int64_t index_for_sample(int px, int py, uint64_t sample_num) const {               // sobol.rs:48-54
        return (int64_t)interval_to_index((uint32_t)log2_resolution, sample_num, px - sb_min[0], py - sb_min[1]);
    }
    Float sample_dimension(int64_t index, int dim, int px, int py) const {              // Q2: pbrt-v3 SobolSampler::SampleDimension
        Float s = sample(index, dim);
        if (dim == 0 || dim == 1) {
            s = s * (Float)resolution + (Float)sb_min[dim];
            s = clampf(s - (Float)(dim == 0 ? px : py), 0.0f, kOneMinusEpsilon);
        }
        return s;
    }
};
// the embedded table (oracle_capi.cpp); nullptr when the blob is missing or malformed
const SobolTables* sobol_tables_base();

// PixelSampler (sampler.rs:257-322): n_sampled_dimensions tabulated 1D and 2D dimensions of spp values each, filled by
// start_pixel of StratifiedSampler (samplers/stratified.rs:44-105) or ZeroTwoSequenceSampler (samplers/zerotwosequence.rs:31-63);
// no sample arrays are requested on this path (PathIntegrator never calls request_*_array).  Where the port cannot run,
// this follows pbrt-v3 (samplers/stratified.cpp, samplers/zerotwosequence.cpp, core/lowdiscrepancy.h):
//   P1 lowdiscrepancy.rs:452-459 van_der_corput shuffles `samples[i * n_pixel_samples..]`: out of range for i >= 1  -> FIX (i * n_samples_per_pixel_sample)
//   P2 stratified.rs:44-105 start_pixel never resets current_pixel_sample_index (no PixelSampler::start_pixel call)  -> FIX (reset)
//   P3 sampler.rs:443-445 impl_pixel_sampler!::set_sample_number calls itself                                        -> FIX (PixelSampler's)
//   P4 the second Sobol' generator matrix (lowdiscrepancy.rs:480-486) is generated (c[j] = c[j-1] ^ (c[j-1] >> 1)), not copied
struct PixelTables {
    int n_dims = 0, spp = 0;
    std::vector<Float> t1, t2;                                   // t1[dim * spp + s]; t2[(dim * spp + s) * 2 + {0,1}]
    void resize(int dims, int samples) { n_dims = dims; spp = samples; t1.assign((size_t)dims * samples, 0.0f); t2.assign((size_t)dims * samples * 2, 0.0f); }
    static void shuffle1(Float* a, int count, int n_dimensions, RNG& rng) {                    // sampling.rs:280-287
        for (int i = 0; i < count; ++i) {
            const int other = i + (int)HaltonTables::uniform_u32_bounded(rng, (uint32_t)(count - i));
            for (int j = 0; j < n_dimensions; ++j) std::swap(a[n_dimensions * i + j], a[n_dimensions * other + j]);
        }
    }
    // StratifiedSampler::start_pixel (stratified.rs:44-76)
    void start_pixel_stratified(RNG& rng, int xs, int ys, bool jitter) {
        const int n = xs * ys;
        for (int d = 0; d < n_dims; ++d) {
            Float* a = t1.data() + (size_t)d * spp;
            const Float inv_n = 1.0f / (Float)n;                                               // sampling.rs:11-17
            for (int i = 0; i < n; ++i) {
                const Float delta = jitter ? rng.uniform_float() : 0.5f;
                a[i] = fmin_(kOneMinusEpsilon, ((Float)i + delta) * inv_n);
            }
            shuffle1(a, n, 1, rng);
        }
        for (int d = 0; d < n_dims; ++d) {
            Float* a = t2.data() + (size_t)d * spp * 2;
            const Float dx = 1.0f / (Float)xs, dy = 1.0f / (Float)ys;                          // sampling.rs:19-41
            int i = 0;
            for (int y = 0; y < ys; ++y)
                for (int x = 0; x < xs; ++x) {
                    Float jx = 0.5f, jy = 0.5f;
                    if (jitter) { jx = rng.uniform_float(); jy = rng.uniform_float(); }
                    a[2 * i] = fmin_(kOneMinusEpsilon, ((Float)x + jx) * dx);
                    a[2 * i + 1] = fmin_(kOneMinusEpsilon, ((Float)y + jy) * dy);
                    ++i;
                }
            shuffle1(a, n, 2, rng);                                                            // whole Point2f elements swap
        }
    }
    // ZeroTwoSequenceSampler::start_pixel (zerotwosequence.rs:31-48) for n_samples_per_pixel_sample = 1
    void start_pixel_zerotwo(RNG& rng) {
        for (int d = 0; d < n_dims; ++d) {                                                     // van_der_corput, lowdiscrepancy.rs:436-460
            Float* a = t1.data() + (size_t)d * spp;
            uint32_t v = rng.uniform_u32();
            for (int i = 0; i < spp; ++i) {                                                    // gray_code_sample :416-422, C[j] = 1 << (31 - j)
                a[i] = fmin_(kOneMinusEpsilon, (Float)v * 2.3283064365386963e-10f);
                v ^= 0x80000000u >> __builtin_ctz((uint32_t)(i + 1));
            }
            for (int i = 0; i < spp; ++i) shuffle1(a + i, 1, 1, rng);                          // P1 FIX
            shuffle1(a, spp, 1, rng);
        }
        uint32_t c1[32];                                                                       // P4
        c1[0] = 0x80000000u;
        for (int j = 1; j < 32; ++j) c1[j] = c1[j - 1] ^ (c1[j - 1] >> 1);
        for (int d = 0; d < n_dims; ++d) {                                                     // sobol_2d, lowdiscrepancy.rs:462-505
            Float* a = t2.data() + (size_t)d * spp * 2;
            uint32_t v0 = rng.uniform_u32(), v1 = rng.uniform_u32();
            for (int i = 0; i < spp; ++i) {                                                    // gray_code_sample_2d :425-434
                a[2 * i] = fmin_(kOneMinusEpsilon, (Float)v0 * 2.3283064365386963e-10f);
                a[2 * i + 1] = fmin_(kOneMinusEpsilon, (Float)v1 * 2.3283064365386963e-10f);
                const int tz = __builtin_ctz((uint32_t)(i + 1));
                v0 ^= 0x80000000u >> tz;
                v1 ^= c1[tz];
            }
            for (int i = 0; i < spp; ++i) shuffle1(a + 2 * i, 1, 2, rng);
            shuffle1(a, spp, 2, rng);
        }
    }
    void start_pixel(int kind, RNG& rng, const PathDesc& pd) {
        if (kind == 2) start_pixel_stratified(rng, pd.x_samples, pd.y_samples, pd.jitter != 0);
        else if (kind == 3) start_pixel_zerotwo(rng);
    }
};

struct Sampler {          // kind 0: RandomSampler (samplers/random.rs:29-56), every dimension straight from PCG32;
    RNG rng;              // kind 1: HaltonSampler through GlobalSampler::get_1d/get_2d (sampler.rs:371-389; no sample arrays)
    int kind = 0;         // kind 2 / 3: PixelSampler::get_1d/get_2d (sampler.rs:289-307): tables first, then `rng`
    const HaltonTables* halton = nullptr;
    const PixelTables* tabs = nullptr;
    const SobolTables* sobol = nullptr;                        // kind 4: SobolSampler, a GlobalSampler like Halton
    int64_t index = 0;
    int dimension = 0;
    int pix_x = 0, pix_y = 0;                                  // current_pixel (Sobol' remaps dimensions 0 / 1 to it)
    int cur1 = 0, cur2 = 0, sample_index = 0;                  // current_1d_dimension, current_2d_dimension, current_pixel_sample_index
    bool tabulated() const { return kind == 2 || kind == 3; }
    void start_sample(int px, int py, uint64_t sample_num) {   // start_pixel / set_sample_number (sampler.rs:347-350,405-409,316-321)
        if (kind == 1) { index = halton->index_for_sample(px, py, sample_num); dimension = 0; }
        if (kind == 4) { index = sobol->index_for_sample(px, py, sample_num); dimension = 0; pix_x = px; pix_y = py; }
        if (tabulated()) { cur1 = cur2 = 0; sample_index = (int)sample_num; }
    }
    Float global_dimension(int dim) const { return kind == 1 ? halton->sample_dimension(index, dim) : sobol->sample_dimension(index, dim, pix_x, pix_y); }
    Float get_1d() {
        if (kind == 1 || kind == 4) return global_dimension(dimension++);
        if (tabulated() && cur1 < tabs->n_dims) return tabs->t1[(size_t)(cur1++) * tabs->spp + sample_index];
        return rng.uniform_float();
    }
    void get_2d(Float* a, Float* b) {                          // x then y
        if (kind == 1 || kind == 4) { *a = global_dimension(dimension); *b = global_dimension(dimension + 1); dimension += 2; return; }
        if (tabulated() && cur2 < tabs->n_dims) {
            const size_t o = ((size_t)(cur2++) * tabs->spp + sample_index) * 2;
            *a = tabs->t2[o]; *b = tabs->t2[o + 1];
            return;
        }
        *a = rng.uniform_float();
        *b = rng.uniform_float();
    }
};

// integrator.rs:136-266 (handle_media = false, specular = false)
inline RGB estimate_direct(const Scene& scene, const SurfaceInteraction& it, const BSDF& bsdf, Float us0, Float us1, const LightRt& light,
                           uint32_t light_index, Float ul0, Float ul1) {
    const uint8_t flags = BSDF_ALL & ~BSDF_SPECULAR;                                            // D23 FIX
    RGB ld = rgb(0);
    V3 wi{0, 0, 0};
    Float light_pdf = 0.0f, scattering_pdf = 0.0f;
    RGB li = rgb(0);
    Ray shadow{};
    // ---- Light::sample_li ----
    if (light.is_delta()) {                                                                    // point.rs:47-66, spot.rs:71-85, distant.rs:50-67
        V3 pl{light.d.p[0], light.d.p[1], light.d.p[2]};
        light_pdf = 1.0f;
        if (light.d.type == LIGHT_DISTANT) {
            wi = light.w_light;
            pl = it.p + light.w_light * (2.0f * light.world_radius);                           // p_outside
            li = light.l();
        } else {
            wi = normalize(pl - it.p);
            if (light.d.type == LIGHT_SPOT) li = light.l() * light.falloff(-wi) / length_squared(pl - it.p);
            else li = light.l() / length_squared(pl - it.p);
        }
        // VisibilityTester: spawn_ray_to(&BaseInteraction) with a bare point (interaction.rs:146-153)
        V3 origin = offset_ray_origin(it.p, it.error, it.n, pl - it.p);
        V3 target = offset_ray_origin(pl, V3{0, 0, 0}, V3{0, 0, 0}, origin - pl);
        shadow = Ray{origin, 1.0f - kShadowEpsilon, target - origin, 0.0f};
    } else if (light.sphere) {                                                                 // diffuse.rs:60-81 + sphere.rs:127-193
        V3 ps, pe, ns;
        Float pdf;
        light.sphere->sample2(it.p, it.error, it.n, ul0, ul1, &ps, &pe, &ns, &pdf);
        light_pdf = pdf;
        if (pdf == 0.0f || length_squared(ps - it.p) == 0.0f) { light_pdf = 0.0f; li = rgb(0); }
        else {
            wi = normalize(ps - it.p);
            li = (light.d.two_sided || dot(ns, -wi) > 0.0f) ? light.l() : rgb(0);
            V3 origin = offset_ray_origin(it.p, it.error, it.n, ps - it.p);
            V3 target = offset_ray_origin(ps, pe, ns, origin - ps);
            shadow = Ray{origin, 1.0f - kShadowEpsilon, target - origin, 0.0f};
        }
    } else {                                                                                   // diffuse.rs:60-81 + shape.rs:38-53 + triangle.rs:330-348
        Float su0 = std::sqrt(ul0);
        Float b0 = 1.0f - su0, b1 = ul1 * su0;                                                 // sampling.rs:275-278
        V3 ps = (light.p0 * b0 + light.p1 * b1) + light.p2 * ((1.0f - b0) - b1);
        V3 ns = normalize(cross(light.p1 - light.p0, light.p2 - light.p0));
        if (light.has_n) ns = faceforward(ns, (light.n0 * b0 + light.n1 * b1) + light.n2 * ((1.0f - b0) - b1));   // triangle.rs:338-341, D6 FIX
        V3 pe = ((vabs(light.p0 * b0) + vabs(light.p1 * b1)) + vabs(light.p2 * ((1.0f - b0) - b1))) * gamma(6.0f);
        Float pdf = 1.0f / light.area;
        V3 w = ps - it.p;
        if (length_squared(w) == 0.0f) pdf = 0.0f;
        else {
            w = normalize(w);
            pdf *= length_squared(it.p - ps) / std::fabs(dot(ns, -w));
            if (std::isinf(pdf)) pdf = 0.0f;
        }
        light_pdf = pdf;
        if (pdf == 0.0f || length_squared(ps - it.p) == 0.0f) { light_pdf = 0.0f; li = rgb(0); }
        else {
            wi = normalize(ps - it.p);
            li = (light.d.two_sided || dot(ns, -wi) > 0.0f) ? light.l() : rgb(0);             // D55 FIX
            V3 origin = offset_ray_origin(it.p, it.error, it.n, ps - it.p);
            V3 target = offset_ray_origin(ps, pe, ns, origin - ps);
            shadow = Ray{origin, 1.0f - kShadowEpsilon, target - origin, 0.0f};
        }
    }
    if (light_pdf > 0.0f && !is_black(li)) {
        scattering_pdf = bsdf.pdf(it.wo, wi, flags);
        RGB f = bsdf.f(it.wo, wi, flags) * std::fabs(dot(wi, bsdf.ns));
        if (!is_black(f)) {
            PathCounters* pc = tl_path_counters;
            if (pc) pc->rays[1]++;
            if (scene.bvh.intersect_p(shadow, pc ? &pc->trav[1] : nullptr)) li = rgb(0);       // light.rs:126-135, D25 FIX
            if (!is_black(li)) {
                if (light.is_delta()) ld = ld + li * f / light_pdf;
                else ld = ld + li * f * power_heuristic(light_pdf, scattering_pdf) / light_pdf;
            }
        }
    }
    // ---- BSDF sampling with MIS ----
    if (!light.is_delta()) {
        uint8_t sampled = 0;
        RGB f = bsdf.sample_f(it.wo, &wi, us0, us1, &scattering_pdf, flags, &sampled);
        f = f * std::fabs(dot(wi, bsdf.ns));
        bool sampled_specular = (sampled & BSDF_SPECULAR) != 0;
        if (!is_black(f) && scattering_pdf > 0.0f) {
            Float weight = 1.0f;
            Interaction base{it.p, it.error, it.n};
            if (!sampled_specular && light.sphere) {                                           // Light::pdf_li -> Sphere::pdf2 (sphere.rs:195-207)
                light_pdf = light.sphere->pdf2(it.p, it.error, it.n, wi);
                if (light_pdf == 0.0f) return ld;
                weight = power_heuristic(scattering_pdf, light_pdf);
            } else if (!sampled_specular) {
                // Light::pdf_li -> Shape::pdf2 (shape.rs:54-69): intersect the light's own triangle
                Ray r = spawn_ray(base, wi);
                TriHit th = triangle_intersect_test(light.p0, light.p1, light.p2, r);
                V3 du, dv;
                if (!th.hit || !triangle_frame(light.p0, light.p1, light.p2, &du, &dv, light.has_uv ? light.uv : nullptr)) return ld;
                Interaction li_it = triangle_interaction(light.p0, light.p1, light.p2, th.b0, th.b1, th.b2);
                Float lp = length_squared(it.p - li_it.p) / (std::fabs(dot(li_it.n, -wi)) * light.area);
                if (std::isinf(lp)) lp = 0.0f;
                light_pdf = lp;
                if (light_pdf == 0.0f) return ld;
                weight = power_heuristic(scattering_pdf, light_pdf);
            }
            Ray ray = spawn_ray(base, wi);
            SurfaceInteraction light_isect;
            RGB lmis = rgb(0);
            tl_ray_kind = 2;
            const bool mis_found = scene.intersect(ray, &light_isect);
            tl_ray_kind = 0;
            if (mis_found) {
                if (scene.tri_light[light_isect.prim] == (int32_t)light_index) lmis = scene.le(light_isect, -wi);   // D56 FIX
            }
            if (!is_black(lmis)) ld = ld + lmis * f * rgb(1.0f) * weight / scattering_pdf;
        }
    }
    return ld;
}

// integrator.rs:92-134
inline RGB uniform_sample_one_light(const Scene& scene, const SurfaceInteraction& it, const BSDF& bsdf, Sampler& s) {
    if (scene.lights.empty()) return rgb(0);
    Float light_pdf;
    size_t num = scene.lookup(it.p).sample_discrete(s.get_1d(), &light_pdf);                  // path.rs:100-104
    if (light_pdf == 0.0f) return rgb(0);
    Float ul0, ul1, us0, us1;
    s.get_2d(&ul0, &ul1);
    s.get_2d(&us0, &us1);
    return estimate_direct(scene, it, bsdf, us0, us1, scene.lights[num], (uint32_t)num, ul0, ul1) / light_pdf;
}

// path.rs:65-213 (BSSRDF branch dead: no subsurface material exists)
inline RGB path_li(const Scene& scene, Ray ray, Sampler& s, int max_depth, Float rr_threshold) {
    RGB l = rgb(0), beta = rgb(1);
    bool specular_bounce = false;
    int bounces = 0;
    Float eta_scale = 1.0f;
    if (tl_path_counters) tl_path_counters->camera_samples++;
    for (;;) {
        SurfaceInteraction isect;
        bool found = scene.intersect(ray, &isect);
        if (bounces == 0 || specular_bounce)
            if (found) l = l + beta * scene.le(isect, -ray.d);
        if (found && tl_path_counters) tl_path_counters->vertices++;
        if (!found || bounces >= max_depth) break;
        BSDF bsdf = scene.make_bsdf(isect);
        if (bsdf.num_components(BSDF_ALL & ~BSDF_SPECULAR) > 0) l = l + beta * uniform_sample_one_light(scene, isect, bsdf, s);
        V3 wo = -ray.d, wi{0, 0, 0};
        Float pdf = 0.0f, u0, u1;
        uint8_t flags = 0;
        s.get_2d(&u0, &u1);
        RGB f = bsdf.sample_f(wo, &wi, u0, u1, &pdf, BSDF_ALL, &flags);
        if (is_black(f) || pdf == 0.0f) break;
        beta = beta * (f * (std::fabs(dot(wi, bsdf.ns)) / pdf));
        specular_bounce = (flags & BSDF_SPECULAR) != 0;
        if ((flags & BSDF_SPECULAR) && (flags & BSDF_TRANSMISSION)) {
            Float eta = bsdf.eta;
            eta_scale *= (dot(wo, isect.n) > 0.0f) ? (eta * eta) : 1.0f / (eta * eta);
        }
        ray = spawn_ray(Interaction{isect.p, isect.error, isect.n}, wi);
        RGB rr_beta = beta * eta_scale;
        if (max_component_value(rr_beta) < rr_threshold && bounces > 3) {
            Float q = fmin_(1.0f - max_component_value(rr_beta), 0.05f);                        // D27 KEEP
            if (s.get_1d() < q) break;
            beta = beta / (1.0f - q);
        }
        bounces += 1;
    }
    return l;
}

// ---------------------------------------------------------------- src/integrators/volpath.rs + src/media/homogeneous.rs
// VolPathIntegrator over HomogeneousMedium.  KEEP/FIX ledger (pbrt-v3 where the Rust code cannot work):
//   D69 FIX  interaction.rs:132-153  spawn_ray / spawn_ray_to hand the new ray `None` as its medium          -> GetMedium(d): outside if
//                                    dot(d, n) > 0 else inside (pbrt-v3 Interaction::GetMedium)
//   D70 FIX  volpath.rs:127-131      `bounces -= 1; continue` skips the loop's `bounces += 1` (and underflows usize at 0)
//                                    -> a material-less surface does not count as a bounce (pbrt-v3's for-loop)
//   D71 FIX  homogeneous.rs:36-38    tr: `.max(Float::MAX)`                                                  -> min
//   D72 FIX  homogeneous.rs:45       sample: t = -(dist / |d|).min(t_max)                                    -> min(dist / |d|, t_max)
//   D73 FIX  homogeneous.rs:56       sample: tr = -sigma_t * min(t, MAX) * |d| without the exponential      -> exp(..)
//   D74 FIX  light.rs:151 / scene.rs:62  the VisibilityTester / intersect_tr loops continue with rays that lost their medium
//                                    (D69) and t_max                                                         -> pbrt-v3
//   KEEP     volpath.rs:96-113       the phase function is sampled (sampler.get_2d) BEFORE uniform_sample_one_light draws its
//                                    three values (pbrt-v3 draws them in the other order)
//   KEEP     volpath.rs:236          Russian roulette q = max(1 - max_component, 0.05) (path.rs has min: D27)
//   KEEP     medium.rs:75-87         HenyeyGreenstein::sample_p as written (= pbrt-v3)
// exp / ln: f32::exp / f32::ln are platform libm; the numerics contract fixes Cephes expf / logf with every operation one
// rounded f32 op (exp_c, log_c), shared with the kernels like sin / cos / acos / atan2.
inline Float exp_c(Float x) {
    if (x > 88.0f) return kInfinity;
    if (x < -87.0f) return 0.0f;                               // below the normal range: flushed (both sides)
    Float z = std::floor(1.44269504088896341f * x + 0.5f);
    x = x - z * 0.693359375f;
    x = x - z * -2.12194440e-4f;
    const int n = (int)z;
    z = x * x;
    z = (((((1.9875691500e-4f * x + 1.3981999507e-3f) * x + 8.3334519073e-3f) * x + 4.1665795894e-2f) * x + 1.6666665459e-1f) * x + 5.0000001201e-1f) * z + x + 1.0f;
    return z * bits_to_float((uint32_t)(n + 127) << 23);       // n in [-126, 127] for |x| <= 88
}
inline Float log_c(Float x) {                                   // x > 0, normal
    if (x <= 0.0f) return -kInfinity;
    const uint32_t u = float_to_bits(x);
    int e = (int)(u >> 23) - 126;
    Float m = bits_to_float((u & 0x007FFFFFu) | 0x3F000000u);  // frexp: m in [0.5, 1)
    if (m < 0.707106781186547524f) { e -= 1; m = (m + m) - 1.0f; }
    else m = m - 1.0f;
    const Float z = m * m;
    Float y = ((((((((7.0376836292e-2f * m - 1.1514610310e-1f) * m + 1.1676998740e-1f) * m - 1.2420140846e-1f) * m + 1.4249322787e-1f) * m - 1.6668057665e-1f) * m +
                 2.0000714765e-1f) * m - 2.4999993993e-1f) * m + 3.3333331174e-1f) * m * z;
    const Float fe = (Float)e;
    y = y + -2.12194440e-4f * fe;
    y = y + -0.5f * z;
    Float r = m + y;
    r = r + 0.693359375f * fe;
    return r;
}

// media/homogeneous.rs:36-38 (D71)
inline RGB medium_tr(const MediumDesc& m, Float t_max, V3 d) {
    const Float s = fmin_(t_max * length(d), 3.402823466e+38f);
    return RGB{exp_c(-((m.sigma_s[0] + m.sigma_a[0]) * s)), exp_c(-((m.sigma_s[1] + m.sigma_a[1]) * s)), exp_c(-((m.sigma_s[2] + m.sigma_a[2]) * s))};
}
// medium.rs:34-37
inline Float phase_hg(Float cos_theta, Float g) {
    const Float denom = 1.0f + g * g + 2.0f * g * cos_theta;
    return (1.0f / kPi / 4.0f) * (1.0f - g * g) / (denom * std::sqrt(denom));        // INV_4_PI = INV_PI / 4 (pbrt.rs:19-21)
}
// medium.rs:71-87
inline Float hg_sample_p(Float g, V3 wo, V3* wi, Float u0, Float u1) {
    Float cos_theta;
    if (std::fabs(g) < 1e-3f) cos_theta = 1.0f - 2.0f * u0;
    else {
        const Float sqr_term = (1.0f - g * g) / (1.0f + g - 2.0f * g * u0);
        cos_theta = -(1.0f + g * g - sqr_term * sqr_term) / (2.0f * g);
    }
    const Float sin_theta = std::sqrt(fmax_(1.0f - cos_theta * cos_theta, 0.0f));
    const Float phi = 2.0f * kPi * u1;
    V3 v1, v2;
    coordinate_system(wo, &v1, &v2);
    Float sp, cp;
    sincos_contract(phi, &sp, &cp);
    *wi = (v1 * sin_theta * cp + v2 * sin_theta * sp) + wo * cos_theta;                 // geometry.rs:1156-1165
    return phase_hg(cos_theta, g);
}
// media/homogeneous.rs:40-74 (D72, D73): returns the throughput factor; *sampled = a medium interaction at *p_out
inline RGB medium_sample(const MediumDesc& m, const Ray& ray, Sampler& s, bool* sampled, V3* p_out) {
    const Float sig_t[3] = {m.sigma_s[0] + m.sigma_a[0], m.sigma_s[1] + m.sigma_a[1], m.sigma_s[2] + m.sigma_a[2]};
    const Float uc = s.get_1d() * 3.0f;
    const int channel = std::min(std::isnan(uc) || uc <= 0.0f ? 0 : (int)uc, 2);
    const Float dist = -log_c(1.0f - s.get_1d()) / sig_t[channel];
    const Float len = length(ray.d);
    const Float t = fmin_(dist / len, ray.t_max);
    *sampled = t < ray.t_max;
    if (*sampled) *p_out = ray.o + ray.d * t;
    const Float tt = fmin_(t, 3.402823466e+38f);
    const RGB tr{exp_c(-sig_t[0] * tt * len), exp_c(-sig_t[1] * tt * len), exp_c(-sig_t[2] * tt * len)};
    const RGB density = *sampled ? RGB{sig_t[0] * tr.r, sig_t[1] * tr.g, sig_t[2] * tr.b} : tr;
    Float pdf = 0.0f;
    pdf += density.r; pdf += density.g; pdf += density.b;
    pdf *= 1.0f / 3.0f;
    if (pdf == 0.0f) pdf = 1.0f;
    return *sampled ? (tr * RGB{m.sigma_s[0], m.sigma_s[1], m.sigma_s[2]}) / pdf : tr / pdf;
}

// The scattering point handed to uniform_sample_one_light: a surface (bsdf != null) or a medium interaction (phase g).
struct VolVertex {
    V3 p, error, n, wo;
    const BSDF* bsdf;       // null: medium interaction
    Float g;
    int med_inside, med_outside;                                // MediumInterface of the interaction
    int medium_towards(V3 w) const { return dot(w, n) > 0.0f ? med_outside : med_inside; }       // D69
};
// primitive.rs:72-76: a hit takes the primitive's interface when that is a transition, else the ray's medium on both sides
inline void hit_interface(const Scene& scene, uint32_t prim, int ray_medium, int* inside, int* outside) {
    const int pi = scene.prim_inside[prim], po = scene.prim_outside[prim];
    if (pi != po) { *inside = pi; *outside = po; }
    else { *inside = ray_medium; *outside = ray_medium; }
}
// VisibilityTester::tr (light.rs:137-160, D74): transmittance from `from` towards the point (p1, p1_err, p1_n); black when a
// surface with a material is in the way
inline RGB visibility_tr(const Scene& scene, const VolVertex& from, V3 p1, V3 p1_err, V3 p1_n) {
    Interaction cur{from.p, from.error, from.n};
    int cur_in = from.med_inside, cur_out = from.med_outside;
    RGB tr = rgb(1.0f);
    for (;;) {
        const V3 origin = offset_ray_origin(cur.p, cur.error, cur.n, p1 - cur.p);          // spawn_ray_to(&BaseInteraction), interaction.rs:146-153
        const V3 target = offset_ray_origin(p1, p1_err, p1_n, origin - p1);
        Ray ray{origin, 1.0f - kShadowEpsilon, target - origin, 0.0f};
        const int medium = dot(ray.d, cur.n) > 0.0f ? cur_out : cur_in;
        SurfaceInteraction isect;
        PathCounters* pc = tl_path_counters;
        tl_ray_kind = 1;
        const bool hit = scene.intersect(ray, &isect);
        tl_ray_kind = 0;
        (void)pc;
        if (hit && scene.tri_material[isect.prim] != kNoMaterial) return rgb(0);
        if (medium >= 0) tr = tr * medium_tr(scene.media[medium], ray.t_max, ray.d);
        if (!hit) break;
        hit_interface(scene, isect.prim, medium, &cur_in, &cur_out);
        cur = Interaction{isect.p, isect.error, isect.n};
    }
    return tr;
}
// Scene::intersect_tr (scene.rs:48-71, D74)
inline bool intersect_tr(const Scene& scene, Ray ray, int medium, SurfaceInteraction* isect, RGB* tr) {
    *tr = rgb(1.0f);
    for (;;) {
        tl_ray_kind = 2;
        const bool hit = scene.intersect(ray, isect);
        tl_ray_kind = 0;
        if (medium >= 0) *tr = *tr * medium_tr(scene.media[medium], ray.t_max, ray.d);
        if (!hit) return false;
        if (scene.tri_material[isect->prim] != kNoMaterial) return true;
        int in, out;
        hit_interface(scene, isect->prim, medium, &in, &out);
        const V3 d = ray.d;
        ray = spawn_ray(Interaction{isect->p, isect->error, isect->n}, d);
        medium = dot(d, isect->n) > 0.0f ? out : in;
    }
}

// estimate_direct (integrator.rs:136-266) with handle_media = true, for a surface or a medium interaction
inline RGB estimate_direct_vol(const Scene& scene, const VolVertex& it, Float us0, Float us1, const LightRt& light, uint32_t light_index, Float ul0,
                               Float ul1) {
    const uint8_t flags = BSDF_ALL & ~BSDF_SPECULAR;
    RGB ld = rgb(0);
    V3 wi{0, 0, 0};
    Float light_pdf = 0.0f, scattering_pdf = 0.0f;
    RGB li = rgb(0);
    V3 p1{0, 0, 0}, p1_err{0, 0, 0}, p1_n{0, 0, 0};
    // ---- Light::sample_li (as in estimate_direct above) ----
    if (light.is_delta()) {
        V3 pl{light.d.p[0], light.d.p[1], light.d.p[2]};
        light_pdf = 1.0f;
        if (light.d.type == LIGHT_DISTANT) {
            wi = light.w_light;
            pl = it.p + light.w_light * (2.0f * light.world_radius);
            li = light.l();
        } else {
            wi = normalize(pl - it.p);
            if (light.d.type == LIGHT_SPOT) li = light.l() * light.falloff(-wi) / length_squared(pl - it.p);
            else li = light.l() / length_squared(pl - it.p);
        }
        p1 = pl;
    } else {
        V3 ps, pe, ns;
        Float pdf;
        if (light.sphere) light.sphere->sample2(it.p, it.error, it.n, ul0, ul1, &ps, &pe, &ns, &pdf);
        else {
            Float su0 = std::sqrt(ul0);
            Float b0 = 1.0f - su0, b1 = ul1 * su0;
            ps = (light.p0 * b0 + light.p1 * b1) + light.p2 * ((1.0f - b0) - b1);
            ns = normalize(cross(light.p1 - light.p0, light.p2 - light.p0));
            if (light.has_n) ns = faceforward(ns, (light.n0 * b0 + light.n1 * b1) + light.n2 * ((1.0f - b0) - b1));
            pe = ((vabs(light.p0 * b0) + vabs(light.p1 * b1)) + vabs(light.p2 * ((1.0f - b0) - b1))) * gamma(6.0f);
            pdf = 1.0f / light.area;
            V3 w = ps - it.p;
            if (length_squared(w) == 0.0f) pdf = 0.0f;
            else {
                w = normalize(w);
                pdf *= length_squared(it.p - ps) / std::fabs(dot(ns, -w));
                if (std::isinf(pdf)) pdf = 0.0f;
            }
        }
        light_pdf = pdf;
        if (pdf == 0.0f || length_squared(ps - it.p) == 0.0f) { light_pdf = 0.0f; li = rgb(0); }
        else {
            wi = normalize(ps - it.p);
            li = (light.d.two_sided || dot(ns, -wi) > 0.0f) ? light.l() : rgb(0);
            p1 = ps; p1_err = pe; p1_n = ns;
        }
    }
    if (light_pdf > 0.0f && !is_black(li)) {
        RGB f;
        if (it.bsdf) {
            scattering_pdf = it.bsdf->pdf(it.wo, wi, flags);
            f = it.bsdf->f(it.wo, wi, flags) * std::fabs(dot(wi, it.bsdf->ns));
        } else {
            const Float p = phase_hg(dot(it.wo, wi), it.g);
            scattering_pdf = p;
            f = rgb(p);
        }
        if (!is_black(f)) {
            if (tl_path_counters) tl_path_counters->rays[1]++;
            li = li * visibility_tr(scene, it, p1, p1_err, p1_n);
            if (!is_black(li)) {
                if (light.is_delta()) ld = ld + li * f / light_pdf;
                else ld = ld + li * f * power_heuristic(light_pdf, scattering_pdf) / light_pdf;
            }
        }
    }
    if (!light.is_delta()) {
        bool sampled_specular = false;
        RGB f;
        if (it.bsdf) {
            uint8_t sampled = 0;
            f = it.bsdf->sample_f(it.wo, &wi, us0, us1, &scattering_pdf, flags, &sampled);
            f = f * std::fabs(dot(wi, it.bsdf->ns));
            sampled_specular = (sampled & BSDF_SPECULAR) != 0;
        } else {
            const Float p = hg_sample_p(it.g, it.wo, &wi, us0, us1);
            scattering_pdf = p;
            f = rgb(p);
        }
        if (!is_black(f) && scattering_pdf > 0.0f) {
            Float weight = 1.0f;
            Interaction base{it.p, it.error, it.n};
            if (!sampled_specular) {
                if (light.sphere) light_pdf = light.sphere->pdf2(it.p, it.error, it.n, wi);
                else {
                    Ray r = spawn_ray(base, wi);
                    TriHit th = triangle_intersect_test(light.p0, light.p1, light.p2, r);
                    V3 du, dv;
                    if (!th.hit || !triangle_frame(light.p0, light.p1, light.p2, &du, &dv, light.has_uv ? light.uv : nullptr)) return ld;
                    Interaction li_it = triangle_interaction(light.p0, light.p1, light.p2, th.b0, th.b1, th.b2);
                    Float lp = length_squared(it.p - li_it.p) / (std::fabs(dot(li_it.n, -wi)) * light.area);
                    if (std::isinf(lp)) lp = 0.0f;
                    light_pdf = lp;
                }
                if (light_pdf == 0.0f) return ld;
                weight = power_heuristic(scattering_pdf, light_pdf);
            }
            Ray ray = spawn_ray(base, wi);
            SurfaceInteraction light_isect;
            RGB tr;
            RGB lmis = rgb(0);
            if (intersect_tr(scene, ray, it.medium_towards(wi), &light_isect, &tr)) {
                if (scene.tri_light[light_isect.prim] == (int32_t)light_index) lmis = scene.le(light_isect, -wi);
            }
            if (!is_black(lmis)) ld = ld + lmis * f * tr * weight / scattering_pdf;
        }
    }
    return ld;
}
inline RGB uniform_sample_one_light_vol(const Scene& scene, const VolVertex& it, Sampler& s) {
    if (scene.lights.empty()) return rgb(0);
    Float light_pdf;
    size_t num = scene.lookup(it.p).sample_discrete(s.get_1d(), &light_pdf);
    if (light_pdf == 0.0f) return rgb(0);
    Float ul0, ul1, us0, us1;
    s.get_2d(&ul0, &ul1);
    s.get_2d(&us0, &us1);
    return estimate_direct_vol(scene, it, us0, us1, scene.lights[num], (uint32_t)num, ul0, ul1) / light_pdf;
}

// volpath.rs:60-244 VolPathIntegrator::li (BSSRDF branch dead: no subsurface material exists)
inline RGB volpath_li(const Scene& scene, Ray ray, int ray_medium, Sampler& s, int max_depth, Float rr_threshold) {
    RGB l = rgb(0), beta = rgb(1);
    bool specular_bounce = false;
    int bounces = 0;
    Float eta_scale = 1.0f;
    if (tl_path_counters) tl_path_counters->camera_samples++;
    for (;;) {
        SurfaceInteraction isect;
        const V3 o0 = ray.o;
        (void)o0;
        bool found = scene.intersect(ray, &isect);
        bool in_medium = false;
        V3 mp{0, 0, 0};
        if (ray_medium >= 0) beta = beta * medium_sample(scene.media[ray_medium], ray, s, &in_medium, &mp);
        if (is_black(beta)) break;
        if (in_medium) {
            if (bounces >= max_depth) break;
            if (tl_path_counters) tl_path_counters->vertices++;
            const MediumDesc& m = scene.media[ray_medium];
            const V3 wo = -ray.d;
            V3 wi{0, 0, 0};
            Float u0, u1;
            s.get_2d(&u0, &u1);
            hg_sample_p(m.g, wo, &wi, u0, u1);                                                  // KEEP: sampled before the light
            VolVertex v{mp, V3{0, 0, 0}, V3{0, 0, 0}, wo, nullptr, m.g, ray_medium, ray_medium};
            ray = Ray{mp, kInfinity, wi, 0.0f};                                                 // mi.spawn_ray(wi): n = 0, no offset; same medium
            specular_bounce = false;
            l = l + beta * uniform_sample_one_light_vol(scene, v, s);
        } else {
            if (bounces == 0 || specular_bounce)
                if (found) l = l + beta * scene.le(isect, -ray.d);
            if (!found || bounces >= max_depth) break;
            int in, out;
            hit_interface(scene, isect.prim, ray_medium, &in, &out);
            if (scene.tri_material[isect.prim] == kNoMaterial) {                                // volpath.rs:127-131 (D70)
                const V3 d = ray.d;
                ray = spawn_ray(Interaction{isect.p, isect.error, isect.n}, d);
                ray_medium = dot(d, isect.n) > 0.0f ? out : in;
                continue;
            }
            if (tl_path_counters) tl_path_counters->vertices++;
            BSDF bsdf = scene.make_bsdf(isect);
            VolVertex v{isect.p, isect.error, isect.n, isect.wo, &bsdf, 0.0f, in, out};
            l = l + beta * uniform_sample_one_light_vol(scene, v, s);
            V3 wo = -ray.d, wi{0, 0, 0};
            Float pdf = 0.0f, u0, u1;
            uint8_t flags = 0;
            s.get_2d(&u0, &u1);
            RGB f = bsdf.sample_f(wo, &wi, u0, u1, &pdf, BSDF_ALL, &flags);
            if (is_black(f) || pdf == 0.0f) break;
            beta = beta * (f * (std::fabs(dot(wi, bsdf.ns)) / pdf));
            specular_bounce = (flags & BSDF_SPECULAR) != 0;
            if ((flags & BSDF_SPECULAR) && (flags & BSDF_TRANSMISSION)) {
                Float eta = bsdf.eta;
                eta_scale *= (dot(wo, isect.n) > 0.0f) ? (eta * eta) : 1.0f / (eta * eta);
            }
            ray = spawn_ray(Interaction{isect.p, isect.error, isect.n}, wi);
            ray_medium = v.medium_towards(wi);
        }
        RGB rr_beta = beta * eta_scale;
        if (max_component_value(rr_beta) < rr_threshold && bounces > 3) {
            Float q = fmax_(1.0f - max_component_value(rr_beta), 0.05f);                        // volpath.rs:236
            if (s.get_1d() < q) break;
            beta = beta / (1.0f - q);
        }
        bounces += 1;
    }
    return l;
}

// ---------------------------------------------------------------- film.rs
struct Film {
    FilmDesc d;
    Float table[16 * 16];
    int px0, py0, px1, py1;                  // cropped_pixel_bounds (film.rs:41-50)
    int sb_x0, sb_y0, sb_x1, sb_y1;          // sample bounds (D42 FIX)
    Float max_lum;
    void init(const FilmDesc& desc) {
        d = desc;
        const Float* cw = d.crop_window;
        const bool full = cw[0] == 0.0f && cw[1] == 0.0f && cw[2] == 0.0f && cw[3] == 0.0f;
        const Float c0x = full ? 0.0f : cw[0], c0y = full ? 0.0f : cw[1], c1x = full ? 1.0f : cw[2], c1y = full ? 1.0f : cw[3];
        px0 = (int)std::ceil((Float)d.res_x * c0x);                                            // film.rs:41-50
        py0 = (int)std::ceil((Float)d.res_y * c0y);
        px1 = (int)std::ceil((Float)d.res_x * c1x);
        py1 = (int)std::ceil((Float)d.res_y * c1y);
        max_lum = d.max_sample_luminance > 0.0f ? d.max_sample_luminance : std::numeric_limits<Float>::infinity();
        for (int y = 0; y < 16; ++y)
            for (int x = 0; x < 16; ++x) {
                Float px = ((Float)x + 0.5f) * d.radius_x / 16.0f, py = ((Float)y + 0.5f) * d.radius_y / 16.0f;   // film.rs:53-63
                table[y * 16 + x] = evaluate(px, py);
            }
        // Film::get_sample_bounds (film.rs:76-81; D42 FIX: pbrt-v3 floor(min + 0.5 - r), ceil(max - 0.5 + r))
        sb_x0 = (int)std::floor((Float)px0 + 0.5f - d.radius_x);
        sb_y0 = (int)std::floor((Float)py0 + 0.5f - d.radius_y);
        sb_x1 = (int)std::ceil((Float)px1 - 0.5f + d.radius_x);
        sb_y1 = (int)std::ceil((Float)py1 - 0.5f + d.radius_y);
    }
    int width() const { return px1 - px0; }
    int height() const { return py1 - py0; }
    size_t index(int x, int y) const { return (size_t)(y - py0) * (size_t)width() + (size_t)(x - px0); }
    Float gaussian(Float v, Float expv) const { return fmax_(std::exp(-d.gaussian_alpha * v * v) - expv, 0.0f); }   // gaussian.rs:29-31
    Float mitchell_1d(Float x) const {                                                         // mitchell.rs:24-38
        const Float B = d.mitchell_b, C = d.mitchell_c;
        x = std::fabs(2.0f * x);
        if (x > 1.0f)
            return ((-B - 6.0f * C) * x * x * x + (6.0f * B + 30.0f * C) * x * x + (-12.0f * B - 48.0f * C) * x + (8.0f * B + 24.0f * C)) * (1.0f / 6.0f);
        return ((12.0f - 9.0f * B - 6.0f * C) * x * x * x + (-18.0f + 12.0f * B + 6.0f * C) * x * x + (6.0f - 2.0f * B)) * (1.0f / 6.0f);
    }
    static Float sinc(Float x) {                                                               // sinc.rs:22-29
        x = std::fabs(x);
        if (x < 1e-5f) return 1.0f;
        return std::sin(kPi * x) / (kPi * x);
    }
    Float windowed_sinc(Float x, Float radius) const {                                         // sinc.rs:30-38
        x = std::fabs(x);
        if (x > radius) return 0.0f;
        const Float lanczos = sinc(x / d.sinc_tau);
        return sinc(x) * lanczos;
    }
    Float evaluate(Float x, Float y) const {
        if (d.filter == FILTER_BOX) return 1.0f;                                               // boxf.rs:26-28
        if (d.filter == FILTER_TRIANGLE) return fmax_(d.radius_x - std::fabs(x), 0.0f) * fmax_(d.radius_y - std::fabs(y), 0.0f);   // triangle.rs:20-22
        if (d.filter == FILTER_MITCHELL) return mitchell_1d(x * (1.0f / d.radius_x)) * mitchell_1d(y * (1.0f / d.radius_y));     // mitchell.rs:42-45
        if (d.filter == FILTER_SINC) return windowed_sinc(x, d.radius_x) * windowed_sinc(y, d.radius_y);                         // sinc.rs:42-44
        Float ex = std::exp(-d.gaussian_alpha * d.radius_x * d.radius_x), ey = std::exp(-d.gaussian_alpha * d.radius_y * d.radius_y);
        return gaussian(x, ex) * gaussian(y, ey);
    }
    // FilmTile::add_sample (film.rs:259-261): luminance clamp
    RGB clamp_luminance(RGB l) const {
        const Float y = y_value(l);
        if (y > max_lum) l = l * (max_lum / y);
        return l;
    }
    // film.rs:252-295 FilmTile::add_sample footprint + weights (D43, D44 FIX); calls fn(px, py, filter_weight)
    template <class F>
    void footprint(Float pfx, Float pfy, F&& fn) const {
        Float dx = pfx - 0.5f, dy = pfy - 0.5f;
        int x0 = (int)std::ceil(dx - d.radius_x), y0 = (int)std::ceil(dy - d.radius_y);
        int x1 = (int)std::floor(dx + d.radius_x) + 1, y1 = (int)std::floor(dy + d.radius_y) + 1;
        x0 = std::max(x0, px0); y0 = std::max(y0, py0);
        x1 = std::min(x1, px1); y1 = std::min(y1, py1);
        for (int y = y0; y < y1; ++y) {
            Float fy = std::fabs(((Float)y - dy) * (1.0f / d.radius_y) * 16.0f);
            int iy = std::min(15, (int)std::floor(fy));
            for (int x = x0; x < x1; ++x) {
                Float fx = std::fabs(((Float)x - dx) * (1.0f / d.radius_x) * 16.0f);
                int ix = std::min(15, (int)std::floor(fx));
                fn(x, y, table[iy * 16 + ix]);
            }
        }
    }
};

struct Stray {
    uint64_t order;       // (source sample-bounds pixel index) * spp + sample
    int x, y;
    RGB c;
    Float w;
};

// SamplerIntegrator::render (integrator.rs:399-480).
//  mode 1 = per-(pixel,sample) sampler streams: RNG::new(((y-sb_y0)*W_sb + (x-sb_x0))*spp + s); film accumulation order
//           per target pixel: its own samples in sample order, then contributions from other pixels' samples ordered
//           by (source pixel, sample).  This is the mode the GPU implements.
//  mode 0 = the reference's order: one sampler stream per 16x16 tile (seed = tile.y*n_tiles.x + tile.x), consumed
//           sequentially over pixels, samples and path vertices.
// out_xyzw: float4 {X, Y, Z, filter_weight_sum} per pixel, ADDED to the buffer.  Returns seconds.
inline double render(const Scene& scene, const CameraDesc& cd, const FilmDesc& fd, const PathDesc& pd, int mode, int threads, Float* out_xyzw,
                     PathCounters* counters_out = nullptr) {
    Camera cam;
    cam.init({cd.pos[0], cd.pos[1], cd.pos[2]}, {cd.look[0], cd.look[1], cd.look[2]}, {cd.up[0], cd.up[1], cd.up[2]}, cd.fov, cd.res_x, cd.res_y);
    cam.lens_radius = cd.lens_radius;
    cam.focal_distance = cd.focal_distance;
    Film film;
    film.init(fd);
    const int W = film.sb_x1 - film.sb_x0, H = film.sb_y1 - film.sb_y0;
    const int tiles_x = (W + 15) / 16, tiles_y = (H + 15) / 16;
    const size_t npix = (size_t)film.width() * film.height();  // cropped_pixel_bounds.area() (film.rs:51)
    HaltonTables halton;
    if (pd.sampler == 1) halton.init(W, H);                   // sample_bounds extent (halton.rs:69)
    SobolTables sobol;
    if (pd.sampler == 4) {
        if (sobol_tables_base()) sobol = *sobol_tables_base();
        sobol.init(film.sb_x0, film.sb_y0, W, H);
    }
    std::vector<RGB> acc(npix, rgb(0));
    std::vector<Float> wsum(npix, 0.0f);
    std::vector<std::vector<Stray>> strays(threads);
    std::atomic<int> next{0};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<PathCounters> thread_counters(threads);
    auto work = [&](int tid) {
        tl_path_counters = counters_out ? &thread_counters[tid] : nullptr;
        tl_ray_kind = 0;
        for (;;) {
            int t = next.fetch_add(1);
            if (t >= tiles_x * tiles_y) { tl_path_counters = nullptr; break; }
            int tx = t % tiles_x, ty = t / tiles_x;
            Sampler tile_sampler;
            tile_sampler.kind = pd.sampler;
            tile_sampler.halton = &halton;
            tile_sampler.sobol = &sobol;
            tile_sampler.rng.set_sequence((uint64_t)(ty * tiles_x + tx));                       // integrator.rs:414-415
            int x0 = film.sb_x0 + tx * 16, x1 = std::min(x0 + 16, film.sb_x1);
            int y0 = film.sb_y0 + ty * 16, y1 = std::min(y0 + 16, film.sb_y1);
            // mode 0 accumulates into a tile-local buffer first (FilmTile), merged below
            PixelTables tabs;
            if (pd.sampler == 2 || pd.sampler == 3) tabs.resize(pd.n_sampled_dimensions, pd.spp);
            tile_sampler.tabs = &tabs;
            for (int y = y0; y < y1; ++y)
                for (int x = x0; x < x1; ++x) {
                    if (pd.sampler == 2 || pd.sampler == 3) {                                   // Sampler::start_pixel (integrator.rs:433)
                        // mode 1: the tables of a pixel come from their own stream, RNG::new(W*H*spp + pixel index)
                        RNG table_rng;
                        table_rng.set_sequence((uint64_t)W * H * (uint64_t)pd.spp + ((uint64_t)(y - film.sb_y0) * W + (uint64_t)(x - film.sb_x0)));
                        tabs.start_pixel(pd.sampler, mode == 1 ? table_rng : tile_sampler.rng, pd);
                    }
                    for (int s = pd.sample_begin; s < pd.sample_end; ++s) {
                        Sampler own;
                        own.kind = pd.sampler;
                        own.halton = &halton;
                        own.sobol = &sobol;
                        own.tabs = &tabs;
                        uint64_t order = ((uint64_t)(y - film.sb_y0) * W + (uint64_t)(x - film.sb_x0)) * (uint64_t)pd.spp + (uint64_t)s;
                        if (mode == 1) own.rng.set_sequence(order);
                        Sampler& smp = mode == 1 ? own : tile_sampler;
                        smp.start_sample(x, y, (uint64_t)s);
                        Float u0, u1, ut, l0, l1;
                        smp.get_2d(&u0, &u1);                                                   // sampler.rs:27-33
                        ut = smp.get_1d();
                        smp.get_2d(&l0, &l1);
                        (void)ut;
                        Float pfx = (Float)x + u0, pfy = (Float)y + u1;
                        Ray ray = cam.generate_ray(pfx, pfy, l0, l1);
                        RGB L = pd.integrator == 1 ? volpath_li(scene, ray, scene.camera_medium, smp, pd.max_depth, pd.rr_threshold)
                                                   : path_li(scene, ray, smp, pd.max_depth, pd.rr_threshold);
                        if (has_nans(L) || y_value(L) < -1e-5f || std::isinf(y_value(L))) L = rgb(0);   // D22 FIX
                        L = film.clamp_luminance(L);
                        film.footprint(pfx, pfy, [&](int px, int py, Float fw) {
                            RGB c = L * 1.0f * fw;                                              // l * sample_weight * filter_weight
                            if (mode == 1 && !(px == x && py == y)) { strays[tid].push_back(Stray{order, px, py, c, fw}); return; }
                            if (mode == 0 && (px < x0 || px >= x1 || py < y0 || py >= y1)) { strays[tid].push_back(Stray{order, px, py, c, fw}); return; }
                            size_t o = film.index(px, py);
                            acc[o] = acc[o] + c;
                            wsum[o] += fw;
                        });
                    }
                }
        }
    };
    std::vector<std::thread> pool;
    for (int i = 1; i < threads; ++i) pool.emplace_back(work, i);
    work(0);
    for (auto& t : pool) t.join();
    if (counters_out)
        for (const PathCounters& c : thread_counters) {
            counters_out->camera_samples += c.camera_samples;
            counters_out->vertices += c.vertices;
            for (int k = 0; k < 3; ++k) {
                counters_out->rays[k] += c.rays[k];
                counters_out->trav[k].nodes_tested += c.trav[k].nodes_tested;
                counters_out->trav[k].tris_tested += c.trav[k].tris_tested;
            }
        }
    std::vector<Stray> all;
    for (auto& v : strays) all.insert(all.end(), v.begin(), v.end());
    std::sort(all.begin(), all.end(), [](const Stray& a, const Stray& b) {
        if (a.order != b.order) return a.order < b.order;
        if (a.y != b.y) return a.y < b.y;
        return a.x < b.x;
    });
    for (const Stray& s : all) {
        size_t o = film.index(s.x, s.y);
        acc[o] = acc[o] + s.c;
        wsum[o] += s.w;
    }
    for (size_t i = 0; i < npix; ++i) {                                                        // Film::merge_film_tile :111-123
        Float xyz[3];
        rgb_to_xyz(acc[i], xyz);
        out_xyzw[4 * i] += xyz[0]; out_xyzw[4 * i + 1] += xyz[1]; out_xyzw[4 * i + 2] += xyz[2];
        out_xyzw[4 * i + 3] += wsum[i];
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// Film::write_image (film.rs:153-178): XYZ -> RGB, / weight, clamp >= 0, * scale
inline void resolve_rgb(const Float* xyzw, size_t npix, Float scale, Float* out_rgb) {
    for (size_t i = 0; i < npix; ++i) {
        Float c[3];
        xyz_to_rgb(xyzw + 4 * i, c);
        Float w = xyzw[4 * i + 3];
        if (w != 0.0f) {
            Float inv = 1.0f / w;
            c[0] = fmax_(c[0] * inv, 0.0f); c[1] = fmax_(c[1] * inv, 0.0f); c[2] = fmax_(c[2] * inv, 0.0f);
        }
        out_rgb[3 * i] = c[0] * scale; out_rgb[3 * i + 1] = c[1] * scale; out_rgb[3 * i + 2] = c[2] * scale;
    }
}

// Film::add_splat (film.rs:137-151) + the splat term of Film::write_image (:167-172).
//   D64 FIX  :139-141 returns when the pixel IS inside cropped_pixel_bounds (and tests the inclusive box) -> pbrt-v3: skip
//            pixels outside, upper bound exclusive
inline void film_add_splats(const Film& film, const Float* p_film, const Float* v_rgb, size_t n, Float* splat_xyz /* 3 per pixel */) {
    for (size_t i = 0; i < n; ++i) {
        const Float fx = std::floor(p_film[2 * i]), fy = std::floor(p_film[2 * i + 1]);
        if (!(fx >= (Float)film.px0 && fx < (Float)film.px1 && fy >= (Float)film.py0 && fy < (Float)film.py1)) continue;
        const RGB v = film.clamp_luminance(RGB{v_rgb[3 * i], v_rgb[3 * i + 1], v_rgb[3 * i + 2]});
        Float xyz[3];
        rgb_to_xyz(v, xyz);
        Float* dst = splat_xyz + 3 * film.index((int)fx, (int)fy);
        dst[0] += xyz[0]; dst[1] += xyz[1]; dst[2] += xyz[2];
    }
}
inline void resolve_rgb_splat(const Float* xyzw, const Float* splat_xyz, size_t npix, Float scale, Float splat_scale, Float* out_rgb) {
    for (size_t i = 0; i < npix; ++i) {
        Float c[3], sc[3];
        xyz_to_rgb(xyzw + 4 * i, c);
        Float w = xyzw[4 * i + 3];
        if (w != 0.0f) {
            Float inv = 1.0f / w;
            c[0] = fmax_(c[0] * inv, 0.0f); c[1] = fmax_(c[1] * inv, 0.0f); c[2] = fmax_(c[2] * inv, 0.0f);
        }
        xyz_to_rgb(splat_xyz + 3 * i, sc);
        for (int k = 0; k < 3; ++k) out_rgb[3 * i + k] = (c[k] + splat_scale * sc[k]) * scale;
    }
}

}  // namespace orc
