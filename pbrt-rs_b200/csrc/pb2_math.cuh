// pb2_math.cuh — numerics contract of the backend (host + device).
//
// Float = f32.  Every + - * / below is one separately rounded IEEE binary32 operation, exactly as rustc emits
// them for the reference (no FMA contraction, no fast-math).  This file is compiled with
//   nvcc -fmad=false  (device)   and   -Xcompiler -ffp-contract=off  (host)
// and never with --use_fast_math; -prec-div / -prec-sqrt / -ftz keep their IEEE defaults.
// Operation order follows SURVEY.md Appendix D (reference file:line cited per function).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define PB2_HD __host__ __device__ __forceinline__
#define PB2_D __device__ __forceinline__
#else
#define PB2_HD inline
#define PB2_D inline
#endif

namespace pb2 {

struct vec3 {
    float x, y, z;
};

PB2_HD vec3 mk(float x, float y, float z) { vec3 r; r.x = x; r.y = y; r.z = z; return r; }
PB2_HD vec3 operator+(vec3 a, vec3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
PB2_HD vec3 operator-(vec3 a, vec3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
PB2_HD vec3 operator-(vec3 a) { return mk(-a.x, -a.y, -a.z); }
PB2_HD vec3 operator*(vec3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
PB2_HD vec3 operator*(vec3 a, vec3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
PB2_HD vec3 operator/(vec3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }   // geometry.rs:268-276: 3 divides
// geometry.rs:228-234: (x*x' + y*y') + z*z'
PB2_HD float dot3(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
PB2_HD float len2(vec3 a) { return (a.x * a.x + a.y * a.y) + a.z * a.z; }
PB2_HD float len(vec3 a) { return sqrtf(len2(a)); }
PB2_HD vec3 unit(vec3 a) { return a / len(a); }                                       // geometry.rs:117-119
PB2_HD vec3 abs3(vec3 a) { return mk(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
PB2_HD vec3 cross3(vec3 a, vec3 b) {                                                  // geometry.rs:361-373
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
PB2_HD float comp(vec3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
// geometry.rs:91-93 (ties go to the higher axis)
PB2_HD int max_dim(vec3 a) { return (a.x > a.y && a.x > a.z) ? 0 : ((a.y > a.z) ? 1 : 2); }
// geometry.rs:95-97: x.max(y.max(z))
PB2_HD float max3(float x, float y, float z) { return fmaxf(x, fmaxf(y, z)); }

// pbrt.rs:26-28, :89-91
#define PB2_MACHINE_EPS 0x1p-24f
#define PB2_ONE_MINUS_EPS 0x1.fffffcp-1f   /* 1 - f32::EPSILON as written in the reference (pbrt.rs:28) */
#define PB2_SHADOW_EPS 0.0001f
#define PB2_PI 3.14159265358979323846f
PB2_HD float gammaf_(float n) { return (n * PB2_MACHINE_EPS) / (1.0f - n * PB2_MACHINE_EPS); }

PB2_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
PB2_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
// pbrt.rs:43-77
PB2_HD float next_up(float v) {
    if (isinf(v) && v > 0.0f) return v;
    if (v == -0.0f) v = 0.0f;
    uint32_t u = f2u(v);
    u = (v >= 0.0f) ? u + 1u : u - 1u;
    return u2f(u);
}
PB2_HD float next_down(float v) {
    if (isinf(v) && v < 0.0f) return v;
    if (v == 0.0f) v = -0.0f;
    uint32_t u = f2u(v);
    u = (v > 0.0f) ? u - 1u : u + 1u;
    return u2f(u);
}

// geometry.rs:375-383
PB2_HD void coord_system(vec3 v1, vec3* v2, vec3* v3) {
    if (fabsf(v1.x) > fabsf(v1.y)) *v2 = unit(mk(-v1.z, 0.0f, v1.x));
    else *v2 = unit(mk(0.0f, v1.z, -v1.y));
    *v3 = cross3(v1, *v2);
}

// off > 0: next_up(po); off < 0: next_down(po); else po (geometry.rs:1146-1152) — one branch-free form of the two functions
// above: a step away from zero is bits + 1 (an infinity stays), a step towards zero is bits - 1, and either zero steps to the
// smallest denormal of the wanted sign; NaNs move as next_up / next_down move them.  Equal to the branching form for all 2^32
// values of po and both signs of off (tools/nudge_check.cpp, exhaustive; tests/test_host_side.py runs it on every 251st value).  k_shade spent 12 % of its warp instructions in the branching form.
PB2_HD float nudge(float po, float off) {
    const bool up = off > 0.0f, dn = off < 0.0f;
    const uint32_t u = f2u(po);
    const bool away = (po > 0.0f) == up;
    const bool is_inf = (u & 0x7fffffffu) == 0x7f800000u;
    uint32_t r = away ? (is_inf ? u : u + 1u) : u - 1u;
    if (po == 0.0f) r = up ? 1u : 0x80000001u;
    return (up || dn) ? u2f(r) : po;
}
// geometry.rs:1139-1154 offset_ray_origin
PB2_HD vec3 offset_ray_origin(vec3 p, vec3 p_error, vec3 n, vec3 w) {
    float d = dot3(abs3(n), p_error);
    vec3 off = n * d;
    if (dot3(w, n) < 0.0f) off = -off;
    const vec3 po = p + off;
    return mk(nudge(po.x, off.x), nudge(po.y, off.y), nudge(po.z, off.z));
}

// src/core/rng.rs:14-48 PCG32 (wrapping u64)
struct Pcg32 {
    uint64_t state, inc;
    PB2_HD uint32_t next_u32() {
        uint64_t old = state;
        state = old * 0x5851f42d4c957f2dULL + inc;
        uint32_t xs = (uint32_t)(((old >> 18) ^ old) >> 27);
        uint32_t rot = (uint32_t)(old >> 59);
        return (xs >> rot) | (xs << ((~rot + 1u) & 31u));
    }
    PB2_HD void set_sequence(uint64_t seq) {
        state = 0;
        inc = (seq << 1) | 1ULL;
        next_u32();
        state += 0x853c49e6748fea9bULL;
        next_u32();
    }
    // rng.rs:46-48: min(ONE_MINUS_EPSILON, (f32)u32 * 2^-32); u32 -> f32 is round-to-nearest-even
    PB2_HD float next_float() {
#if defined(__CUDA_ARCH__)
        float f = __uint2float_rn(next_u32());
#else
        float f = (float)next_u32();
#endif
        return fminf(PB2_ONE_MINUS_EPS, f * 2.3283064365386963e-10f);
    }
};

// Deterministic sin/cos used wherever the reference calls f32::sin / f32::cos (sampling.rs:258-273,
// microfacet.rs:336-384).  Rust forwards those to the platform libm, whose last-bit results differ between
// platforms; the backend fixes one definition so host and device agree bit for bit: Cody-Waite reduction by pi/2
// in three f32 steps, then the Cephes single-precision minimax polynomials on [-pi/4, pi/4] (<= 2 ulp for |x| < 100).
PB2_HD void det_sincos(float x, float* s_out, float* c_out) {
    const float q = rintf(x * 0.636619772367581343f);
    const int k = (int)q;
    float r = x - q * 1.5703125f;
    r = r - q * 4.837512969970703125e-4f;
    r = r - q * 7.549789948768648e-8f;
    const float z = r * r;
    const float sp = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * r + r;
    const float cp = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z - 0.5f * z + 1.0f;
    float s, c;
    switch (k & 3) {
        case 0: s = sp; c = cp; break;
        case 1: s = cp; c = -sp; break;
        case 2: s = -sp; c = -cp; break;
        default: s = -cp; c = sp; break;
    }
    *s_out = s;
    *c_out = c;
}
PB2_HD float det_sin(float x) { float s, c; det_sincos(x, &s, &c); return s; }
PB2_HD float det_cos(float x) { float s, c; det_sincos(x, &s, &c); return c; }

// Deterministic exp / ln for HomogeneousMedium (media/homogeneous.rs: f32::exp / f32::ln = platform libm in the reference):
// Cephes expf / logf, every operation one rounded f32 op; exp flushes results below the normal range to 0.
PB2_HD float det_exp(float x) {
    if (x > 88.0f) return u2f(0x7f800000u);
    if (x < -87.0f) return 0.0f;
    float z = floorf(1.44269504088896341f * x + 0.5f);
    x = x - z * 0.693359375f;
    x = x - z * -2.12194440e-4f;
    const int n = (int)z;
    z = x * x;
    z = (((((1.9875691500e-4f * x + 1.3981999507e-3f) * x + 8.3334519073e-3f) * x + 4.1665795894e-2f) * x + 1.6666665459e-1f) * x + 5.0000001201e-1f) * z + x + 1.0f;
    return z * u2f((uint32_t)(n + 127) << 23);
}
PB2_HD float det_log(float x) {
    if (x <= 0.0f) return -u2f(0x7f800000u);
    const uint32_t u = f2u(x);
    int e = (int)(u >> 23) - 126;
    float m = u2f((u & 0x007FFFFFu) | 0x3F000000u);
    if (m < 0.707106781186547524f) { e -= 1; m = (m + m) - 1.0f; }
    else m = m - 1.0f;
    const float z = m * m;
    float y = ((((((((7.0376836292e-2f * m - 1.1514610310e-1f) * m + 1.1676998740e-1f) * m - 1.2420140846e-1f) * m + 1.4249322787e-1f) * m - 1.6668057665e-1f) * m +
                 2.0000714765e-1f) * m - 2.4999993993e-1f) * m + 3.3333331174e-1f) * m * z;
    const float fe = (float)e;
    y = y + -2.12194440e-4f * fe;
    y = y + -0.5f * z;
    float r = m + y;
    r = r + 0.693359375f * fe;
    return r;
}

struct mat4 {
    float m[4][4];
};

}  // namespace pb2
