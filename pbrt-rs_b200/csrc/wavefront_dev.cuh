// wavefront_dev.cuh — device code shared by the translation units of the wavefront PathIntegrator (wavefront.cu: queues,
// ray generation, extend / shadow, film; wavefront_shade.cu: k_shade, one object per <sampler tables, shading geometry>
// combination; wavefront_volpath.cu: k_volpath): slot <-> (pixel, sample), the sampler streams, the path vertex rebuilt from
// a hit record and estimate_direct.  One .cu held all of it at first and took six minutes to compile.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "camera.cuh"
#include "trace_persistent.cuh"
#include "wavefront.cuh"

namespace pb2 {

TraceTuning trace_tuning();

// k_shade<material class, PixelSampler tables, mesh shading geometry> for the three material queues of one bounce; each
// <TABLES, SG> combination is defined by its own object file (wavefront_shade.cu compiled four times).
template <bool TABLES, bool SG>
void launch_shade_t(Wavefront* wf, const SceneView& sv, const ShadeView& sh, const PathBuffers& b, const PathMap& map, const FilmView& film,
                    const PathParams& pp, int cur, uint64_t n, cudaStream_t st);
template <> void launch_shade_t<false, false>(Wavefront* wf, const SceneView& sv, const ShadeView& sh, const PathBuffers& b, const PathMap& map, const FilmView& film,
                    const PathParams& pp, int cur, uint64_t n, cudaStream_t st);
template <> void launch_shade_t<false, true>(Wavefront* wf, const SceneView& sv, const ShadeView& sh, const PathBuffers& b, const PathMap& map, const FilmView& film,
                    const PathParams& pp, int cur, uint64_t n, cudaStream_t st);
template <> void launch_shade_t<true, false>(Wavefront* wf, const SceneView& sv, const ShadeView& sh, const PathBuffers& b, const PathMap& map, const FilmView& film,
                    const PathParams& pp, int cur, uint64_t n, cudaStream_t st);
template <> void launch_shade_t<true, true>(Wavefront* wf, const SceneView& sv, const ShadeView& sh, const PathBuffers& b, const PathMap& map, const FilmView& film,
                    const PathParams& pp, int cur, uint64_t n, cudaStream_t st);
// VolPathIntegrator::li for n camera samples (wavefront_volpath.cu)
void launch_volpath(Wavefront* wf, unsigned grid, uint64_t n, const SceneView& sv, const ShadeView& sh, const PathBuffers& b, const PathMap& map,
                    const FilmView& film, const CameraView& cam, const PathParams& pp, cudaStream_t st);

// the wavefront VolPathIntegrator: all iterations of one batch of n path slots (wavefront_volpath.cu)
void trace_batch_vol(Wavefront* wf, const SceneView& sv, const ShadeView& sh, const CameraView& cam, const FilmView& film, const PathMap& map,
                     const PathParams& pp, uint64_t n, cudaStream_t st);
// wavefront.cu: the stable three-way queue select (k_select3) over `state` (null: PathBuffers::state) and k_raygen
void select_queues(Wavefront* wf, const uint32_t* in, int n_in, bool by_class, uint32_t* o0, int c0, uint32_t* o1, int c1, uint32_t* o2, int c2,
                   cudaStream_t st, const uint8_t* state = nullptr);
void launch_raygen(Wavefront* wf, uint64_t n, const PathMap& map, const FilmView& film, const CameraView& cam, cudaStream_t st);

namespace {

constexpr float kInf = __builtin_huge_valf();
constexpr int kThreads = 256;

// PathBuffers::state (see the order-preserving queues in wavefront.cu)
constexpr unsigned kStateDead = 3u;            // bits 0-1: shading class of the hit, 3 = the ray escaped
constexpr unsigned kStateContinues = 4u;       // bit 2: the path continues with the ray k_shade wrote
                                               // bits 3, 4: shadow ray / MIS ray of the NEE record pending

// ---- slot <-> (pixel, sample) -----------------------------------------------------------------------------------
struct SlotInfo {
    int x, y;                 // pixel in image coordinates (may lie outside the image for wide filters)
    uint32_t sample;          // sample index of the pixel
    unsigned long long seq;   // sampler stream: ((y - sb_y0) * sb_w + (x - sb_x0)) * spp + sample
    uint32_t pix;             // (y - sb_y0) * sb_w + (x - sb_x0)
};
// (32-bit arithmetic: a wavefront holds at most 2^28 slots, and a 64-bit modulo costs ~100 instructions in every kernel
// that resumes a sampler)
__device__ __forceinline__ SlotInfo slot_info(const PathMap& m, const FilmView& f, uint64_t slot64) {
    const uint32_t slot = (uint32_t)slot64;
    SlotInfo s;
    if (m.explicit_xy) {
        s.x = m.explicit_xy[2ull * slot];
        s.y = m.explicit_xy[2ull * slot + 1];
        s.sample = m.explicit_s[slot];
        s.pix = (uint32_t)(s.y - f.sb_y0) * (uint32_t)f.sb_w + (uint32_t)(s.x - f.sb_x0);
    } else {
        const uint32_t s_local = fast_div(slot, m.n_pix_magic);
        s.pix = slot - s_local * m.n_pix;
        s.sample = (uint32_t)m.sample0 + s_local;
        const uint32_t row = fast_div(s.pix, f.sb_w_magic);
        s.x = f.sb_x0 + (int)(s.pix - row * (uint32_t)f.sb_w);
        s.y = f.sb_y0 + (int)row;
    }
    s.seq = (unsigned long long)s.pix * (unsigned long long)m.spp + s.sample;
    return s;
}

// ---- samplers --------------------------------------------------------------------------------------------------------------
// HaltonSampler (samplers/halton.rs, core/lowdiscrepancy.rs:293-390; pbrt-v3 semantics where the port is broken).
__device__ __forceinline__ unsigned long long inverse_radical_inverse(unsigned base, unsigned long long inverse, int n_digits) {
    unsigned long long index = 0;
    for (int i = 0; i < n_digits; ++i) {
        const unsigned long long digit = inverse % base;
        inverse /= base;
        index = index * base + digit;
    }
    return index;
}
__device__ __forceinline__ long long halton_index(const SamplerView& h, int px, int py, unsigned long long sample_num) {    // halton.rs:117-141
    long long offset = 0;
    if (h.sample_stride > 1ull) {
        const int pm0 = ((px % 128) + 128) % 128, pm1 = ((py % 128) + 128) % 128;
        offset += (long long)(inverse_radical_inverse(2u, (unsigned long long)pm0, h.base_exponents[0]) *
                              (h.sample_stride / (unsigned long long)h.base_scales[0]) * h.mult_inverse[0]);
        offset += (long long)(inverse_radical_inverse(3u, (unsigned long long)pm1, h.base_exponents[1]) *
                              (h.sample_stride / (unsigned long long)h.base_scales[1]) * h.mult_inverse[1]);
        offset %= (long long)h.sample_stride;
    }
    return offset + (long long)(sample_num * h.sample_stride);
}
// Digit loop of radical_inverse_specialized / scramble_radical_inverse_specialized (lowdiscrepancy.rs:293-320): integer
// digits are exact, so a 32-bit index (every practical frame: index < stride * spp) takes 32-bit divisions.
template <class UInt, bool SCRAMBLED>
__device__ __forceinline__ void halton_digits(UInt a, UInt base, const uint16_t* perm, float inv_base, unsigned long long* reversed, float* inv_base_n) {
    while (a != 0) {
        const UInt next = a / base, digit = a - next * base;
        *reversed = *reversed * base + (SCRAMBLED ? (unsigned long long)perm[digit] : (unsigned long long)digit);
        *inv_base_n = *inv_base_n * inv_base;
        a = next;
    }
}
__device__ __forceinline__ float halton_dimension(const SamplerView& h, unsigned long long index, unsigned dim) {            // halton.rs:143-155
    if (dim == 0u) return __ull2float_rn(__brevll(index >> h.base_exponents[0])) * 5.4210108624275222e-20f;
    const unsigned long long a = dim == 1u ? index / (unsigned long long)h.base_scales[1] : index;
    const unsigned base = h.primes[dim];
    const float inv_base = 1.0f / (float)base;
    unsigned long long reversed = 0;
    float inv_base_n = 1.0f;
    if (dim == 1u) {                                                                       // radical_inverse
        if (a >> 32) halton_digits<unsigned long long, false>(a, base, nullptr, inv_base, &reversed, &inv_base_n);
        else halton_digits<unsigned, false>((unsigned)a, base, nullptr, inv_base, &reversed, &inv_base_n);
        return fminf(__ull2float_rn(reversed) * inv_base_n, PB2_ONE_MINUS_EPS);
    }
    const uint16_t* perm = h.perms + h.prime_sums[dim];                                    // scramble_radical_inverse
    if (a >> 32) halton_digits<unsigned long long, true>(a, base, perm, inv_base, &reversed, &inv_base_n);
    else halton_digits<unsigned, true>((unsigned)a, base, perm, inv_base, &reversed, &inv_base_n);
    return fminf(inv_base_n * (__ull2float_rn(reversed) + inv_base * (float)perm[0] / (1.0f - inv_base)), PB2_ONE_MINUS_EPS);
}
// SobolSampler (samplers/sobol.rs:48-58, lowdiscrepancy.rs:507-560; pbrt-v3 semantics where the port cannot run: DESIGN.md).
constexpr unsigned kSobolMatrixSize = 52u;                                 // sobolmatrices.rs:2
__device__ __forceinline__ unsigned long long sobol_index(const SamplerView& h, int px, int py, unsigned long long frame) {     // sobol_interval_to_index
    const unsigned m = (unsigned)h.sobol_log2_resolution;
    if (m == 0u) return 0ull;
    unsigned long long index = frame << (m << 1), delta = 0ull;
    const unsigned long long* vdc = h.sobol_vdc + (size_t)(m - 1u) * kSobolMatrixSize;
    const unsigned long long* inv = h.sobol_vdc_inv + (size_t)(m - 1u) * kSobolMatrixSize;
    while (frame) {                                                        // XOR over the set bits, in any order
        delta ^= __ldg(vdc + (__ffsll((long long)frame) - 1));
        frame &= frame - 1ull;
    }
    unsigned long long b = (unsigned long long)((((unsigned)px) << m) | (unsigned)py) ^ delta;
    while (b) {
        index ^= __ldg(inv + (__ffsll((long long)b) - 1));
        b &= b - 1ull;
    }
    return index;
}
__device__ __forceinline__ float sobol_raw(const SamplerView& h, unsigned long long a, unsigned dim) {       // sobol_sample, scramble = 0
    const uint32_t* m = h.sobol_m32 + (size_t)dim * kSobolMatrixSize;
    unsigned v = 0u;
    while (a) {
        v ^= __ldg(m + (__ffsll((long long)a) - 1));
        a &= a - 1ull;
    }
    return fminf(PB2_ONE_MINUS_EPS, __uint2float_rn(v) * 2.3283064365386963e-10f);
}
// SobolSampler::sample_dimension for dimensions 0 / 1 (pbrt-v3: the film position inside pixel (x, y))
__device__ __forceinline__ float sobol_pixel_dimension(const SamplerView& h, unsigned long long index, unsigned dim, int pixel) {
    float s = sobol_raw(h, index, dim);
    s = s * (float)h.sobol_resolution + (float)h.sobol_min[dim];
    s = s - (float)pixel;
    return s < 0.0f ? 0.0f : (s > PB2_ONE_MINUS_EPS ? PB2_ONE_MINUS_EPS : s);
}
// One path's sampler.  RandomSampler: the PCG32 stream; HaltonSampler: (index, dimension) of the sequence; PixelSamplers
// (stratified, (0,2)): PixelSampler::get_1d / get_2d (sampler.rs:289-307) — the next tabulated dimension of this pixel's
// sample while one is left, then the PCG32 stream.  TABLES = false compiles the table branch out (k_shade is at its register limit).
struct PathSampler {
    Pcg32 rng;
    unsigned long long index;
    unsigned dim;
    unsigned cur1, cur2;          // current_1d_dimension, current_2d_dimension
    unsigned tab_base;            // sample * tab_n_pix + pixel
    const SamplerView* h;
    __device__ __forceinline__ bool sobol() const { return h->kind == 4; }
    __device__ __forceinline__ bool global() const { return h->kind == 1 || h->kind == 4; }          // GlobalSampler (sampler.rs:324-410)
    __device__ __forceinline__ bool tables() const { return h->kind == 2 || h->kind == 3; }         // PixelSampler (:257-322)
    __device__ __forceinline__ unsigned long long global_index(const SlotInfo& si) const {         // get_index_for_sample
        return sobol() ? sobol_index(*h, si.x - h->sobol_min[0], si.y - h->sobol_min[1], si.sample)
                       : (unsigned long long)halton_index(*h, si.x, si.y, si.sample);
    }
    __device__ __forceinline__ float global_dimension(unsigned d) const {                          // sample_dimension, d >= 2 for Sobol'
        return sobol() ? sobol_raw(*h, index, d) : halton_dimension(*h, index, d);
    }
    __device__ __forceinline__ void start(const SamplerView& view, const SlotInfo& si) {   // start of a pixel sample
        h = &view;
        dim = 0u;
        cur1 = cur2 = 0u;
        tab_base = si.sample * view.tab_n_pix + si.pix;
        if (global()) index = global_index(si);
        else rng.set_sequence(si.seq);
    }
    // The first draw of every pixel sample: CameraSample::p_film's offset inside the pixel (sampler.rs:27-33).  Sobol' remaps
    // dimensions 0 / 1 to the pixel, which needs the pixel's coordinates — known here, not carried in the path's sampler state.
    __device__ __forceinline__ void film_offset(const SlotInfo& si, float* u0, float* u1) {
        if (sobol()) {
            *u0 = sobol_pixel_dimension(*h, index, 0u, si.x);
            *u1 = sobol_pixel_dimension(*h, index, 1u, si.y);
            dim = 2u;
            return;
        }
        next2(u0, u1);
    }
    // `extra` = the PixelSampler dimension counters kept in bits 17-30 of the path's state word
    __device__ __forceinline__ void resume(const SamplerView& view, const SlotInfo& si, unsigned long long saved, unsigned extra) {
        h = &view;
        cur1 = extra & 0x7Fu;
        cur2 = (extra >> 7) & 0x7Fu;
        tab_base = si.sample * view.tab_n_pix + si.pix;
        if (global()) { index = global_index(si); dim = (unsigned)saved; }
        else { rng.state = saved; rng.inc = (si.seq << 1) | 1ull; }
    }
    __device__ __forceinline__ unsigned long long save() const { return global() ? (unsigned long long)dim : rng.state; }
    __device__ __forceinline__ unsigned extra() const { return cur1 | (cur2 << 7); }
    template <bool TABLES = true>
    __device__ __forceinline__ float next1() {                                             // Sampler::get_1d
        if (TABLES && tables() && cur1 < (unsigned)h->n_dims) {
            const float v = __ldg(h->t1 + (size_t)(cur1 * (unsigned)h->spp_tab) * h->tab_n_pix + tab_base);
            ++cur1;
            return v;
        }
        if (global()) return global_dimension(dim++);
        return rng.next_float();
    }
    template <bool TABLES = true>
    __device__ __forceinline__ void next2(float* a, float* b) {                            // Sampler::get_2d, x then y
        if (TABLES && tables() && cur2 < (unsigned)h->n_dims) {
            const float2 v = __ldg(h->t2 + (size_t)(cur2 * (unsigned)h->spp_tab) * h->tab_n_pix + tab_base);
            ++cur2;
            *a = v.x; *b = v.y;
            return;
        }
        if (global()) { *a = global_dimension(dim); *b = global_dimension(dim + 1u); dim += 2u; return; }
        *a = rng.next_float();
        *b = rng.next_float();
    }
};

// ---- shade -----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ vec3 ld3(const float* p) { return mk(p[0], p[1], p[2]); }

struct Vertex {           // SurfaceInteraction subset rebuilt from the hit record (triangle.rs:193-311, D59)
    vec3 p, err, n;
    vec3 sn;              // shading.n: n unless the mesh has vertex normals / tangents
    vec3 ss, ts;          // the BSDF's frame (reflection.rs:220-234): ss = normalize(shading.dpdu), ts = cross(shading.n, ss)
    vec3 wo;              // SurfaceInteraction::wo: -ray.d for a triangle; normalize(o2w * -ray_obj.d) for a sphere (sphere.rs:79)
};
// SG = the mesh carries per-vertex normals, tangents or UVs (compiled out otherwise: k_shade is at its register limit).
// ro / rd: the ray that hit (a sphere's interaction is rebuilt from the ray and the hit distance, which a sphere hit carries in
// place of b0; sphere.rs:38-93).
// A primitive's record in SceneView::tris_prim is 6 x float4: the PackedTri (3 x float4) and, for a plain triangle, what
// Triangle::intersect + BSDF::new derive from the three vertices alone — n, ss, ts (k_tris_by_prim) — so that the plain-mesh
// vertex costs three more loads instead of two normalisations, tri_frame and a cross product (11 % of k_shade<0>'s instructions).
template <bool SG>
__device__ __forceinline__ Vertex rebuild_vertex(const SceneView& s, const ShadeView& sh, uint32_t prim, float b0, float b1, float b2,
                                                 vec3 ro = mk(0.f, 0.f, 0.f), vec3 rd = mk(0.f, 0.f, 1.f)) {
    const float4* tp = s.tris_prim + 6ull * prim;
    const float4 a = ldg4(tp), b = ldg4(tp + 1), c = ldg4(tp + 2);
    const vec3 p0 = mk(a.x, a.y, a.z), p1 = mk(b.x, b.y, b.z), p2 = mk(c.x, c.y, c.z);
    Vertex v;
    v.wo = -rd;
    if (SG && s.spheres && (__float_as_uint(c.w) & 2u)) {
        const SphereVertex sv = sphere_vertex_at(sphere_of(s, a), ro, rd, b0);
        v.p = sv.p; v.err = sv.err; v.n = sv.n; v.sn = sv.sn; v.wo = sv.wo;
        v.ss = unit(sv.dpdu);
        v.ts = cross3(v.sn, v.ss);
        return v;
    }
    const float xs = (fabsf(b0 * p0.x) + fabsf(b1 * p1.x)) + fabsf(b2 * p2.x);
    const float ys = (fabsf(b0 * p0.y) + fabsf(b1 * p1.y)) + fabsf(b2 * p2.y);
    const float zs = (fabsf(b0 * p0.z) + fabsf(b1 * p1.z)) + fabsf(b2 * p2.z);
    v.err = mk(xs, ys, zs) * gammaf_(7.0f);
    v.p = (p0 * b0 + p1 * b1) + p2 * b2;
    if (!SG || !sh.indices) {                                  // plain triangle (also in a scene with analytic spheres): the per-primitive frame
        const float4 fn = ldg4(tp + 3), fs = ldg4(tp + 4), ft = ldg4(tp + 5);
        v.n = mk(fn.x, fn.y, fn.z);
        v.sn = v.n;
        v.ss = mk(fs.x, fs.y, fs.z);
        v.ts = mk(ft.x, ft.y, ft.z);
        return v;
    }
    v.n = unit(cross3(p0 - p2, p1 - p2));
    vec3 dpdu, dv;
    const uint32_t i0 = __ldg(sh.indices + 3ull * prim), i1 = __ldg(sh.indices + 3ull * prim + 1), i2 = __ldg(sh.indices + 3ull * prim + 2);
    if (sh.uvs) tri_frame_uv(p0, p1, p2, __ldg(sh.uvs + i0), __ldg(sh.uvs + i1), __ldg(sh.uvs + i2), &dpdu, &dv);   // Triangle::get_uvs
    else tri_frame(p0, p1, p2, &dpdu, &dv);
    v.sn = v.n;
    vec3 sdpdu = dpdu;
    if (sh.normals || sh.tangents) {                                     // triangle.rs:251-311
        vec3 ns = v.n;
        if (sh.normals) {
            ns = (ld3(sh.normals + 3ull * i0) * b0 + ld3(sh.normals + 3ull * i1) * b1) + ld3(sh.normals + 3ull * i2) * b2;
            ns = len2(ns) > 0.0f ? unit(ns) : v.n;
        }
        vec3 ss = unit(dpdu);
        if (sh.tangents) {
            const vec3 st = (ld3(sh.tangents + 3ull * i0) * b0 + ld3(sh.tangents + 3ull * i1) * b1) + ld3(sh.tangents + 3ull * i2) * b2;
            if (len2(st) > 0.0f) ss = unit(st);
        }
        vec3 ts = cross3(ss, ns);
        if (len2(ts) > 0.0f) { ts = unit(ts); ss = cross3(ts, ns); }
        else coord_system(ns, &ss, &ts);
        // SurfaceInteraction::set_shading_geometry(ss, ts, .., true) (interaction.rs:297-316)
        v.sn = unit(cross3(ss, ts));
        v.n = face_toward(v.n, v.sn);                                    // D6 FIX
        sdpdu = ss;
    }
    v.ss = unit(sdpdu);                                                  // BSDF::new (reflection.rs:220-234)
    v.ts = cross3(v.sn, v.ss);
    return v;
}
// estimate_direct (integrator.rs:136-266) up to the two visibility queries: fills the NEE record of `slot`.
// out != nullptr (VolPathIntegrator, k_volpath): nothing is stored in the path buffers; the factors of both terms are handed back
// unmultiplied, because with handle_media the transmittance enters the product before f (integrator.rs:172-173, 259-261).
struct NeeOut {
    rgb3 li, f1;            // light sample: radiance and f (* |cos|)
    float w1, light_pdf;    // MIS weight (unused for a delta light) and the light's pdf
    bool delta;
    vec3 p1, p1_err, p1_n;  // the sampled point on the light (VisibilityTester's p1)
    rgb3 lmis, f2;          // BSDF / phase sample: the light's radiance towards the vertex and f (* |cos|)
    float w2, scattering_pdf;
    vec3 mis_o, mis_d;
    vec3 sh_o, sh_d;        // the first segment of the shadow ray (spawn_ray_to, interaction.rs:146-153)
    unsigned light_prim;
};
template <bool SG, class BsdfType>
__device__ __forceinline__ unsigned direct_lighting(const SceneView& s, const ShadeView& sh, const PathBuffers& b, uint32_t slot, const Vertex& v, vec3 wo,
                                                const BsdfType& bsdf, const DLight& light, float pick_pdf, float ul0, float ul1, float us0,
                                                float us1, rgb3 beta, NeeOut* out = nullptr) {
    const unsigned flags = kAllLobes & ~kSpecular;                       // D23 FIX
    const rgb3 l_emit = mkc(light.l[0], light.l[1], light.l[2]);
    const bool delta = light.type != 1;                                  // light.rs:28-31, D24 FIX: point, spot, distant
    vec3 wi = mk(0.f, 0.f, 0.f);
    float light_pdf = 0.0f, scattering_pdf = 0.0f;
    rgb3 li = gray(0.0f);
    vec3 sh_o = mk(0.f, 0.f, 0.f), sh_d = mk(0.f, 0.f, 0.f);
    const vec3 lp0 = ld3(light.p0), lp1 = ld3(light.p1), lp2 = ld3(light.p2);
    if (delta) {                                                         // point.rs:47-66, spot.rs:71-85, distant.rs:50-67
        vec3 pl = ld3(light.p);
        light_pdf = 1.0f;
        if (light.type == 3) {                                           // DistantLight: the tester's far end is p_outside
            wi = ld3(light.axis);
            pl = v.p + wi * (2.0f * light.world_radius);
            li = l_emit;
        } else {
            wi = unit(pl - v.p);
            if (light.type == 2) {                                       // SpotLight::falloff(-wi), spot.rs:51-63
                const vec3 w = -wi;
                const float cos_theta = (light.axis[0] * w.x + light.axis[1] * w.y) + light.axis[2] * w.z;
                float fall = 1.0f;
                if (cos_theta < light.cos_total_width) fall = 0.0f;
                else if (!(cos_theta >= light.cos_falloff_start)) {
                    const float dl = (cos_theta - light.cos_total_width) / (light.cos_falloff_start - light.cos_total_width);
                    fall = (dl * dl) * (dl * dl);
                }
                li = l_emit * fall / len2(pl - v.p);
            } else li = l_emit / len2(pl - v.p);
        }
        sh_o = offset_ray_origin(v.p, v.err, v.n, pl - v.p);             // interaction.rs:146-153
        const vec3 target = offset_ray_origin(pl, mk(0.f, 0.f, 0.f), mk(0.f, 0.f, 0.f), sh_o - pl);
        sh_d = target - sh_o;
        if (out) { out->p1 = pl; out->p1_err = mk(0.f, 0.f, 0.f); out->p1_n = mk(0.f, 0.f, 0.f); }
    } else if (SG && light.sphere >= 0) {                                // diffuse.rs:60-81 + sphere.rs:127-193
        vec3 ps, pe, ns;
        float pdf;
        sphere_sample2(reinterpret_cast<const DSphere*>(s.spheres)[light.sphere], v.p, v.err, v.n, ul0, ul1, &ps, &pe, &ns, &pdf);
        if (pdf == 0.0f || len2(ps - v.p) == 0.0f) { light_pdf = 0.0f; }
        else {
            light_pdf = pdf;
            wi = unit(ps - v.p);
            li = (light.two_sided || dot3(ns, -wi) > 0.0f) ? l_emit : gray(0.0f);
            sh_o = offset_ray_origin(v.p, v.err, v.n, ps - v.p);
            const vec3 target = offset_ray_origin(ps, pe, ns, sh_o - ps);
            sh_d = target - sh_o;
            if (out) { out->p1 = ps; out->p1_err = pe; out->p1_n = ns; }
        }
    } else {                                                             // diffuse.rs:60-81, shape.rs:38-53, triangle.rs:330-348
        const float su0 = sqrtf(ul0);
        const float b0 = 1.0f - su0, b1 = ul1 * su0;                     // sampling.rs:275-278
        const float b2 = (1.0f - b0) - b1;
        const vec3 ps = (lp0 * b0 + lp1 * b1) + lp2 * b2;
        vec3 ns = ld3(light.ns_sample);                                    // normalize(cross(p1 - p0, p2 - p0)), triangle.rs:336
        if (light.has_n) ns = face_toward(ns, (ld3(light.n0) * b0 + ld3(light.n1) * b1) + ld3(light.n2) * b2);     // triangle.rs:338-341, D6 FIX
        const vec3 pe = ((abs3(lp0 * b0) + abs3(lp1 * b1)) + abs3(lp2 * b2)) * gammaf_(6.0f);
        float pdf = light.inv_area;                                        // 1 / area (shape.rs:43)
        vec3 w = ps - v.p;
        if (len2(w) == 0.0f) pdf = 0.0f;
        else {
            w = unit(w);
            pdf = pdf * (len2(v.p - ps) / fabsf(dot3(ns, -w)));
            if (isinf(pdf)) pdf = 0.0f;
        }
        if (pdf == 0.0f || len2(ps - v.p) == 0.0f) { light_pdf = 0.0f; }
        else {
            light_pdf = pdf;
            wi = unit(ps - v.p);
            li = (light.two_sided || dot3(ns, -wi) > 0.0f) ? l_emit : gray(0.0f);      // D55 FIX
            sh_o = offset_ray_origin(v.p, v.err, v.n, ps - v.p);
            const vec3 target = offset_ray_origin(ps, pe, ns, sh_o - ps);
            sh_d = target - sh_o;
            if (out) { out->p1 = ps; out->p1_err = pe; out->p1_n = ns; }
        }
    }
    unsigned pending = 0u;
    rgb3 t1 = gray(0.0f), t2 = gray(0.0f);
    if (light_pdf > 0.0f && !black(li)) {
        scattering_pdf = bsdf_pdf(bsdf, wo, wi, flags);
        const rgb3 f = bsdf_f(bsdf, wo, wi, flags) * cos_factor(bsdf, wi);
        if (!black(f)) {
            pending |= 1u;                                               // VisibilityTester::un_occluded decides (D25 FIX)
            if (out) { out->li = li; out->f1 = f; out->light_pdf = light_pdf; out->delta = delta; out->w1 = delta ? 1.0f : power_heuristic(light_pdf, scattering_pdf); }
            else t1 = delta ? li * f / light_pdf : li * f * power_heuristic(light_pdf, scattering_pdf) / light_pdf;
        }
    }
    vec3 mis_o = mk(0.f, 0.f, 0.f), mis_d = mk(0.f, 0.f, 1.f);
    if (!delta) {
        unsigned sampled = 0u;
        rgb3 f = bsdf_sample_f(bsdf, wo, &wi, us0, us1, &scattering_pdf, flags, &sampled);
        f = f * cos_factor(bsdf, wi);
        const bool sampled_specular = (sampled & kSpecular) != 0u;
        if (!black(f) && scattering_pdf > 0.0f) {
            float weight = 1.0f;
            bool go = true;
            const vec3 ro = offset_ray_origin(v.p, v.err, v.n, wi);      // it.spawn_ray(wi)
            float lb0 = 0.0f, lb1 = 0.0f, lb2 = 0.0f;
            if (SG && light.sphere >= 0) {
                // Light::pdf_li -> Sphere::pdf2 (sphere.rs:195-207); then the light's normal where this ray meets the sphere (the
                // closest hit of the MIS ray is this sphere or the contribution is dropped in k_extend)
                const DSphere& sp = reinterpret_cast<const DSphere*>(s.spheres)[light.sphere];
                if (!sampled_specular) {
                    const float lp = sphere_pdf2(sp, v.p, v.err, v.n, wi);
                    if (lp == 0.0f) go = false;
                    else weight = power_heuristic(scattering_pdf, lp);
                }
                if (go) {
                    float t_l;
                    SphereVertex lv;
                    if (!sphere_intersect(sp, ro, wi, kInf, &t_l, &lv)) go = false;
                    else {
                        const rgb3 lmis = (light.two_sided || dot3(lv.n, -wi) > 0.0f) ? l_emit : gray(0.0f);
                        if (!black(lmis)) {
                            pending |= 2u;
                            if (out) { out->lmis = lmis; out->f2 = f; out->w2 = weight; out->scattering_pdf = scattering_pdf; }
                            else t2 = lmis * f * gray(1.0f) * weight / scattering_pdf;
                            mis_o = ro;
                            mis_d = wi;
                        }
                    }
                }
                go = false;                                              // handled
            } else
            if (!sampled_specular) {
                // Light::pdf_li -> Shape::pdf2 (shape.rs:54-69): the light's own triangle
                const RayCtx rc = make_ray_ctx(ro, wi);
                float t;
                const bool frame_ok = light.frame_ok != 0;               // Triangle::intersect's degenerate-frame exit, per triangle
                if (!tri_test(rc, kInf, lp0, lp1, lp2, &t, &lb0, &lb1, &lb2) || !frame_ok) go = false;
                else {
                    const vec3 p_l = (lp0 * lb0 + lp1 * lb1) + lp2 * lb2;
                    const vec3 n_l = ld3(light.n_hit);
                    float lp = len2(v.p - p_l) / (fabsf(dot3(n_l, -wi)) * light.area);
                    if (isinf(lp)) lp = 0.0f;
                    if (lp == 0.0f) go = false;
                    else weight = power_heuristic(scattering_pdf, lp);
                }
            }
            if (go) {
                // li = light_isect.le(-wi) if the closest hit is this light's triangle (D56 FIX); its normal is known here:
                // the geometric one, or — on a mesh with vertex normals / tangents — the one Triangle::intersect leaves in the
                // interaction at these barycentrics (flipped towards the shading normal, set_shading_geometry)
                vec3 n_l = ld3(light.n_hit);
                if (SG && !sampled_specular) n_l = rebuild_vertex<true>(s, sh, light.prim, lb0, lb1, lb2).n;
                const rgb3 lmis = (light.two_sided || dot3(n_l, -wi) > 0.0f) ? l_emit : gray(0.0f);
                if (!black(lmis)) {
                    pending |= 2u;
                    if (out) { out->lmis = lmis; out->f2 = f; out->w2 = weight; out->scattering_pdf = scattering_pdf; }
                    else t2 = lmis * f * gray(1.0f) * weight / scattering_pdf;
                    mis_o = ro;
                    mis_d = wi;
                }
            }
        }
    }
    if (pending == 0u) return 0u;
    if (out) {
        out->mis_o = mis_o;
        out->mis_d = mis_d;
        out->sh_o = sh_o;
        out->sh_d = sh_d;
        out->light_prim = light.prim;
        return pending;
    }
    b.sh_o[slot] = make_float4(sh_o.x, sh_o.y, sh_o.z, 1.0f - PB2_SHADOW_EPS);
    b.sh_d[slot] = make_float4(sh_d.x, sh_d.y, sh_d.z, pick_pdf);
    b.t1[slot] = make_float4(t1.r, t1.g, t1.b, __uint_as_float(pending));
    b.beta_nee[slot] = make_float4(beta.r, beta.g, beta.b, 0.0f);
    if (pending & 2u) {
        b.mis_o[slot] = make_float4(mis_o.x, mis_o.y, mis_o.z, 0.0f);
        b.mis_d[slot] = make_float4(mis_d.x, mis_d.y, mis_d.z, 0.0f);
        b.t2[slot] = make_float4(t2.r, t2.g, t2.b, __uint_as_float(light.prim));
    }
    (void)s;
    return pending;                                  // bit 0: shadow ray, bit 1: MIS ray — queued by compact_queues()
}

}  // namespace

}  // namespace pb2
