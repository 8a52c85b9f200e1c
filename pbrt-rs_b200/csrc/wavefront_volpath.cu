// wavefront_volpath.cu — VolPathIntegrator over HomogeneousMedium: k_volpath (sm_100a).
#include "wavefront_dev.cuh"

namespace pb2 {

namespace {

// ---- VolPathIntegrator (src/integrators/volpath.rs) over HomogeneousMedium (src/media/homogeneous.rs) ----------------------
// One thread carries one camera sample through the whole of VolPathIntegrator::li (volpath.rs:60-244): medium sampling, the
// phase-function or BSDF vertex, next-event estimation with transmittance (VisibilityTester::tr, light.rs:137-160) and the
// MIS ray through Scene::intersect_tr (scene.rs:48-71).  Every ray is a closest-hit walk (transmittance rays pass through
// material-less interface surfaces segment by segment), done here with the literal one-level walk of traverse.cuh; the
// wavefront stages of the PathIntegrator are not involved.  Defect ledger D69-D74: spawned rays take
// GetMedium(d), a material-less surface is not a bounce, tr / sample use min and the exponential.
struct VolHit {
    HitRec h;
    Vertex v;
};
__device__ __forceinline__ bool vol_intersect(const SceneView& s, const ShadeView& sh, vec3 o, vec3 d, float t_max, VolHit* out) {
    if (!traverse<false, true>(s, o, d, t_max, &out->h)) return false;
    out->v = rebuild_vertex<true>(s, sh, out->h.prim, out->h.sphere ? out->h.t : out->h.b0, out->h.b1, out->h.b2, o, d);
    return true;
}
// homogeneous.rs:36-38 (D71)
__device__ __forceinline__ rgb3 medium_tr(const DMedium& m, float t_max, vec3 d) {
    const float sdist = fminf(t_max * len(d), 3.402823466e+38f);
    return mkc(det_exp(-(m.sigma_t[0] * sdist)), det_exp(-(m.sigma_t[1] * sdist)), det_exp(-(m.sigma_t[2] * sdist)));
}
// primitive.rs:72-76
__device__ __forceinline__ void hit_interface(const ShadeView& sh, uint32_t prim, int ray_medium, int* inside, int* outside) {
    const int pi = sh.prim_inside ? sh.prim_inside[prim] : -1, po = sh.prim_outside ? sh.prim_outside[prim] : -1;
    if (pi != po) { *inside = pi; *outside = po; }
    else { *inside = ray_medium; *outside = ray_medium; }
}
// VisibilityTester::tr (light.rs:137-160, D74)
__device__ __forceinline__ rgb3 visibility_tr(const SceneView& s, const ShadeView& sh, vec3 p, vec3 err, vec3 n, int med_in, int med_out, vec3 p1,
                                              vec3 p1_err, vec3 p1_n, unsigned long long* n_rays) {
    rgb3 tr = gray(1.0f);
    for (;;) {
        const vec3 origin = offset_ray_origin(p, err, n, p1 - p);
        const vec3 target = offset_ray_origin(p1, p1_err, p1_n, origin - p1);
        const vec3 d = target - origin;
        const int medium = dot3(d, n) > 0.0f ? med_out : med_in;
        VolHit hit;
        ++*n_rays;
        const bool found = vol_intersect(s, sh, origin, d, 1.0f - PB2_SHADOW_EPS, &hit);
        if (found && sh.tri_material[hit.h.prim] != 0xFFFFFFFFu) return gray(0.0f);
        if (medium >= 0) tr = tr * medium_tr(sh.media[medium], found ? hit.h.t : 1.0f - PB2_SHADOW_EPS, d);
        if (!found) break;
        hit_interface(sh, hit.h.prim, medium, &med_in, &med_out);
        p = hit.v.p; err = hit.v.err; n = hit.v.n;
    }
    return tr;
}
// Scene::intersect_tr (scene.rs:48-71, D74): true when the ray ends on a surface with a material (its primitive in *prim)
__device__ __forceinline__ bool intersect_tr(const SceneView& s, const ShadeView& sh, vec3 o, vec3 d, int medium, uint32_t* prim, rgb3* tr,
                                             unsigned long long* n_rays) {
    *tr = gray(1.0f);
    for (;;) {
        VolHit hit;
        ++*n_rays;
        const bool found = vol_intersect(s, sh, o, d, kInf, &hit);
        if (medium >= 0) *tr = *tr * medium_tr(sh.media[medium], found ? hit.h.t : kInf, d);
        if (!found) return false;
        if (sh.tri_material[hit.h.prim] != 0xFFFFFFFFu) { *prim = hit.h.prim; return true; }
        int in, out;
        hit_interface(sh, hit.h.prim, medium, &in, &out);
        o = offset_ray_origin(hit.v.p, hit.v.err, hit.v.n, d);
        medium = dot3(d, hit.v.n) > 0.0f ? out : in;
    }
}
// uniform_sample_one_light + estimate_direct with handle_media (integrator.rs:92-266) at a surface or medium vertex
template <class BsdfType>
__device__ __forceinline__ rgb3 vol_sample_one_light(const SceneView& s, const ShadeView& sh, const PathBuffers& b, const Vertex& v, vec3 wo_si,
                                                     const BsdfType& bsdf, int med_in, int med_out, PathSampler& smp, unsigned long long* n_shadow,
                                                     unsigned long long* n_mis) {
    if (sh.n_lights <= 0) return gray(0.0f);
    float pick_pdf;
    const float *l_cdf = sh.light_cdf, *l_func = sh.light_func;
    float l_int = sh.light_func_int;
    if (sh.spatial.func) {
        const size_t vox = spatial_voxel(sh.spatial, v.p);
        l_cdf = sh.spatial.cdf + vox * (size_t)(sh.n_lights + 1);
        l_func = sh.spatial.func + vox * (size_t)sh.n_lights;
        l_int = __ldg(sh.spatial.func_int + vox);
    }
    const int li_idx = sample_discrete(l_cdf, l_func, sh.n_lights, l_int, smp.next1(), &pick_pdf);
    if (pick_pdf == 0.0f) return gray(0.0f);
    float ul0, ul1, us0, us1;
    smp.next2(&ul0, &ul1);
    smp.next2(&us0, &us1);
    NeeOut ne;
    const unsigned pending = direct_lighting<true>(s, sh, b, 0u, v, wo_si, bsdf, sh.lights[li_idx], pick_pdf, ul0, ul1, us0, us1, gray(1.0f), &ne);
    rgb3 ld = gray(0.0f);
    if (pending & 1u) {
        const rgb3 li = ne.li * visibility_tr(s, sh, v.p, v.err, v.n, med_in, med_out, ne.p1, ne.p1_err, ne.p1_n, n_shadow);
        if (!black(li)) ld = ld + (ne.delta ? li * ne.f1 / ne.light_pdf : li * ne.f1 * ne.w1 / ne.light_pdf);
    }
    if (pending & 2u) {
        uint32_t prim;
        rgb3 tr;
        const int medium = dot3(ne.mis_d, v.n) > 0.0f ? med_out : med_in;
        if (intersect_tr(s, sh, ne.mis_o, ne.mis_d, medium, &prim, &tr, n_mis) && prim == ne.light_prim)
            ld = ld + ne.lmis * ne.f2 * tr * ne.w2 / ne.scattering_pdf;
    }
    return ld / pick_pdf;
}
// The surface vertex of volpath.rs:133-187 for shading class CLS: NEE, then the BSDF sample that continues the path.
// Returns false when the path ends (black f or zero pdf).
template <int CLS>
__device__ __forceinline__ bool vol_surface(const SceneView& s, const ShadeView& sh, const PathBuffers& b, const VolHit& hit, vec3 ray_d, int med_in,
                                            int med_out, PathSampler& smp, rgb3* L, rgb3* beta, float* eta_scale, bool* specular, vec3* wi_out,
                                            unsigned long long* n_shadow, unsigned long long* n_mis) {
    const Vertex& v = hit.v;
    const auto bsdf = make_bsdf<CLS>(sh.mats[sh.tri_material[hit.h.prim]], v.n, v.sn, v.ss, v.ts);
    *L = *L + *beta * vol_sample_one_light(s, sh, b, v, v.wo, bsdf, med_in, med_out, smp, n_shadow, n_mis);      // at every surface vertex (volpath.rs:137-146)
    const vec3 wo = -ray_d;
    float u0, u1, pdf = 0.0f;
    smp.next2(&u0, &u1);
    unsigned sampled = 0u;
    vec3 wi = mk(0.f, 0.f, 0.f);
    const rgb3 f = bsdf_sample_f(bsdf, wo, &wi, u0, u1, &pdf, kAllLobes, &sampled);
    if (black(f) || pdf == 0.0f) return false;
    *beta = *beta * (f * (fabsf(dot3(wi, bsdf.ns)) / pdf));
    *specular = (sampled & kSpecular) != 0u;
    if ((sampled & kSpecular) && (sampled & kTransmission)) {
        const float eta = bsdf.eta;
        *eta_scale = *eta_scale * ((dot3(wo, v.n) > 0.0f) ? (eta * eta) : 1.0f / (eta * eta));
    }
    *wi_out = wi;
    return true;
}
__global__ void __launch_bounds__(128) k_volpath(uint64_t n, SceneView s, ShadeView sh, PathBuffers b, PathMap map, FilmView film, CameraView cam,
                                                 PathParams pp) {
    unsigned long long n_extend = 0, n_shadow = 0, n_mis = 0;
    for (uint64_t slot = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += (uint64_t)gridDim.x * blockDim.x) {
        const SlotInfo si = slot_info(map, film, slot);
        PathSampler smp;
        smp.start(map.smp, si);
        float u0, u1, l0 = 0.0f, l1 = 0.0f;
        smp.film_offset(si, &u0, &u1);
        if (smp.global() && !(cam.lens_radius > 0.0f)) smp.dim += 3u;
        else { (void)smp.next1(); smp.next2(&l0, &l1); }
        vec3 o, d;
        float t_max;
        camera_ray(cam, (float)si.x + u0, (float)si.y + u1, l0, l1, &o, &d, &t_max);
        int ray_medium = sh.camera_medium;
        rgb3 L = gray(0.0f), beta = gray(1.0f);
        bool specular_bounce = false;
        int bounces = 0;
        float eta_scale = 1.0f;
        for (;;) {
            VolHit hit;
            ++n_extend;
            const bool found = vol_intersect(s, sh, o, d, t_max, &hit);
            bool in_medium = false;
            vec3 mp = mk(0.f, 0.f, 0.f);
            if (ray_medium >= 0) {                                       // HomogeneousMedium::sample (homogeneous.rs:40-74; D72, D73)
                const DMedium& m = sh.media[ray_medium];
                const float ray_t_max = found ? hit.h.t : t_max;
                const float uc = smp.next1() * 3.0f;
                const int channel = min(__float2int_rz(uc), 2);
                const float dist = -det_log(1.0f - smp.next1()) / m.sigma_t[channel < 0 ? 0 : channel];
                const float dl = len(d);
                const float t = fminf(dist / dl, ray_t_max);
                in_medium = t < ray_t_max;
                if (in_medium) mp = o + d * t;
                const float tt = fminf(t, 3.402823466e+38f);
                const rgb3 tr = mkc(det_exp(-m.sigma_t[0] * tt * dl), det_exp(-m.sigma_t[1] * tt * dl), det_exp(-m.sigma_t[2] * tt * dl));
                const rgb3 density = in_medium ? mkc(m.sigma_t[0] * tr.r, m.sigma_t[1] * tr.g, m.sigma_t[2] * tr.b) : tr;
                float pdf = 0.0f;
                pdf += density.r; pdf += density.g; pdf += density.b;
                pdf *= 1.0f / 3.0f;
                if (pdf == 0.0f) pdf = 1.0f;
                beta = beta * (in_medium ? (tr * mkc(m.sigma_s[0], m.sigma_s[1], m.sigma_s[2])) / pdf : tr / pdf);
            }
            if (black(beta)) break;
            if (in_medium) {
                if (bounces >= pp.max_depth) break;
                const DMedium& m = sh.media[ray_medium];
                const vec3 wo = -d;
                vec3 wi = mk(0.f, 0.f, 0.f);
                float p0, p1;
                smp.next2(&p0, &p1);
                hg_sample_p(m.g, wo, &wi, p0, p1);                       // volpath.rs:96-103: sampled before the light (KEEP)
                Vertex v;
                v.p = mp; v.err = mk(0.f, 0.f, 0.f); v.n = mk(0.f, 0.f, 0.f); v.sn = v.n; v.ss = v.n; v.ts = v.n; v.wo = wo;
                const PhaseHG ph{m.g, mk(0.f, 0.f, 0.f)};
                o = mp;                                                  // mi.spawn_ray(wi): no normal, no offset; the medium stays
                d = wi;
                t_max = kInf;
                specular_bounce = false;
                L = L + beta * vol_sample_one_light(s, sh, b, v, wo, ph, ray_medium, ray_medium, smp, &n_shadow, &n_mis);
            } else {
                if (bounces == 0 || specular_bounce) {
                    if (found) {
                        const int li = sh.tri_light[hit.h.prim];
                        if (li >= 0) {
                            const DLight& lt = sh.lights[li];
                            const rgb3 le = (lt.two_sided || dot3(hit.v.n, -d) > 0.0f) ? mkc(lt.l[0], lt.l[1], lt.l[2]) : gray(0.0f);
                            L = L + beta * le;
                        }
                    }
                }
                if (!found || bounces >= pp.max_depth) break;
                int in, out;
                hit_interface(sh, hit.h.prim, ray_medium, &in, &out);
                const uint32_t mat = sh.tri_material[hit.h.prim];
                if (mat == 0xFFFFFFFFu) {                                // volpath.rs:127-131 (D70): crosses the interface, not a bounce
                    o = offset_ray_origin(hit.v.p, hit.v.err, hit.v.n, d);
                    t_max = kInf;
                    ray_medium = dot3(d, hit.v.n) > 0.0f ? out : in;
                    continue;
                }
                vec3 wi;
                bool go;
                const int cls = sh.mats[mat].cls;
                if (cls == 0) go = vol_surface<0>(s, sh, b, hit, d, in, out, smp, &L, &beta, &eta_scale, &specular_bounce, &wi, &n_shadow, &n_mis);
                else if (cls == 1) go = vol_surface<1>(s, sh, b, hit, d, in, out, smp, &L, &beta, &eta_scale, &specular_bounce, &wi, &n_shadow, &n_mis);
                else go = vol_surface<2>(s, sh, b, hit, d, in, out, smp, &L, &beta, &eta_scale, &specular_bounce, &wi, &n_shadow, &n_mis);
                if (!go) break;
                o = offset_ray_origin(hit.v.p, hit.v.err, hit.v.n, wi);
                d = wi;
                t_max = kInf;
                ray_medium = dot3(wi, hit.v.n) > 0.0f ? out : in;
            }
            const rgb3 rr_beta = beta * eta_scale;
            if (max_channel(rr_beta) < pp.rr_threshold && bounces > 3) {
                const float q = fmaxf(1.0f - max_channel(rr_beta), 0.05f);    // volpath.rs:236
                if (smp.next1() < q) break;
                beta = beta / (1.0f - q);
            }
            bounces += 1;
        }
        b.L[slot] = make_float4(L.r, L.g, L.b, 0.0f);
    }
    // ray totals of the launch (pb2_render_counters): one atomic per warp and counter
    for (int off = 16; off > 0; off >>= 1) {
        n_extend += __shfl_down_sync(0xFFFFFFFFu, n_extend, off);
        n_shadow += __shfl_down_sync(0xFFFFFFFFu, n_shadow, off);
        n_mis += __shfl_down_sync(0xFFFFFFFFu, n_mis, off);
    }
    if ((threadIdx.x & 31u) == 0u) {
        atomicAdd(&b.counters[T_EXTEND], n_extend);
        atomicAdd(&b.counters[T_SHADOW], n_shadow);
        atomicAdd(&b.counters[T_MIS], n_mis);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&b.counters[T_CAMERA], (unsigned long long)n);
}


// ============================ the wavefront VolPathIntegrator ====================================================================
// The same transport as k_volpath above (which stays as the one-thread-per-path form, PB2_VOLPATH_MEGAKERNEL=1), cut into stages
// that talk through the PathBuffers / VolBuffers arrays and the slot queues of the wavefront PathIntegrator, so that every ray —
// path rays, and each segment of the shadow and MIS rays that gather transmittance — is walked by the persistent QuadNode
// traversal (trace_persistent.cuh) with full warps instead of by one lane of a divergent loop.  One iteration:
//   k_vol_extend   closest hit of every active path ray (hit record + distance)
//   k_vol_medium   HomogeneousMedium::sample on the segment; a medium vertex is finished here (phase sample, NEE record, Russian
//                  roulette); a surface hit adds Le, crosses a material-less interface (not a bounce) or is left to ...
//   k_vol_surface  ... the kernel of its shading class (select by class in between, as k_shade): NEE record, BSDF sample, RR
//   select         next active queue, shadow-ray queue, MIS-ray queue
//   k_vol_trace + k_vol_tr_post   one segment of every pending shadow / MIS ray (closest hit), then VisibilityTester::tr /
//                  Scene::intersect_tr's loop body: blocked, arrived, or through an interface into the next segment — repeated
//                  (queues re-selected, counts read back) while a ray has a segment left; scenes without interfaces take one round
//   k_vol_resolve  l += beta * ld / pick_pdf for the iteration's NEE records, in the reference's order of terms
// A path's additions to L happen in the order of volpath.rs (Le at a vertex, then that vertex's direct light, then the next
// vertex), and every path draws its sampler in that order too, so the radiance equals k_volpath's bit for bit.
#ifndef PB2_VOL_BLIND_ROUNDS
#define PB2_VOL_BLIND_ROUNDS 1  /* transmittance rounds run without a read-back of the pending-ray counts (interface scenes) */
#endif
#ifndef PB2_VOL_BLOCKS
#define PB2_VOL_BLOCKS 4        /* CTAs of 128 threads per SM of k_vol_medium / k_vol_surface: 128 registers with 200-440 B of spills beat the 156-195 registers / 3 CTAs the compiler takes freely (18.22 -> 17.54 ms on the fog + smoke frame) */
#endif
constexpr uint8_t kTrDone = 3, kTrAgain = 0;       // VolBuffers::st_s / st_m after k_vol_tr_post (the select's class predicate)

__device__ __forceinline__ uint32_t pack_path_word(unsigned bounces, bool spec, unsigned extra) { return bounces | ((spec ? 1u : 0u) << 16) | (extra << 17); }

// uniform_sample_one_light + estimate_direct with handle_media (integrator.rs:92-266) up to the visibility queries: draws the
// light and the two 2D samples, writes the NEE record of `slot` and returns its pending bits (1 shadow ray, 2 MIS ray).
template <class BsdfType>
__device__ __forceinline__ unsigned vol_nee_record(const SceneView& s, const ShadeView& sh, const PathBuffers& b, const VolBuffers& vb, uint32_t slot,
                                                   const Vertex& v, vec3 wo_si, const BsdfType& bsdf, int med_in, int med_out, PathSampler& smp, rgb3 beta) {
    if (sh.n_lights <= 0) return 0u;
    float pick_pdf;
    const float *l_cdf = sh.light_cdf, *l_func = sh.light_func;
    float l_int = sh.light_func_int;
    if (sh.spatial.func) {
        const size_t vox = spatial_voxel(sh.spatial, v.p);
        l_cdf = sh.spatial.cdf + vox * (size_t)(sh.n_lights + 1);
        l_func = sh.spatial.func + vox * (size_t)sh.n_lights;
        l_int = __ldg(sh.spatial.func_int + vox);
    }
    const int li_idx = sample_discrete(l_cdf, l_func, sh.n_lights, l_int, smp.next1(), &pick_pdf);
    if (pick_pdf == 0.0f) return 0u;
    float ul0, ul1, us0, us1;
    smp.next2(&ul0, &ul1);
    smp.next2(&us0, &us1);
    NeeOut ne;
    const unsigned pending = direct_lighting<true>(s, sh, b, 0u, v, wo_si, bsdf, sh.lights[li_idx], pick_pdf, ul0, ul1, us0, us1, gray(1.0f), &ne);
    if (pending == 0u) return 0u;
    b.beta_nee[slot] = make_float4(beta.r, beta.g, beta.b, 0.0f);
    int med_s = -1, med_m = -1;
    if (pending & 1u) {
        b.sh_o[slot] = make_float4(ne.sh_o.x, ne.sh_o.y, ne.sh_o.z, 1.0f - PB2_SHADOW_EPS);
        b.sh_d[slot] = make_float4(ne.sh_d.x, ne.sh_d.y, ne.sh_d.z, pick_pdf);
        b.t1[slot] = make_float4(ne.li.r, ne.li.g, ne.li.b, ne.w1);
        vb.f1[slot] = make_float4(ne.f1.r, ne.f1.g, ne.f1.b, ne.delta ? -ne.light_pdf : ne.light_pdf);
        vb.p1[slot] = make_float4(ne.p1.x, ne.p1.y, ne.p1.z, 0.0f);
        vb.p1_err[slot] = make_float4(ne.p1_err.x, ne.p1_err.y, ne.p1_err.z, 0.0f);
        vb.p1_n[slot] = make_float4(ne.p1_n.x, ne.p1_n.y, ne.p1_n.z, 0.0f);
        vb.tr_s[slot] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
        med_s = dot3(ne.sh_d, v.n) > 0.0f ? med_out : med_in;            // light.rs:146
    } else b.sh_d[slot] = make_float4(0.0f, 0.0f, 0.0f, pick_pdf);
    if (pending & 2u) {
        const rgb3 a2 = ne.lmis * ne.f2;
        b.mis_o[slot] = make_float4(ne.mis_o.x, ne.mis_o.y, ne.mis_o.z, ne.scattering_pdf);
        b.mis_d[slot] = make_float4(ne.mis_d.x, ne.mis_d.y, ne.mis_d.z, ne.w2);
        b.t2[slot] = make_float4(a2.r, a2.g, a2.b, __uint_as_float(ne.light_prim));
        vb.tr_m[slot] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
        med_m = dot3(ne.mis_d, v.n) > 0.0f ? med_out : med_in;
    }
    vb.ray_med[slot] = make_int2(med_s, med_m);
    return pending;
}

// Russian roulette of volpath.rs:227-238; false: the path ends here.
__device__ __forceinline__ bool vol_roulette(const PathParams& pp, PathSampler& smp, rgb3* beta, float eta_scale, unsigned bounces) {
    const rgb3 rr_beta = *beta * eta_scale;
    if (max_channel(rr_beta) < pp.rr_threshold && bounces > 3u) {
        const float q = fmaxf(1.0f - max_channel(rr_beta), 0.05f);       // volpath.rs:236
        if (smp.next1() < q) return false;
        *beta = *beta / (1.0f - q);
    }
    return true;
}

// ---- closest hit of the active path rays -------------------------------------------------------------------------------------------
struct VolExtendSink {
    PathBuffers b;
    VolBuffers vb;
    const uint32_t* queue;
    PB2_D bool load(uint32_t i, vec3* o, vec3* d, float* t_max) const {
        const uint32_t slot = queue[i];
        const float4 ro = b.ray_o[slot], rd = b.ray_d[slot];
        *o = mk(ro.x, ro.y, ro.z);
        *d = mk(rd.x, rd.y, rd.z);
        *t_max = ro.w;
        return true;
    }
    PB2_D void accept(uint32_t i, uint32_t prim, float, float b0, float b1, float b2) const {
        b.hit[queue[i]] = make_uint4(prim, __float_as_uint(b0), __float_as_uint(b1), __float_as_uint(b2));
    }
    PB2_D void accept_sphere(uint32_t i, uint32_t prim, float t, float u, float v) const {
        b.hit[queue[i]] = make_uint4(prim, __float_as_uint(t), __float_as_uint(u), __float_as_uint(v));
    }
    PB2_D void finish(uint32_t i, bool found, float t) const {
        const uint32_t slot = queue[i];
        vb.t_path[slot] = t;
        b.state[slot] = found ? (uint8_t)0 : (uint8_t)kStateDead;
    }
    PB2_D void occluded(uint32_t, bool) const {}
};
__global__ void __launch_bounds__(128, PB2_MIN_BLOCKS) k_vol_extend(SceneView s, PathBuffers b, VolBuffers vb, int cur, TraceTuning tune) {
    const VolExtendSink sink{b, vb, b.q_active[cur]};
    trace_persistent<false, false>(s, (uint32_t)b.counters[C_ACTIVE_A + cur], &b.counters[C_WORK_EXTEND], sink, tune);
}
__global__ void __launch_bounds__(128, 4) k_vol_extend_spheres(SceneView s, PathBuffers b, VolBuffers vb, int cur, TraceTuning tune) {
    const VolExtendSink sink{b, vb, b.q_active[cur]};
    trace_persistent<false, true>(s, (uint32_t)b.counters[C_ACTIVE_A + cur], &b.counters[C_WORK_EXTEND], sink, tune);
}

// ---- one segment of every pending shadow / MIS ray -----------------------------------------------------------------------------------
struct VolTrSink {
    PathBuffers b;
    VolBuffers vb;
    const uint32_t* qs;
    const uint32_t* qm;
    uint32_t n_s;
    PB2_D bool load(uint32_t i, vec3* o, vec3* d, float* t_max) const {
        const bool mis = i >= n_s;
        const uint32_t slot = mis ? qm[i - n_s] : qs[i];
        const float4 ro = mis ? b.mis_o[slot] : b.sh_o[slot], rd = mis ? b.mis_d[slot] : b.sh_d[slot];
        *o = mk(ro.x, ro.y, ro.z);
        *d = mk(rd.x, rd.y, rd.z);
        *t_max = mis ? kInf : ro.w;
        return true;
    }
    PB2_D void accept(uint32_t i, uint32_t prim, float, float b0, float b1, float b2) const {
        const uint4 h = make_uint4(prim, __float_as_uint(b0), __float_as_uint(b1), __float_as_uint(b2));
        if (i >= n_s) vb.hit_m[qm[i - n_s]] = h; else vb.hit_s[qs[i]] = h;
    }
    PB2_D void accept_sphere(uint32_t i, uint32_t prim, float t, float u, float v) const {
        const uint4 h = make_uint4(prim, __float_as_uint(t), __float_as_uint(u), __float_as_uint(v));
        if (i >= n_s) vb.hit_m[qm[i - n_s]] = h; else vb.hit_s[qs[i]] = h;
    }
    PB2_D void finish(uint32_t i, bool found, float t) const {
        if (i >= n_s) { const uint32_t slot = qm[i - n_s]; vb.t_m[slot] = t; vb.st_m[slot] = found ? 1 : 0; }
        else { const uint32_t slot = qs[i]; vb.t_s[slot] = t; vb.st_s[slot] = found ? 1 : 0; }
    }
    PB2_D void occluded(uint32_t, bool) const {}
};
__global__ void __launch_bounds__(128, PB2_MIN_BLOCKS) k_vol_trace(SceneView s, PathBuffers b, VolBuffers vb, const uint32_t* qs, int cs, const uint32_t* qm,
                                                                   int cm, TraceTuning tune) {
    const uint32_t n_s = (uint32_t)b.counters[cs], n_m = (uint32_t)b.counters[cm];
    const VolTrSink sink{b, vb, qs, qm, n_s};
    trace_persistent<false, false>(s, n_s + n_m, &b.counters[C_WORK_SHADOW], sink, tune);
}
__global__ void __launch_bounds__(128, 4) k_vol_trace_spheres(SceneView s, PathBuffers b, VolBuffers vb, const uint32_t* qs, int cs, const uint32_t* qm, int cm,
                                                              TraceTuning tune) {
    const uint32_t n_s = (uint32_t)b.counters[cs], n_m = (uint32_t)b.counters[cm];
    const VolTrSink sink{b, vb, qs, qm, n_s};
    trace_persistent<false, true>(s, n_s + n_m, &b.counters[C_WORK_SHADOW], sink, tune);
}
// The loop bodies of VisibilityTester::tr (light.rs:137-160) and Scene::intersect_tr (scene.rs:48-71) for the segment just walked.
// round > 0: the rays of this round were not counted by the select that queued the first segments.
__global__ void __launch_bounds__(kThreads) k_vol_tr_post(SceneView s, ShadeView sh, PathBuffers b, VolBuffers vb, const uint32_t* qs, int cs,
                                                          const uint32_t* qm, int cm, int round) {
    const uint64_t n_s = b.counters[cs], n_m = b.counters[cm];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        b.counters[C_WORK_SHADOW] = 0;                                   // the walk is over: ready for the next round / iteration
        if (round > 0) { b.counters[T_SHADOW] += n_s; b.counters[T_MIS] += n_m; }
    }
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_s + n_m; i += (uint64_t)gridDim.x * blockDim.x) {
        if (i < n_s) {
            const uint32_t slot = qs[i];
            const bool found = vb.st_s[slot] != 0;
            const uint4 h = vb.hit_s[slot];
            uint8_t st = kTrDone;
            if (found && sh.tri_material[h.x] != 0xFFFFFFFFu) vb.tr_s[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);       // blocked
            else {
                const float4 ro = b.sh_o[slot], rd = b.sh_d[slot];
                const vec3 o = mk(ro.x, ro.y, ro.z), d = mk(rd.x, rd.y, rd.z);
                const int medium = vb.ray_med[slot].x;
                if (medium >= 0) {
                    const float4 t4 = vb.tr_s[slot];
                    const rgb3 tr = mkc(t4.x, t4.y, t4.z) * medium_tr(sh.media[medium], vb.t_s[slot], d);      // t = the hit's, or the ray's 1 - eps
                    vb.tr_s[slot] = make_float4(tr.r, tr.g, tr.b, 0.0f);
                }
                if (found) {                                             // through the interface, re-aimed at p1
                    int in, out;
                    hit_interface(sh, h.x, medium, &in, &out);
                    const Vertex v = rebuild_vertex<true>(s, sh, h.x, __uint_as_float(h.y), __uint_as_float(h.z), __uint_as_float(h.w), o, d);
                    const float4 q1 = vb.p1[slot], qe = vb.p1_err[slot], qn = vb.p1_n[slot];
                    const vec3 p1 = mk(q1.x, q1.y, q1.z);
                    const vec3 origin = offset_ray_origin(v.p, v.err, v.n, p1 - v.p);
                    const vec3 target = offset_ray_origin(p1, mk(qe.x, qe.y, qe.z), mk(qn.x, qn.y, qn.z), origin - p1);
                    const vec3 nd = target - origin;
                    b.sh_o[slot] = make_float4(origin.x, origin.y, origin.z, ro.w);
                    b.sh_d[slot] = make_float4(nd.x, nd.y, nd.z, rd.w);
                    vb.ray_med[slot].x = dot3(nd, v.n) > 0.0f ? out : in;
                    st = kTrAgain;
                }
            }
            vb.st_s[slot] = st;
        } else {
            const uint32_t slot = qm[i - n_s];
            const bool found = vb.st_m[slot] != 0;
            const uint4 h = vb.hit_m[slot];
            const float4 ro = b.mis_o[slot], rd = b.mis_d[slot];
            const vec3 o = mk(ro.x, ro.y, ro.z), d = mk(rd.x, rd.y, rd.z);
            const int medium = vb.ray_med[slot].y;
            float4 t4 = vb.tr_m[slot];
            if (medium >= 0) {
                const rgb3 tr = mkc(t4.x, t4.y, t4.z) * medium_tr(sh.media[medium], vb.t_m[slot], d);
                t4 = make_float4(tr.r, tr.g, tr.b, 0.0f);
            }
            uint8_t st = kTrDone;
            if (found) {
                if (sh.tri_material[h.x] != 0xFFFFFFFFu) { t4.w = 1.0f; b.mis_prim[slot] = h.x; }       // ended on a surface
                else {
                    int in, out;
                    hit_interface(sh, h.x, medium, &in, &out);
                    const Vertex v = rebuild_vertex<true>(s, sh, h.x, __uint_as_float(h.y), __uint_as_float(h.z), __uint_as_float(h.w), o, d);
                    const vec3 no = offset_ray_origin(v.p, v.err, v.n, d);
                    b.mis_o[slot] = make_float4(no.x, no.y, no.z, ro.w);
                    vb.ray_med[slot].y = dot3(d, v.n) > 0.0f ? out : in;
                    st = kTrAgain;
                }
            }
            vb.tr_m[slot] = t4;
            vb.st_m[slot] = st;
        }
    }
}
// estimate_direct's sum (integrator.rs:172-191, 243-261) and l += beta * ld / pick_pdf for the NEE records of this iteration.
__global__ void __launch_bounds__(kThreads) k_vol_resolve(PathBuffers b, VolBuffers vb, int cur) {
    const uint64_t n = b.counters[C_ACTIVE_A + cur];
    const uint32_t* queue = b.q_active[cur];
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t slot = queue[i];
        const unsigned pending = (b.state[slot] >> 3) & 3u;
        if (pending == 0u) continue;
        rgb3 ld = gray(0.0f);
        if (pending & 1u) {
            const float4 t1 = b.t1[slot], f1 = vb.f1[slot], tr = vb.tr_s[slot];
            const rgb3 li = mkc(t1.x, t1.y, t1.z) * mkc(tr.x, tr.y, tr.z);
            const rgb3 f = mkc(f1.x, f1.y, f1.z);
            const float light_pdf = fabsf(f1.w);
            if (!black(li)) ld = ld + (f1.w < 0.0f ? li * f / light_pdf : li * f * t1.w / light_pdf);
        }
        if (pending & 2u) {
            const float4 t2 = b.t2[slot], tr = vb.tr_m[slot];
            if (tr.w == 1.0f && b.mis_prim[slot] == __float_as_uint(t2.w))
                ld = ld + mkc(t2.x, t2.y, t2.z) * mkc(tr.x, tr.y, tr.z) * b.mis_d[slot].w / b.mis_o[slot].w;
        }
        const float pick_pdf = b.sh_d[slot].w;
        const float4 bn = b.beta_nee[slot];
        float4 Lf = b.L[slot];
        const rgb3 L = mkc(Lf.x, Lf.y, Lf.z) + mkc(bn.x, bn.y, bn.z) * (ld / pick_pdf);
        b.L[slot] = make_float4(L.r, L.g, L.b, Lf.w);
    }
}

// ---- the medium stage: volpath.rs:76-131 for every active path -----------------------------------------------------------------------
__global__ void __launch_bounds__(128, PB2_VOL_BLOCKS) k_vol_medium(SceneView s, ShadeView sh, PathBuffers b, VolBuffers vb, PathMap map, FilmView film, PathParams pp, int cur) {
    const uint64_t n = b.counters[C_ACTIVE_A + cur];
    const uint32_t* queue = b.q_active[cur];
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t slot = queue[i];
        const bool found = b.state[slot] != kStateDead;
        const float4 ro = b.ray_o[slot], rd = b.ray_d[slot];
        const vec3 o = mk(ro.x, ro.y, ro.z), d = mk(rd.x, rd.y, rd.z);
        float4 Lf = b.L[slot];
        const float4 bt = b.beta[slot];
        rgb3 L = mkc(Lf.x, Lf.y, Lf.z), beta = mkc(bt.x, bt.y, bt.z);
        const float eta_scale = bt.w;
        const unsigned word = __float_as_uint(Lf.w);
        unsigned bounces = word & 0xFFFFu;
        const bool specular_bounce = (word >> 16) & 1u;
        const int ray_medium = vb.med[slot];
        PathSampler smp;
        smp.resume(map.smp, slot_info(map, film, slot), b.rng[slot], (word >> 17) & 0x3FFFu);
        bool in_medium = false;
        vec3 mp = mk(0.f, 0.f, 0.f);
        if (ray_medium >= 0) {                                           // HomogeneousMedium::sample (homogeneous.rs:40-74; D72, D73)
            const DMedium& m = sh.media[ray_medium];
            const float ray_t_max = vb.t_path[slot];                     // found ? the hit's t : the ray's t_max
            const float uc = smp.next1() * 3.0f;
            const int channel = min(__float2int_rz(uc), 2);
            const float dist = -det_log(1.0f - smp.next1()) / m.sigma_t[channel < 0 ? 0 : channel];
            const float dl = len(d);
            const float t = fminf(dist / dl, ray_t_max);
            in_medium = t < ray_t_max;
            if (in_medium) mp = o + d * t;
            const float tt = fminf(t, 3.402823466e+38f);
            const rgb3 tr = mkc(det_exp(-m.sigma_t[0] * tt * dl), det_exp(-m.sigma_t[1] * tt * dl), det_exp(-m.sigma_t[2] * tt * dl));
            const rgb3 density = in_medium ? mkc(m.sigma_t[0] * tr.r, m.sigma_t[1] * tr.g, m.sigma_t[2] * tr.b) : tr;
            float pdf = 0.0f;
            pdf += density.r; pdf += density.g; pdf += density.b;
            pdf *= 1.0f / 3.0f;
            if (pdf == 0.0f) pdf = 1.0f;
            beta = beta * (in_medium ? (tr * mkc(m.sigma_s[0], m.sigma_s[1], m.sigma_s[2])) / pdf : tr / pdf);
        }
        unsigned st = kStateDead;                                        // class 3: no surface stage; bit 2 continues; bits 3-4 NEE rays
        bool spec = specular_bounce;
        if (black(beta)) {
            // volpath.rs:93: the path ends
        } else if (in_medium) {
            if (bounces < (unsigned)pp.max_depth) {
                const DMedium& m = sh.media[ray_medium];
                const vec3 wo = -d;
                vec3 wi = mk(0.f, 0.f, 0.f);
                float p0, p1;
                smp.next2(&p0, &p1);
                hg_sample_p(m.g, wo, &wi, p0, p1);                       // volpath.rs:96-103: sampled before the light (KEEP)
                Vertex v;
                v.p = mp; v.err = mk(0.f, 0.f, 0.f); v.n = mk(0.f, 0.f, 0.f); v.sn = v.n; v.ss = v.n; v.ts = v.n; v.wo = wo;
                const PhaseHG ph{m.g, mk(0.f, 0.f, 0.f)};
                st |= vol_nee_record(s, sh, b, vb, slot, v, wo, ph, ray_medium, ray_medium, smp, beta) << 3;
                spec = false;
                if (vol_roulette(pp, smp, &beta, eta_scale, bounces)) {
                    bounces += 1u;
                    b.ray_o[slot] = make_float4(mp.x, mp.y, mp.z, kInf); // mi.spawn_ray(wi): no normal, no offset; the medium stays
                    b.ray_d[slot] = make_float4(wi.x, wi.y, wi.z, 0.0f);
                    st |= kStateContinues;
                }
            }
        } else {
            const uint4 h = b.hit[slot];
            if ((bounces == 0u || specular_bounce) && found) {
                const int li = sh.tri_light[h.x];
                if (li >= 0) {
                    // (the interaction's normal: rebuilt only for an emitter that is not two-sided)
                    const DLight& lt = sh.lights[li];
                    rgb3 le = mkc(lt.l[0], lt.l[1], lt.l[2]);
                    if (!lt.two_sided) {
                        const Vertex v = rebuild_vertex<true>(s, sh, h.x, __uint_as_float(h.y), __uint_as_float(h.z), __uint_as_float(h.w), o, d);
                        if (!(dot3(v.n, -d) > 0.0f)) le = gray(0.0f);
                    }
                    L = L + beta * le;
                }
            }
            if (found && bounces < (unsigned)pp.max_depth) {
                const uint32_t mat = sh.tri_material[h.x];
                if (mat == 0xFFFFFFFFu) {                                // volpath.rs:127-131 (D70): crosses the interface, not a bounce
                    int in, out;
                    hit_interface(sh, h.x, ray_medium, &in, &out);
                    const Vertex v = rebuild_vertex<true>(s, sh, h.x, __uint_as_float(h.y), __uint_as_float(h.z), __uint_as_float(h.w), o, d);
                    const vec3 no = offset_ray_origin(v.p, v.err, v.n, d);
                    b.ray_o[slot] = make_float4(no.x, no.y, no.z, kInf);
                    vb.med[slot] = dot3(d, v.n) > 0.0f ? out : in;
                    st |= kStateContinues;
                } else st = (unsigned)sh.mats[mat].cls;                  // k_vol_surface<cls> goes on from here
            }
        }
        b.beta[slot] = make_float4(beta.r, beta.g, beta.b, eta_scale);
        b.rng[slot] = smp.save();
        b.L[slot] = make_float4(L.r, L.g, L.b, __uint_as_float(pack_path_word(bounces, spec, smp.extra())));
        b.state[slot] = (uint8_t)st;
    }
}

// ---- the surface vertex of volpath.rs:133-187 for the hits of shading class CLS ----------------------------------------------------
template <int CLS>
__global__ void __launch_bounds__(128, PB2_VOL_BLOCKS) k_vol_surface(SceneView s, ShadeView sh, PathBuffers b, VolBuffers vb, PathMap map, FilmView film, PathParams pp) {
    const uint64_t n = b.counters[C_MAT0 + CLS];
    const uint32_t* queue = b.q_mat[CLS];
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t slot = queue[i];
        const uint4 h = b.hit[slot];
        const float4 ro = b.ray_o[slot], rd = b.ray_d[slot];
        const vec3 o = mk(ro.x, ro.y, ro.z), d = mk(rd.x, rd.y, rd.z);
        const float4 Lf = b.L[slot];
        const float4 bt = b.beta[slot];
        rgb3 beta = mkc(bt.x, bt.y, bt.z);
        float eta_scale = bt.w;
        const unsigned word = __float_as_uint(Lf.w);
        unsigned bounces = word & 0xFFFFu;
        const int ray_medium = vb.med[slot];
        PathSampler smp;
        smp.resume(map.smp, slot_info(map, film, slot), b.rng[slot], (word >> 17) & 0x3FFFu);
        const Vertex v = rebuild_vertex<true>(s, sh, h.x, __uint_as_float(h.y), __uint_as_float(h.z), __uint_as_float(h.w), o, d);
        int in, out;
        hit_interface(sh, h.x, ray_medium, &in, &out);
        const auto bsdf = make_bsdf<CLS>(sh.mats[sh.tri_material[h.x]], v.n, v.sn, v.ss, v.ts);
        unsigned st = (unsigned)CLS;
        st |= vol_nee_record(s, sh, b, vb, slot, v, v.wo, bsdf, in, out, smp, beta) << 3;          // at every surface vertex (volpath.rs:137-146)
        const vec3 wo = -d;
        float u0, u1, pdf = 0.0f;
        smp.next2(&u0, &u1);
        unsigned sampled = 0u;
        vec3 wi = mk(0.f, 0.f, 0.f);
        const rgb3 f = bsdf_sample_f(bsdf, wo, &wi, u0, u1, &pdf, kAllLobes, &sampled);
        bool spec = false;
        if (!(black(f) || pdf == 0.0f)) {
            beta = beta * (f * (fabsf(dot3(wi, bsdf.ns)) / pdf));
            spec = (sampled & kSpecular) != 0u;
            if ((sampled & kSpecular) && (sampled & kTransmission)) {
                const float eta = bsdf.eta;
                eta_scale = eta_scale * ((dot3(wo, v.n) > 0.0f) ? (eta * eta) : 1.0f / (eta * eta));
            }
            if (vol_roulette(pp, smp, &beta, eta_scale, bounces)) {
                bounces += 1u;
                const vec3 no = offset_ray_origin(v.p, v.err, v.n, wi);
                b.ray_o[slot] = make_float4(no.x, no.y, no.z, kInf);
                b.ray_d[slot] = make_float4(wi.x, wi.y, wi.z, 0.0f);
                b.beta[slot] = make_float4(beta.r, beta.g, beta.b, eta_scale);
                vb.med[slot] = dot3(wi, v.n) > 0.0f ? out : in;
                st |= kStateContinues;
            }
        }
        b.rng[slot] = smp.save();
        b.L[slot] = make_float4(Lf.x, Lf.y, Lf.z, __uint_as_float(pack_path_word(bounces, spec, smp.extra())));
        b.state[slot] = (uint8_t)st;
    }
}
__global__ void __launch_bounds__(kThreads) k_vol_begin(uint64_t n, VolBuffers vb, int camera_medium) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) vb.med[i] = camera_medium;
}

int vol_buffers_create(Wavefront* wf) {
    if (wf->vol_arena) return 0;
    const size_t cap = wf->capacity, f4 = cap * 16;
    const size_t bytes = 8 * f4 + cap * 8 + 4 * cap * 4 + 2 * cap + 16 * 256;
    cudaError_t e = cudaMalloc(&wf->vol_arena, bytes);
    if (e != cudaSuccess) { wf->vol_arena = nullptr; return (int)e; }
    char* p = (char*)wf->vol_arena;
    auto take = [&](size_t n) { char* r = p; p += (n + 255) & ~(size_t)255; return r; };
    VolBuffers& vb = wf->vol;
    vb.f1 = (float4*)take(f4); vb.p1 = (float4*)take(f4); vb.p1_err = (float4*)take(f4); vb.p1_n = (float4*)take(f4);
    vb.tr_s = (float4*)take(f4); vb.tr_m = (float4*)take(f4);
    vb.hit_s = (uint4*)take(f4); vb.hit_m = (uint4*)take(f4);
    vb.ray_med = (int2*)take(cap * 8);
    vb.med = (int32_t*)take(cap * 4); vb.t_path = (float*)take(cap * 4); vb.t_s = (float*)take(cap * 4); vb.t_m = (float*)take(cap * 4);
    vb.st_s = (uint8_t*)take(cap); vb.st_m = (uint8_t*)take(cap);
    return 0;
}

}  // namespace

void launch_volpath(Wavefront* wf, unsigned grid, uint64_t n, const SceneView& sv, const ShadeView& sh, const PathBuffers& b, const PathMap& map,
                    const FilmView& film, const CameraView& cam, const PathParams& pp, cudaStream_t st) {
    (void)wf;
    k_volpath<<<grid, 128, 0, st>>>(n, sv, sh, b, map, film, cam, pp);
}


// All iterations of one batch of n path slots of the wavefront VolPathIntegrator.
void trace_batch_vol(Wavefront* wf, const SceneView& sv, const ShadeView& sh, const CameraView& cam, const FilmView& film, const PathMap& map,
                     const PathParams& pp, uint64_t n, cudaStream_t st) {
    if (vol_buffers_create(wf) != 0) {
        fprintf(stderr, "pbrt_b200: out of device memory for the VolPathIntegrator's path state; falling back to k_volpath\n");
        launch_volpath(wf, (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n + 127) / 128, (uint64_t)wf->sm_count * 8)), n, sv, sh, wf->b, map, film, cam, pp, st);
        return;
    }
    PathBuffers& b = wf->b;
    const VolBuffers& vb = wf->vol;
    const TraceTuning tune = trace_tuning();
    const unsigned trace_grid = (unsigned)wf->sm_count * (unsigned)PB2_MIN_BLOCKS, trace_grid_sph = (unsigned)wf->sm_count * 4u;
    const unsigned wide = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n + kThreads - 1) / kThreads, (uint64_t)wf->sm_count * 8));
    const unsigned stage = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n + 127) / 128, (uint64_t)wf->sm_count * 8));
    auto read_counts = [&]() {                             // queue counts of the stream so far, on the host
        cudaMemcpyAsync(wf->h_counters, b.counters, C_COUNT * 8, cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
    };
    launch_raygen(wf, n, map, film, cam, st);
    k_vol_begin<<<wide, kThreads, 0, st>>>(n, vb, sh.camera_medium);
    uint64_t launches = 2;
    for (int it = 0;; ++it) {
        const int cur = it & 1;
        bool counts_fresh = false;                         // h_counters read back after this iteration's queue select
        if (sv.spheres) k_vol_extend_spheres<<<trace_grid_sph, 128, 0, st>>>(sv, b, vb, cur, tune);
        else k_vol_extend<<<trace_grid, 128, 0, st>>>(sv, b, vb, cur, tune);
        k_vol_medium<<<stage, 128, 0, st>>>(sv, sh, b, vb, map, film, pp, cur);
        select_queues(wf, b.q_active[cur], C_ACTIVE_A + cur, true, b.q_mat[0], C_MAT0, b.q_mat[1], C_MAT1, b.q_mat[2], C_MAT2, st);
        if (sh.class_mask & 1u) k_vol_surface<0><<<stage, 128, 0, st>>>(sv, sh, b, vb, map, film, pp);
        if (sh.class_mask & 2u) k_vol_surface<1><<<stage, 128, 0, st>>>(sv, sh, b, vb, map, film, pp);
        if (sh.class_mask & 4u) k_vol_surface<2><<<stage, 128, 0, st>>>(sv, sh, b, vb, map, film, pp);
        launches += 3 + __builtin_popcount(sh.class_mask & 7u);
        // without interfaces every path is at bounce `it`: none survives iteration max_depth and none leaves an NEE record there
        if (!sh.has_interfaces && it >= pp.max_depth) break;
        select_queues(wf, b.q_active[cur], C_ACTIVE_A + cur, false, b.q_active[cur ^ 1], C_ACTIVE_A + (cur ^ 1), b.q_shadow, C_SHADOW, b.q_mis, C_MIS, st);
        launches += 1;
        if (sh.n_lights > 0) {
            uint32_t *qs = b.q_shadow, *qm = b.q_mis;
            uint32_t *qs_alt = b.q_mat[0], *qm_alt = b.q_mat[1];         // (the class queues are free until the next iteration)
            int cs = C_SHADOW, cm = C_MIS, cs_alt = C_VOL_S_ALT, cm_alt = C_VOL_M_ALT;
            for (int round = 0;; ++round) {
                if (sv.spheres) k_vol_trace_spheres<<<trace_grid_sph, 128, 0, st>>>(sv, b, vb, qs, cs, qm, cm, tune);
                else k_vol_trace<<<trace_grid, 128, 0, st>>>(sv, b, vb, qs, cs, qm, cm, tune);
                k_vol_tr_post<<<wide, kThreads, 0, st>>>(sv, sh, b, vb, qs, cs, qm, cm, round);
                launches += 2;
                if (!sh.has_interfaces) break;                           // every ray ended on its first segment
                select_queues(wf, qs, cs, true, qs_alt, cs_alt, b.q_mat[2], C_VOL_SCRATCH, b.q_mat[2], C_VOL_SCRATCH, st, vb.st_s);
                select_queues(wf, qm, cm, true, qm_alt, cm_alt, b.q_mat[2], C_VOL_SCRATCH, b.q_mat[2], C_VOL_SCRATCH, st, vb.st_m);
                launches += 2;
                // (round 0 of PB2_VOL_BLIND_ROUNDS is followed by the next one without asking: a scene with interfaces usually has
                // rays that cross one, and an empty round costs less than the read-back it saves)
                if (round >= PB2_VOL_BLIND_ROUNDS) {
                    read_counts();
                    counts_fresh = true;
                    if (wf->h_counters[cs_alt] == 0 && wf->h_counters[cm_alt] == 0) break;
                }
                std::swap(qs, qs_alt);
                std::swap(qm, qm_alt);
                std::swap(cs, cs_alt);
                std::swap(cm, cm_alt);
            }
            k_vol_resolve<<<wide, kThreads, 0, st>>>(b, vb, cur);
            launches += 1;
        }
        // crossing an interface is not a bounce: go on while a path is alive.  Before iteration max_depth some path is (nothing
        // counts down faster than one bounce per iteration), and the count of the next active queue was fixed by the select above:
        // a read-back of the transmittance rounds already holds it.
        if (sh.has_interfaces && it >= pp.max_depth) {
            if (!counts_fresh) read_counts();
            if (wf->h_counters[C_ACTIVE_A + (cur ^ 1)] == 0) break;
        }
    }
    wf->totals[4] += launches;
}

}  // namespace pb2
