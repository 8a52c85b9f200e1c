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

}  // namespace

void launch_volpath(Wavefront* wf, unsigned grid, uint64_t n, const SceneView& sv, const ShadeView& sh, const PathBuffers& b, const PathMap& map,
                    const FilmView& film, const CameraView& cam, const PathParams& pp, cudaStream_t st) {
    (void)wf;
    k_volpath<<<grid, 128, 0, st>>>(n, sv, sh, b, map, film, cam, pp);
}

}  // namespace pb2
