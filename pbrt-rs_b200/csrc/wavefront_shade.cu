// wavefront_shade.cu — k_shade: one path vertex of PathIntegrator::li for every hit of one shading class (sm_100a).
// Compiled once per <PB2_SHADE_TABLES, PB2_SHADE_SG> combination (Makefile): PixelSampler tables or stream samplers, plain mesh
// or the general vertex (mesh attributes / analytic spheres).
#include "wavefront_dev.cuh"

#ifndef PB2_SHADE_TABLES
#error "compile with -DPB2_SHADE_TABLES=0|1 -DPB2_SHADE_SG=0|1"
#endif

namespace pb2 {

namespace {

// One path vertex of PathIntegrator::li (path.rs:79-209) for every hit of material type `mat`.
#ifndef PB2_SHADE_BLOCKS
#define PB2_SHADE_BLOCKS 2
#endif
#ifndef PB2_SHADE_THREADS
#define PB2_SHADE_THREADS 256
#endif
#ifndef PB2_SHADE_PREFETCH
#define PB2_SHADE_PREFETCH 1    /* 0: no prefetch of the next vertex's path state; 2: into L1 (tuning builds) */
#endif
__device__ __forceinline__ void prefetch_state(const void* p) {
#if PB2_SHADE_PREFETCH == 2
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}
template <int MAT, bool TABLES, bool SG>
__global__ void __launch_bounds__(PB2_SHADE_THREADS, PB2_SHADE_BLOCKS) k_shade(SceneView s, ShadeView sh, PathBuffers b, PathMap map, FilmView film, PathParams pp, int cur) {
    constexpr int Q = MAT == 3 ? 1 : MAT;       // MAT 3 = the class-1 queue shaded by the plastic-only kernel (shade.cuh: may_be<3>)
    const uint64_t n = b.counters[C_MAT0 + Q];
    const uint32_t* queue = b.q_mat[Q];
#ifndef PB2_SHADE_PIPE
#define PB2_SHADE_PIPE 1        /* 0: load the queue entry and the hit record where they are used (the round-1 loop; tuning builds) */
#endif
    // The head of an iteration is a chain of dependent loads — queue entry -> hit record -> leaf slot -> triangle — that held
    // 22 % of the kernel's stall samples on its first two links alone (profiles/r02_tuning.md); they are issued two / one
    // iteration ahead, so their latency runs under the previous vertices' arithmetic.
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t slot_n = 0, slot_nn = 0;
    uint4 h_n = make_uint4(0u, 0u, 0u, 0u);
    if (PB2_SHADE_PIPE) {
        if (i0 < n) { slot_n = queue[i0]; h_n = b.hit[slot_n]; }
        if (i0 + stride < n) slot_nn = queue[i0 + stride];
    }
    for (uint64_t i = i0; i < n; i += stride) {
        uint32_t slot;
        uint4 h;
        if (PB2_SHADE_PIPE) {
            slot = slot_n;
            h = h_n;
            slot_n = slot_nn;
            if (i + stride < n) {
                h_n = b.hit[slot_n];
#if PB2_SHADE_PREFETCH
                // the next vertex's path state (its slot has been in a register since the previous iteration): four lines on
                // their way into L2 / L1 while this vertex is shaded — the head-of-iteration loads of ray_d, L, beta and rng
                // held 18 % of the kernel's stall samples (profiles/r02_tuning.md; prefetching the next triangle record from
                // the middle of the iteration as well measured slower and is not done)
                prefetch_state(b.ray_d + slot_n);
                prefetch_state(b.L + slot_n);
                prefetch_state(b.beta + slot_n);
                prefetch_state(b.rng + slot_n);
#endif
            }
            if (i + 2 * stride < n) slot_nn = queue[i + 2 * stride];
        } else {
            slot = queue[i];
            h = b.hit[slot];
        }
        const float4 rd = b.ray_d[slot];
        float4 Lf = b.L[slot];
        float4 bt = b.beta[slot];
        rgb3 L = mkc(Lf.x, Lf.y, Lf.z), beta = mkc(bt.x, bt.y, bt.z);
        float eta_scale = bt.w;
        const unsigned state = __float_as_uint(Lf.w);
        unsigned bounces = state & 0xFFFFu;
        const bool specular_bounce = (state >> 16) & 1u;
        vec3 ray_o = mk(0.f, 0.f, 0.f);
        if (SG && s.spheres) { const float4 ro = b.ray_o[slot]; ray_o = mk(ro.x, ro.y, ro.z); }
        const Vertex v = rebuild_vertex<SG>(s, sh, h.x, __uint_as_float(h.y), __uint_as_float(h.z), __uint_as_float(h.w), ray_o, mk(rd.x, rd.y, rd.z));
        const vec3 wo = -mk(rd.x, rd.y, rd.z);
        if (bounces == 0u || specular_bounce) {                          // path.rs:80-82 + interaction.rs:387-395
            const int li = sh.tri_light[h.x];
            if (li >= 0) {
                const DLight& lt = sh.lights[li];
                const rgb3 le = (lt.two_sided || dot3(v.n, wo) > 0.0f) ? mkc(lt.l[0], lt.l[1], lt.l[2]) : gray(0.0f);
                L = L + beta * le;
            }
        }
        bool alive = bounces < (unsigned)pp.max_depth;                   // path.rs:90-92
        unsigned queued = (unsigned)Q;                                 // b.state[slot]: class | continues << 2 | NEE rays << 3
        if (alive) {
            const auto bsdf = make_bsdf<MAT>(sh.mats[sh.tri_material[h.x]], v.n, v.sn, v.ss, v.ts);
            PathSampler rng;
            rng.resume(map.smp, slot_info(map, film, slot), b.rng[slot], TABLES ? (state >> 17) & 0x3FFFu : 0u);
            // (class 2 holds specular lobes only — FresnelSpecular, SpecularReflection — so estimate_direct is never reached there)
            if (MAT != 2 && bsdf_count(bsdf, kAllLobes & ~kSpecular) > 0 && sh.n_lights > 0) {      // path.rs:105-121, integrator.rs:99-134
                float pick_pdf;
                // light_distribution.lookup(&isect.p) (path.rs:100-104)
                const float *l_cdf = sh.light_cdf, *l_func = sh.light_func;
                float l_int = sh.light_func_int;
                if (sh.spatial.func) {
                    const size_t vox = spatial_voxel(sh.spatial, v.p);
                    l_cdf = sh.spatial.cdf + vox * (size_t)(sh.n_lights + 1);
                    l_func = sh.spatial.func + vox * (size_t)sh.n_lights;
                    l_int = __ldg(sh.spatial.func_int + vox);
                }
                const int li = sample_discrete(l_cdf, l_func, sh.n_lights, l_int, rng.next1<TABLES>(), &pick_pdf);
                if (pick_pdf != 0.0f) {
                    float ul0, ul1, us0, us1;
                    rng.next2<TABLES>(&ul0, &ul1);
                    rng.next2<TABLES>(&us0, &us1);
                    queued |= direct_lighting<SG>(s, sh, b, slot, v, SG ? v.wo : wo, bsdf, sh.lights[li], pick_pdf, ul0, ul1, us0, us1, beta) << 3;   // estimate_direct reads it.wo
                }
            }
            float u0, u1;
            rng.next2<TABLES>(&u0, &u1);                                                  // path.rs:123-134
            vec3 wi = mk(0.f, 0.f, 0.f);
            float pdf = 0.0f;
            unsigned sampled = 0u;
            const rgb3 f = bsdf_sample_f(bsdf, wo, &wi, u0, u1, &pdf, kAllLobes, &sampled);
            if (black(f) || pdf == 0.0f) alive = false;
            else {
                beta = beta * (f * (fabsf(dot3(wi, bsdf.ns)) / pdf));
                const bool spec = (sampled & kSpecular) != 0u;
                if (spec && (sampled & kTransmission)) {
                    const float eta = bsdf.eta;
                    eta_scale = eta_scale * ((dot3(wo, v.n) > 0.0f) ? (eta * eta) : 1.0f / (eta * eta));
                }
                const vec3 o = offset_ray_origin(v.p, v.err, v.n, wi);
                const rgb3 rr_beta = beta * eta_scale;                                   // path.rs:200-207, D27 KEEP
                if (max_channel(rr_beta) < pp.rr_threshold && bounces > 3u) {
                    const float q = fminf(1.0f - max_channel(rr_beta), 0.05f);
                    if (rng.next1<TABLES>() < q) alive = false;
                    else beta = beta / (1.0f - q);
                }
                if (alive) {
                    bounces += 1u;
                    b.ray_o[slot] = make_float4(o.x, o.y, o.z, kInf);
                    b.ray_d[slot] = make_float4(wi.x, wi.y, wi.z, 0.0f);
                    b.beta[slot] = make_float4(beta.r, beta.g, beta.b, eta_scale);
                    b.rng[slot] = rng.save();
                    Lf.w = __uint_as_float(bounces | ((spec ? 1u : 0u) << 16) | (TABLES ? rng.extra() << 17 : 0u));
                    queued |= kStateContinues;
                }
            }
        }
        b.state[slot] = (uint8_t)queued;
        b.L[slot] = make_float4(L.r, L.g, L.b, Lf.w);
    }
}

}  // namespace

// k_shade<material, PixelSampler tables, mesh shading geometry> for the three material queues of one bounce.
template <>
void launch_shade_t<PB2_SHADE_TABLES != 0, PB2_SHADE_SG != 0>(Wavefront* wf, const SceneView& sv, const ShadeView& sh, const PathBuffers& b, const PathMap& map, const FilmView& film,
                    const PathParams& pp, int cur, uint64_t n, cudaStream_t st) {
    constexpr bool TABLES = PB2_SHADE_TABLES != 0, SG = PB2_SHADE_SG != 0;
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n + PB2_SHADE_THREADS - 1) / PB2_SHADE_THREADS,
                                                                             (uint64_t)wf->sm_count * 2 * PB2_SHADE_BLOCKS));
    // (a class no material of the scene has leaves its queue empty every bounce: no launch)
    if (sh.class_mask & 1u) k_shade<0, TABLES, SG><<<grid, PB2_SHADE_THREADS, 0, st>>>(sv, sh, b, map, film, pp, cur);
    if ((sh.class_mask & 10u) == 10u) k_shade<3, TABLES, SG><<<grid, PB2_SHADE_THREADS, 0, st>>>(sv, sh, b, map, film, pp, cur);   // bit 3: every class-1 material is a two-lobe plastic
    else if (sh.class_mask & 2u) k_shade<1, TABLES, SG><<<grid, PB2_SHADE_THREADS, 0, st>>>(sv, sh, b, map, film, pp, cur);
    if (sh.class_mask & 4u) k_shade<2, TABLES, SG><<<grid, PB2_SHADE_THREADS, 0, st>>>(sv, sh, b, map, film, pp, cur);
}

}  // namespace pb2
