// kernels_traverse.cu — batched Primitive::intersect / intersect_p kernels and the ray builders of the
// ray-casting workloads (sm_100a; compile with -fmad=false).
#include "kernels.hpp"
#include "camera.cuh"
#include "trace_persistent.cuh"

#include <cstdlib>

namespace pb2 {

// ---------------------------------------------------------------------------------------------------------
// k_closest_hit / k_any_hit: bvh.rs:828-879 / :881-932 over a batch.  Rays are 2 x float4 (o,t_max | d,time), hits one
// uint4 {prim_id, t, b1, b2}.  The walk itself is trace_persistent.cuh.
// ---------------------------------------------------------------------------------------------------------

struct BatchSink {
    const float4* __restrict__ rays;
    uint4* __restrict__ hits;
    float* __restrict__ b0_out;
    uint8_t* __restrict__ occ_out;
    PB2_D bool load(uint32_t i, vec3* o, vec3* d, float* t_max) const {
        const float4 ro = __ldg(rays + 2ull * i);
        const float4 rd = __ldg(rays + 2ull * i + 1);
        *o = mk(ro.x, ro.y, ro.z);
        *d = mk(rd.x, rd.y, rd.z);
        *t_max = ro.w;
        return true;
    }
    PB2_D void accept(uint32_t i, uint32_t prim, float t, float b0, float b1, float b2) const {
        hits[i] = make_uint4(prim, __float_as_uint(t), __float_as_uint(b1), __float_as_uint(b2));
        if (b0_out) b0_out[i] = b0;
    }
    PB2_D void accept_sphere(uint32_t i, uint32_t prim, float t, float u, float v) const {
        hits[i] = make_uint4(prim, __float_as_uint(t), __float_as_uint(u), __float_as_uint(v));
        if (b0_out) b0_out[i] = 0.0f;
    }
    PB2_D void finish(uint32_t i, bool found, float t_max) const {
        if (found) return;
        hits[i] = make_uint4(0xFFFFFFFFu, __float_as_uint(t_max), 0u, 0u);
        if (b0_out) b0_out[i] = 0.0f;
    }
    PB2_D void occluded(uint32_t i, bool occ) const { occ_out[i] = occ ? 1 : 0; }
};

__global__ void __launch_bounds__(128, PB2_MIN_BLOCKS) k_closest_hit(SceneView s, const float4* __restrict__ rays, uint32_t n,
                                                      unsigned long long* __restrict__ counter, uint4* __restrict__ hits,
                                                      float* __restrict__ b0_out, TraceTuning tune) {
    const BatchSink sink{rays, hits, b0_out, nullptr};
    trace_persistent<false, false>(s, n, counter, sink, tune);
}

__global__ void __launch_bounds__(128, PB2_MIN_BLOCKS) k_any_hit(SceneView s, const float4* __restrict__ rays, uint32_t n,
                                                  unsigned long long* __restrict__ counter, uint8_t* __restrict__ out,
                                                  TraceTuning tune) {
    const BatchSink sink{rays, nullptr, nullptr, out};
    trace_persistent<true, false>(s, n, counter, sink, tune);
}

// The same walks over a scene that holds analytic spheres next to its triangles (sphere.cuh): the EFloat quadratic needs more
// registers than the triangle-only kernels are allowed, so these are separate kernels and triangle scenes never pay for them.
__global__ void __launch_bounds__(128, 4) k_closest_hit_spheres(SceneView s, const float4* __restrict__ rays, uint32_t n,
                                                                unsigned long long* __restrict__ counter, uint4* __restrict__ hits,
                                                                float* __restrict__ b0_out, TraceTuning tune) {
    const BatchSink sink{rays, hits, b0_out, nullptr};
    trace_persistent<false, true>(s, n, counter, sink, tune);
}
__global__ void __launch_bounds__(128, 4) k_any_hit_spheres(SceneView s, const float4* __restrict__ rays, uint32_t n,
                                                            unsigned long long* __restrict__ counter, uint8_t* __restrict__ out, TraceTuning tune) {
    const BatchSink sink{rays, nullptr, nullptr, out};
    trace_persistent<true, true>(s, n, counter, sink, tune);
}

// Scheduling knobs (results do not depend on them).  Defaults tuned on B200 with the C3 workload; the environment
// variables exist for the tuning sweeps recorded in profiles/.
static TraceTuning& tuning_ref() {
    static TraceTuning t = [] {
        TraceTuning v{12, 8, 18, 0};   // profiles/r01_tuning.md: sweep on the QuadNode kernel
        if (const char* e = getenv("PB2_REFILL_BELOW")) v.refill_below = atoi(e);
        if (const char* e = getenv("PB2_NODE_QUORUM")) v.node_quorum = atoi(e);
        if (const char* e = getenv("PB2_LEAF_QUORUM")) v.leaf_quorum = atoi(e);
        if (const char* e = getenv("PB2_PREFETCH")) v.prefetch = atoi(e);
        return v;
    }();
    return t;
}
TraceTuning trace_tuning() { return tuning_ref(); }
void set_trace_tuning(int refill_below, int node_quorum, int leaf_quorum, int prefetch) {
    TraceTuning& t = tuning_ref();
    if (refill_below >= 0) t.refill_below = refill_below;
    if (node_quorum >= 0) t.node_quorum = node_quorum;
    if (leaf_quorum >= 0) t.leaf_quorum = leaf_quorum;
    if (prefetch >= 0) t.prefetch = prefetch;
}

// Grid = every SM filled to the occupancy the kernel reaches (queried once), capped by the amount of work.
static unsigned persistent_grid(const void* kernel, uint64_t n) {
    static int sm_count = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!sm_count) cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 128, 0);
    if (per_sm < 1) per_sm = 1;
    static const int cap = getenv("PB2_GRID_PER_SM") ? atoi(getenv("PB2_GRID_PER_SM")) : 0;      // occupancy experiments
    if (cap > 0 && cap < per_sm) per_sm = cap;
    const uint64_t want = (n + 127) / 128;
    const uint64_t full = (uint64_t)sm_count * (uint64_t)per_sm;
    return (unsigned)(want < full ? want : full);
}

// PackedTri.pad (v2.w) = 1 when Triangle::intersect rejects every hit of the triangle (degenerate dpdu/dpdv AND zero
// geometric normal, triangle.rs:193-215): computed once per scene with the same tri_frame() the walk used to call per hit.
// With mesh UVs (TriangleMesh::uv) the frame, and so the flag, follows them (indices: 3 vertex ids per caller triangle).
__global__ void __launch_bounds__(256) k_mark_degenerate(float4* __restrict__ tris, uint32_t n, const uint32_t* __restrict__ indices,
                                                         const float2* __restrict__ uvs) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = tris[3ull * i], b = tris[3ull * i + 1];
    float4 c = tris[3ull * i + 2];
    if (__float_as_uint(c.w) & 2u) return;              // an analytic sphere's slot (bvh_build.hpp: kPrimSphere)
    vec3 du, dv;
    bool ok;
    if (uvs) {
        const uint32_t prim = __float_as_uint(a.w);
        ok = tri_frame_uv(mk(a.x, a.y, a.z), mk(b.x, b.y, b.z), mk(c.x, c.y, c.z), uvs[indices[3ull * prim]], uvs[indices[3ull * prim + 1]],
                          uvs[indices[3ull * prim + 2]], &du, &dv);
    } else ok = tri_frame(mk(a.x, a.y, a.z), mk(b.x, b.y, b.z), mk(c.x, c.y, c.z), &du, &dv);
    c.w = __uint_as_float(ok ? 0u : 1u);
    tris[3ull * i + 2] = c;
}
// tris_prim[prim] = tris[slot of prim] + the primitive's shading frame: the leaf-order records regrouped in the caller's primitive
// order (SceneView::tris_prim, 6 x float4 per primitive), followed by what Triangle::intersect and BSDF::new derive from the three
// vertices of a plain triangle — n = normalize(cross(dp02, dp12)) (triangle.rs:234-236), ss = normalize(dpdu) with the default
// UVs (triangle.rs:193-215), ts = cross(n, ss) (reflection.rs:220-234) — computed here once with the functions k_shade used per vertex.
__global__ void __launch_bounds__(256) k_tris_by_prim(const float4* __restrict__ tris, uint32_t n, float4* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = tris[3ull * i], b = tris[3ull * i + 1], c = tris[3ull * i + 2];
    const uint32_t prim = __float_as_uint(a.w);
    float4* o = out + 6ull * prim;
    o[0] = a;
    o[1] = b;
    o[2] = c;
    vec3 nn = mk(0.f, 0.f, 0.f), ss = nn, ts = nn;
    if (!(__float_as_uint(c.w) & 2u)) {                        // (an analytic sphere's interaction is rebuilt from the ray)
        const vec3 p0 = mk(a.x, a.y, a.z), p1 = mk(b.x, b.y, b.z), p2 = mk(c.x, c.y, c.z);
        vec3 dpdu, dv;
        nn = unit(cross3(p0 - p2, p1 - p2));
        tri_frame(p0, p1, p2, &dpdu, &dv);
        ss = unit(dpdu);
        ts = cross3(nn, ss);
    }
    o[3] = make_float4(nn.x, nn.y, nn.z, 0.0f);
    o[4] = make_float4(ss.x, ss.y, ss.z, 0.0f);
    o[5] = make_float4(ts.x, ts.y, ts.z, 0.0f);
}
void launch_tris_by_prim(const void* d_tris, uint64_t n, void* d_out, cudaStream_t st) {
    if (n) k_tris_by_prim<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const float4*)d_tris, (uint32_t)n, (float4*)d_out);
}
void launch_mark_degenerate(void* d_tris, uint64_t n, const void* d_indices, const void* d_uvs, cudaStream_t st) {
    if (n == 0) return;
    k_mark_degenerate<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((float4*)d_tris, (uint32_t)n, (const uint32_t*)d_indices, (const float2*)d_uvs);
}

void launch_closest_hit(const SceneView& s, const void* d_rays, uint64_t n, void* d_hits, void* d_b0, unsigned long long* d_counter,
                        cudaStream_t st) {
    if (n == 0) return;
    cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), st);
    if (s.spheres)
        k_closest_hit_spheres<<<persistent_grid((const void*)k_closest_hit_spheres, n), 128, 0, st>>>(s, (const float4*)d_rays, (uint32_t)n, d_counter,
                                                                                                   (uint4*)d_hits, (float*)d_b0, trace_tuning());
    else
    k_closest_hit<<<persistent_grid((const void*)k_closest_hit, n), 128, 0, st>>>(s, (const float4*)d_rays, (uint32_t)n, d_counter, (uint4*)d_hits,
                                                                                 (float*)d_b0, trace_tuning());
}

void launch_any_hit(const SceneView& s, const void* d_rays, uint64_t n, void* d_out, unsigned long long* d_counter, cudaStream_t st) {
    if (n == 0) return;
    cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), st);
    if (s.spheres)
        k_any_hit_spheres<<<persistent_grid((const void*)k_any_hit_spheres, n), 128, 0, st>>>(s, (const float4*)d_rays, (uint32_t)n, d_counter,
                                                                                           (uint8_t*)d_out, trace_tuning());
    else
    k_any_hit<<<persistent_grid((const void*)k_any_hit, n), 128, 0, st>>>(s, (const float4*)d_rays, (uint32_t)n, d_counter, (uint8_t*)d_out, trace_tuning());
}

// Camera::generate_ray: camera.cuh

__global__ void __launch_bounds__(256) k_camera_rays(CameraView cam, const float2* __restrict__ p_film, const float2* __restrict__ p_lens,
                                                      uint64_t n, float4* __restrict__ rays) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float fx, fy;
    if (p_film) {
        const float2 p = p_film[i];
        fx = p.x; fy = p.y;
    } else {
        fx = (float)(uint32_t)(i % (uint64_t)cam.res_x) + 0.5f;
        fy = (float)(uint32_t)(i / (uint64_t)cam.res_x) + 0.5f;
    }
    vec3 o, d;
    float t_max;
    const float2 pl = p_lens ? p_lens[i] : make_float2(0.5f, 0.5f);           // lens centre when no lens samples are given
    camera_ray(cam, fx, fy, pl.x, pl.y, &o, &d, &t_max);
    rays[2 * i] = make_float4(o.x, o.y, o.z, t_max);
    rays[2 * i + 1] = make_float4(d.x, d.y, d.z, 0.0f);
}

void launch_camera_rays(const CameraView& cam, const void* d_pfilm, const void* d_plens, uint64_t n, void* d_rays, cudaStream_t st) {
    if (n == 0) return;
    k_camera_rays<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(cam, (const float2*)d_pfilm, (const float2*)d_plens, n, (float4*)d_rays);
}

// ---------------------------------------------------------------------------------------------------------
// Secondary rays of the C3 workload.  The interaction is rebuilt from the hit record as Triangle::intersect does
// (triangle.rs:217-250): p_hit = b0*p0 + b1*p1 + b2*p2, p_error = gamma(7) * sum|b_i p_i|, n = normalize(dp02 x dp12).
// ---------------------------------------------------------------------------------------------------------
struct SurfPoint {
    vec3 p, p_err, n;
};

PB2_D bool rebuild_hit(const SceneView& s, vec3 o, vec3 d, uint32_t slot, SurfPoint* sp) {
    const float4 a = ldg4(s.tris + 3ull * slot), b = ldg4(s.tris + 3ull * slot + 1), c = ldg4(s.tris + 3ull * slot + 2);
    const vec3 p0 = mk(a.x, a.y, a.z), p1 = mk(b.x, b.y, b.z), p2 = mk(c.x, c.y, c.z);
    const RayCtx r = make_ray_ctx(o, d);
    float t, b0, b1, b2;
    if (!tri_test(r, __int_as_float(0x7f800000), p0, p1, p2, &t, &b0, &b1, &b2)) return false;
    const float xs = (fabsf(b0 * p0.x) + fabsf(b1 * p1.x)) + fabsf(b2 * p2.x);
    const float ys = (fabsf(b0 * p0.y) + fabsf(b1 * p1.y)) + fabsf(b2 * p2.y);
    const float zs = (fabsf(b0 * p0.z) + fabsf(b1 * p1.z)) + fabsf(b2 * p2.z);
    sp->p_err = mk(xs, ys, zs) * gammaf_(7.0f);
    sp->p = (p0 * b0 + p1 * b1) + p2 * b2;
    sp->n = unit(cross3(p0 - p2, p1 - p2));
    return true;
}

// interaction.rs:138-153 spawn_ray_to(point): o = offset_ray_origin(p, err, n, target - p), d = target - o,
// t_max = 1 - SHADOW_EPSILON.
__global__ void __launch_bounds__(256) k_spawn_shadow(SceneView s, const float4* __restrict__ rays, const uint4* __restrict__ hits,
                                                       uint64_t n, float lx, float ly, float lz,
                                                       float4* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 h = hits[i];
    float4 o4 = make_float4(0.f, 0.f, 0.f, -1.0f), d4 = make_float4(0.f, 0.f, 1.f, 0.f);
    if (h.x != 0xFFFFFFFFu) {
        const float4 ro = __ldg(rays + 2 * i), rd = __ldg(rays + 2 * i + 1);
        SurfPoint sp;
        if (rebuild_hit(s, mk(ro.x, ro.y, ro.z), mk(rd.x, rd.y, rd.z), __ldg(s.slot_of_prim + h.x), &sp)) {
            const vec3 target = mk(lx, ly, lz);
            const vec3 d = target - sp.p;                       // interaction.rs:140: d = p2 - self.p
            const vec3 o = offset_ray_origin(sp.p, sp.p_err, sp.n, d);
            o4 = make_float4(o.x, o.y, o.z, 1.0f - PB2_SHADOW_EPS);
            d4 = make_float4(d.x, d.y, d.z, 0.0f);
        }
    }
    out[2 * i] = o4;
    out[2 * i + 1] = d4;
}

// (cosine_hemisphere: shade.cuh — sampling.rs:258-273, :289-294)

// interaction.rs:132-135 spawn_ray(d): o = offset_ray_origin(p, err, n, d), t_max = inf.  Direction: cosine-weighted
// about the geometric normal flipped towards the incoming side, frame from coordinate_system(n); PCG32 stream = ray index.
__global__ void __launch_bounds__(256) k_spawn_bounce(SceneView s, const float4* __restrict__ rays, const uint4* __restrict__ hits,
                                                       uint64_t n, float4* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 h = hits[i];
    float4 o4 = make_float4(0.f, 0.f, 0.f, -1.0f), d4 = make_float4(0.f, 0.f, 1.f, 0.f);
    if (h.x != 0xFFFFFFFFu) {
        const float4 ro = __ldg(rays + 2 * i), rd = __ldg(rays + 2 * i + 1);
        SurfPoint sp;
        const vec3 d_in = mk(rd.x, rd.y, rd.z);
        if (rebuild_hit(s, mk(ro.x, ro.y, ro.z), d_in, __ldg(s.slot_of_prim + h.x), &sp)) {
            vec3 n = sp.n;
            if (dot3(n, d_in) > 0.0f) n = -n;
            Pcg32 rng;
            rng.set_sequence(i);
            const float u0 = rng.next_float();
            const float u1 = rng.next_float();
            const vec3 l = cosine_hemisphere(u0, u1);
            vec3 sdir, tdir;
            coord_system(n, &sdir, &tdir);
            const vec3 wi = (sdir * l.x + tdir * l.y) + n * l.z;
            const vec3 o = offset_ray_origin(sp.p, sp.p_err, sp.n, wi);
            o4 = make_float4(o.x, o.y, o.z, __int_as_float(0x7f800000));
            d4 = make_float4(wi.x, wi.y, wi.z, 0.0f);
        }
    }
    out[2 * i] = o4;
    out[2 * i + 1] = d4;
}

// Both ray sets of the C3 workload from one pass over the hits: the hit's triangle is fetched (a scattered gather in a
// 10 M-triangle scene) and its interaction rebuilt once instead of once per builder; outputs equal the two kernels above.
__global__ void __launch_bounds__(256) k_spawn_shadow_bounce(SceneView s, const float4* __restrict__ rays, const uint4* __restrict__ hits,
                                                              uint64_t n, float lx, float ly, float lz, float4* __restrict__ out_shadow,
                                                              float4* __restrict__ out_bounce) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 h = hits[i];
    float4 so = make_float4(0.f, 0.f, 0.f, -1.0f), sd = make_float4(0.f, 0.f, 1.f, 0.f), bo = so, bd = sd;
    if (h.x != 0xFFFFFFFFu) {
        const float4 ro = __ldg(rays + 2 * i), rd = __ldg(rays + 2 * i + 1);
        SurfPoint sp;
        const vec3 d_in = mk(rd.x, rd.y, rd.z);
        if (rebuild_hit(s, mk(ro.x, ro.y, ro.z), d_in, __ldg(s.slot_of_prim + h.x), &sp)) {
            {
                const vec3 target = mk(lx, ly, lz);
                const vec3 d = target - sp.p;
                const vec3 o = offset_ray_origin(sp.p, sp.p_err, sp.n, d);
                so = make_float4(o.x, o.y, o.z, 1.0f - PB2_SHADOW_EPS);
                sd = make_float4(d.x, d.y, d.z, 0.0f);
            }
            vec3 nn = sp.n;
            if (dot3(nn, d_in) > 0.0f) nn = -nn;
            Pcg32 rng;
            rng.set_sequence(i);
            const float u0 = rng.next_float();
            const float u1 = rng.next_float();
            const vec3 l = cosine_hemisphere(u0, u1);
            vec3 sdir, tdir;
            coord_system(nn, &sdir, &tdir);
            const vec3 wi = (sdir * l.x + tdir * l.y) + nn * l.z;
            const vec3 o = offset_ray_origin(sp.p, sp.p_err, sp.n, wi);
            bo = make_float4(o.x, o.y, o.z, __int_as_float(0x7f800000));
            bd = make_float4(wi.x, wi.y, wi.z, 0.0f);
        }
    }
    out_shadow[2 * i] = so;
    out_shadow[2 * i + 1] = sd;
    out_bounce[2 * i] = bo;
    out_bounce[2 * i + 1] = bd;
}

void launch_spawn_shadow(const SceneView& s, const void* d_rays, const void* d_hits, uint64_t n, const float light[3],
                         void* d_out, cudaStream_t st) {
    if (n == 0) return;
    k_spawn_shadow<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s, (const float4*)d_rays, (const uint4*)d_hits, n, light[0], light[1],
                                                                  light[2], (float4*)d_out);
}

void launch_spawn_bounce(const SceneView& s, const void* d_rays, const void* d_hits, uint64_t n, void* d_out, cudaStream_t st) {
    if (n == 0) return;
    k_spawn_bounce<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s, (const float4*)d_rays, (const uint4*)d_hits, n, (float4*)d_out);
}

void launch_spawn_shadow_bounce(const SceneView& s, const void* d_rays, const void* d_hits, uint64_t n, const float light[3], void* d_out_shadow,
                                void* d_out_bounce, cudaStream_t st) {
    if (n == 0) return;
    k_spawn_shadow_bounce<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s, (const float4*)d_rays, (const uint4*)d_hits, n, light[0], light[1], light[2],
                                                                         (float4*)d_out_shadow, (float4*)d_out_bounce);
}

__global__ void __launch_bounds__(256) k_rng_floats(uint64_t first_seq, uint32_t n_seq, uint32_t n_per, float* __restrict__ out) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seq) return;
    Pcg32 rng;
    rng.set_sequence(first_seq + s);
    for (uint32_t k = 0; k < n_per; ++k) out[(uint64_t)s * n_per + k] = rng.next_float();
}

void launch_rng_floats(uint64_t first_seq, uint32_t n_seq, uint32_t n_per, float* d_out, cudaStream_t st) {
    if (n_seq == 0 || n_per == 0) return;
    k_rng_floats<<<(n_seq + 255) / 256, 256, 0, st>>>(first_seq, n_seq, n_per, d_out);
}

}  // namespace pb2
