// traverse.cuh — device-side BVHAccel traversal + watertight triangle test (sm_100a).
//
// Replaces, per ray: BVHAccel::intersect / intersect_p (src/accelerators/bvh.rs:828-879, :881-932),
// Bounds3f::intersect_p (src/core/geometry.rs:709-751), GeometricPrimitive::intersect (src/core/primitive.rs:65-78)
// and Triangle::intersect_test (src/shapes/triangle.rs:74-158).
//
// Exactness: results (primitive id, t, barycentrics, any-hit boolean) are bit-identical to the reference
// traversal because (1) every float op is a separately rounded IEEE op in the reference's order (-fmad=false),
// (2) triangles of a leaf are tested in leaf order and children are visited near-first by dir_is_neg[axis],
// exactly as bvh.rs:856-866, and (3) the child-pair node layout tests the far child's box early but keeps its
// entry distance on the stack and re-applies the `t_min < ray.t_max` clause (geometry.rs:749) at pop time with
// the then-current t_max — the only clause of the slab test that depends on ray.t_max — so the set and order of
// triangles tested equals the reference's.
#pragma once
#include "pb2_math.cuh"
#include "sphere.cuh"

namespace pb2 {

constexpr uint32_t kLeafFlag = 0x80000000u;
constexpr int kStackDepth = 64;     // bvh.rs:839

struct SceneView {
    const float4* __restrict__ quads;   // 8 x float4 per record (QuadNode): two tree levels per fetch — the kernels' layout
    uint32_t quad_root_ref;
    const float4* __restrict__ pairs;   // 4 x float4 per interior node (PairNode): literal one-level walk (traverse())
    const float4* __restrict__ tris;    // 3 x float4 per triangle (PackedTri), BVH leaf order
    const uint32_t* __restrict__ slot_of_prim;   // caller's triangle id -> leaf-order slot
    const float4* __restrict__ tris_prim; // shading scenes only, 6 x float4 per primitive in the CALLER's primitive order: the PackedTri and
                                          // the plain triangle's n, ss, ts (k_tris_by_prim) — a path vertex is rebuilt from hit.prim directly
    const float4* __restrict__ spheres;  // 8 x float4 per analytic sphere (DSphere, sphere.cuh); null when the scene has none
    uint32_t root_ref;
    uint32_t n_tris;                     // primitives in the tree (triangles + spheres)
    float root_lo[3];
    float root_hi[3];
};

struct HitRec {
    uint32_t prim;      // caller's triangle id
    uint32_t slot;      // position in leaf order (index into tris)
    float t, b0, b1, b2;
    bool sphere;        // the hit is an analytic sphere: b0 = 0, (b1, b2) = (u, v)
};

// Per-ray constants: inverse direction, slab selectors and the shear of the watertight test.
struct RayCtx {
    vec3 o, d, inv;
    bool nx, ny, nz;
    int kx, ky, kz;
    float sx, sy, sz;
};

PB2_D RayCtx make_ray_ctx(vec3 o, vec3 d) {
    RayCtx r;
    r.o = o;
    r.d = d;
    r.inv = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);             // bvh.rs:831
    r.nx = r.inv.x < 0.0f;                                      // bvh.rs:832-836
    r.ny = r.inv.y < 0.0f;
    r.nz = r.inv.z < 0.0f;
    r.kz = max_dim(abs3(d));                                    // triangle.rs:84-92
    r.kx = r.kz + 1; if (r.kx == 3) r.kx = 0;
    r.ky = r.kx + 1; if (r.ky == 3) r.ky = 0;
    const float dx = comp(d, r.kx), dy = comp(d, r.ky), dz = comp(d, r.kz);
    r.sx = -dx / dz;                                            // triangle.rs:99-101
    r.sy = -dy / dz;
    r.sz = 1.0f / dz;
    return r;
}

// geometry.rs:709-751 without its final `t_min < ray.t_max` clause: returns the t_max-independent verdict and the
// entry distance.  D2 FIX: the z far plane is widened by 1 + 2*gamma(3) like x and y.
PB2_D bool slab_entry(const RayCtx& r, float lx, float ly, float lz, float hx, float hy, float hz, float* t_entry) {
    const float widen = 1.0f + 2.0f * gammaf_(3.0f);            // folds to 0x3F800003
    float t_min = ((r.nx ? hx : lx) - r.o.x) * r.inv.x;
    float t_max = ((r.nx ? lx : hx) - r.o.x) * r.inv.x;
    const float ty_min = ((r.ny ? hy : ly) - r.o.y) * r.inv.y;
    float ty_max = ((r.ny ? ly : hy) - r.o.y) * r.inv.y;
    t_max = t_max * widen;
    ty_max = ty_max * widen;
    bool ok = !(t_min > ty_max || ty_min > t_max);
    if (ty_min > t_min) t_min = ty_min;
    if (ty_max < t_max) t_max = ty_max;
    const float tz_min = ((r.nz ? hz : lz) - r.o.z) * r.inv.z;
    float tz_max = ((r.nz ? lz : hz) - r.o.z) * r.inv.z;
    tz_max = tz_max * widen;
    ok = ok && !(t_min > tz_max || tz_min > t_max);
    if (tz_min > t_min) t_min = tz_min;
    if (tz_max < t_max) t_max = tz_max;
    *t_entry = t_min;
    return ok && (t_max > 0.0f);
}

PB2_D vec3 permute3(vec3 v, int kx, int ky, int kz) { return mk(comp(v, kx), comp(v, ky), comp(v, kz)); }
// permute3(v, kx, ky, kz) for the triple Triangle::intersect_test builds (triangle.rs:84-92: kz = max dimension of |d|,
// kx = kz + 1, ky = kx + 1, both mod 3): always a rotation, selected by two predicates the caller evaluates once — (y, z, x) for
// kz = 0, (z, x, y) for kz = 1, v itself for kz = 2.  Six selects per vertex; the three index-driven comp() calls cost k_extend
// 46 instructions per triangle test, 10 % of its warp instructions (profiles/r02_tuning.md).
PB2_D vec3 rotate3(vec3 v, bool kz0, bool kz1) {
    return mk(kz0 ? v.y : (kz1 ? v.z : v.x), kz0 ? v.z : (kz1 ? v.x : v.y), kz0 ? v.x : (kz1 ? v.y : v.z));
}

// (f32)((f64)a*(f64)b - (f64)c*(f64)d): both products are exact in binary64, so one fused multiply-subtract
// rounds exactly like the reference's f64 subtract (triangle.rs:109-111); then one f64 -> f32 rounding.
PB2_D float edge_fn(float a, float b, float c, float d) {
    return __double2float_rn(__fma_rn((double)a, (double)b, -__dmul_rn((double)c, (double)d)));
}

// Triangle::intersect_test (triangle.rs:74-158; D7, D8 fixed; D9, D10 kept).
PB2_D bool tri_test(const RayCtx& r, float ray_t_max, vec3 p0, vec3 p1, vec3 p2, float* t_out, float* b0, float* b1, float* b2) {
    const bool kz0 = r.kz == 0, kz1 = r.kz == 1;
    vec3 p0t = rotate3(p0 - r.o, kz0, kz1);
    vec3 p1t = rotate3(p1 - r.o, kz0, kz1);
    vec3 p2t = rotate3(p2 - r.o, kz0, kz1);
    p0t.x = p0t.x + r.sx * p0t.z;  p0t.y = p0t.y + r.sy * p0t.z;
    p1t.x = p1t.x + r.sx * p1t.z;  p1t.y = p1t.y + r.sy * p1t.z;
    p2t.x = p2t.x + r.sx * p2t.z;  p2t.y = p2t.y + r.sy * p2t.z;
    const float e0 = edge_fn(p1t.x, p2t.y, p1t.y, p2t.x);
    const float e1 = edge_fn(p2t.x, p0t.y, p2t.y, p0t.x);
    const float e2 = edge_fn(p0t.x, p1t.y, p0t.y, p1t.x);
    if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return false;
    const float det = (e0 + e1) + e2;
    if (det == 0.0f) return false;
    p0t.z = p0t.z * r.sz;
    p1t.z = p1t.z * r.sz;
    p2t.z = p2t.z * r.sz;
    const float t_scaled = (e0 * p0t.z + e1 * p1t.z) + e2 * p2t.z;
    if (det < 0.0f && (t_scaled >= 0.0f || t_scaled < ray_t_max * det)) return false;
    else if (det > 0.0f && (t_scaled <= 0.0f || t_scaled > ray_t_max * det)) return false;
    const float inv_det = 1.0f / det;
    const float t = t_scaled * inv_det;
    const float max_zt = max3(fabsf(p0t.z), fabsf(p1t.z), fabsf(p2t.z));
    const float delta_z = gammaf_(3.0f) * max_zt;
    const float max_xt = max3(fabsf(p0t.x), fabsf(p1t.x), fabsf(p2t.x));
    const float max_yt = max3(fabsf(p0t.y), fabsf(p1t.y), fabsf(p2t.y));
    const float delta_y = gammaf_(5.0f) * (max_yt + max_zt);
    const float delta_e = 2.0f * ((gammaf_(2.0f) * max_xt * max_yt + delta_y * max_xt) + delta_y * max_yt);
    const float max_e = max3(fabsf(e0), fabsf(e1), fabsf(e2));
    const float delta_t = 3.0f * ((gammaf_(3.0f) * max_e * max_zt + delta_e * max_zt) + delta_z * max_e) * fabsf(inv_det);
    if (t <= delta_t) return false;
    *t_out = t;
    *b0 = e0 * inv_det;
    *b1 = e1 * inv_det;
    *b2 = e2 * inv_det;
    return true;
}

// triangle.rs:193-215: dpdu, dpdv from the triangle's UVs (Triangle::get_uvs, :60-72; uv0/uv1/uv2, default (0,0),(1,0),(1,1));
// false when Triangle::intersect bails out on a degenerate frame (closest-hit only — intersect_p never runs this).
PB2_HD bool tri_frame_uv(vec3 p0, vec3 p1, vec3 p2, float2 uv0, float2 uv1, float2 uv2, vec3* dpdu, vec3* dpdv) {
    const vec3 dp02 = p0 - p2, dp12 = p1 - p2;
    const float duv02x = uv0.x - uv2.x, duv02y = uv0.y - uv2.y, duv12x = uv1.x - uv2.x, duv12y = uv1.y - uv2.y;
    const float determinant = duv02x * duv12y - duv02y * duv12x;
    const bool degenerate_uv = fabsf(determinant) < 1e-8f;               // D11 FIX
    vec3 du = mk(0.f, 0.f, 0.f), dv = mk(0.f, 0.f, 0.f);
    if (!degenerate_uv) {
        const float inv_det = 1.0f / determinant;
        du = (dp02 * duv12y - dp12 * duv02y) * inv_det;
        dv = (dp02 * -duv12x + dp12 * duv02x) * inv_det;
    }
    if (degenerate_uv || len2(cross3(du, dv)) == 0.0f) {
        const vec3 ng = cross3(p2 - p0, p1 - p0);
        if (len2(ng) == 0.0f) return false;
        coord_system(unit(ng), &du, &dv);
    }
    *dpdu = du;
    *dpdv = dv;
    return true;
}
PB2_HD bool tri_frame(vec3 p0, vec3 p1, vec3 p2, vec3* dpdu, vec3* dpdv) {
    const vec3 dp02 = p0 - p2, dp12 = p1 - p2;
    const float duv02x = 0.0f - 1.0f, duv02y = 0.0f - 1.0f, duv12x = 1.0f - 1.0f, duv12y = 0.0f - 1.0f;
    const float determinant = duv02x * duv12y - duv02y * duv12x;
    const float inv_det = 1.0f / determinant;
    vec3 du = (dp02 * duv12y - dp12 * duv02y) * inv_det;
    vec3 dv = (dp02 * -duv12x + dp12 * duv02x) * inv_det;
    if (len2(cross3(du, dv)) == 0.0f) {
        const vec3 ng = cross3(p2 - p0, p1 - p0);
        if (len2(ng) == 0.0f) return false;
        coord_system(unit(ng), &du, &dv);
    }
    *dpdu = du;
    *dpdv = dv;
    return true;
}

PB2_D float4 ldg4(const float4* p) { return __ldg(p); }
// One 256-bit read-only load (sm_100: LDG.E.256) of two consecutive float4 at a 32-byte aligned address: half the L1
// wavefronts and load instructions of two 128-bit loads.
PB2_D void ldg8(const float4* p, float4* a, float4* b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a->x), "=f"(a->y), "=f"(a->z), "=f"(a->w), "=f"(b->x), "=f"(b->y), "=f"(b->z), "=f"(b->w)
                 : "l"(p));
}

// One ray through the BVH.  ANY = intersect_p semantics (first accepted triangle ends the walk).
// A sphere's leaf slot (bvh_build.hpp kPrimSphere): c.w bit 1 set, a.x = bits of its index in SceneView::spheres.
PB2_D const DSphere& sphere_of(const SceneView& s, float4 a) { return reinterpret_cast<const DSphere*>(s.spheres)[__float_as_uint(a.x)]; }

template <bool ANY, bool SPH = false>
PB2_D bool traverse(const SceneView& s, vec3 o, vec3 d, float ray_t_max, HitRec* hit) {
    if (s.n_tris == 0) return false;
    const RayCtx r = make_ray_ctx(o, d);
    float t_max = ray_t_max;
    float te;
    if (!(slab_entry(r, s.root_lo[0], s.root_lo[1], s.root_lo[2], s.root_hi[0], s.root_hi[1], s.root_hi[2], &te) && te < t_max))
        return false;
    uint32_t stack_ref[kStackDepth];
    float stack_t[kStackDepth];
    int sp = 0;
    uint32_t cur = s.root_ref;
    bool found = false;
    for (;;) {
        if (cur & kLeafFlag) {
            uint32_t slot = cur & ~kLeafFlag;
            for (;;) {
                const float4 a = ldg4(s.tris + 3ull * slot);
                const float4 b = ldg4(s.tris + 3ull * slot + 1);
                const float4 c = ldg4(s.tris + 3ull * slot + 2);
                const vec3 p0 = mk(a.x, a.y, a.z), p1 = mk(b.x, b.y, b.z), p2 = mk(c.x, c.y, c.z);
                float t, b0, b1, b2;
                if (SPH && (__float_as_uint(c.w) & 2u)) {       // GeometricPrimitive(Sphere) (primitive.rs:65-78, sphere.rs:38-98)
                    vec3 p_hit, od;
                    float phi;
                    if (sphere_test(sphere_of(s, a), o, d, t_max, &p_hit, &phi, &od, &t)) {
                        if (ANY) return true;
                        const SphereVertex sv = sphere_vertex(sphere_of(s, a), p_hit, phi, od);
                        t_max = t;
                        hit->prim = __float_as_uint(a.w);
                        hit->slot = slot;
                        hit->t = t; hit->b0 = 0.0f; hit->b1 = sv.u; hit->b2 = sv.v;
                        hit->sphere = true;
                        found = true;
                    }
                } else
                if (tri_test(r, t_max, p0, p1, p2, &t, &b0, &b1, &b2)) {
                    if (ANY) return true;
                    if (__float_as_uint(c.w) == 0u) {          // frame not degenerate (k_mark_degenerate, triangle.rs:193-215)
                        t_max = t;                              // primitive.rs:70
                        hit->prim = __float_as_uint(a.w);
                        hit->slot = slot;
                        hit->t = t; hit->b0 = b0; hit->b1 = b1; hit->b2 = b2;
                        hit->sphere = false;
                        found = true;
                    }
                }
                if (__float_as_uint(b.w) != 0u) break;          // last triangle of the leaf
                ++slot;
            }
        } else {
            const float4* np = s.pairs + 4ull * cur;
            const float4 a = ldg4(np), b = ldg4(np + 1), c = ldg4(np + 2);
            const uint4 m = __ldg(reinterpret_cast<const uint4*>(np + 3));
            float tl, tr;
            const bool okl = slab_entry(r, a.x, a.y, a.z, a.w, b.x, b.y, &tl) && (tl < t_max);
            const bool okr = slab_entry(r, b.z, b.w, c.x, c.y, c.z, c.w, &tr) && (tr < t_max);
            // bvh.rs:856-866: near child first, chosen by the sign of the direction on the split axis
            const bool neg = (m.z == 0u) ? r.nx : ((m.z == 1u) ? r.ny : r.nz);
            const uint32_t near_ref = neg ? m.y : m.x, far_ref = neg ? m.x : m.y;
            const bool ok_near = neg ? okr : okl, ok_far = neg ? okl : okr;
            const float t_far = neg ? tl : tr;
            if (ok_near) {
                if (ok_far) { stack_ref[sp] = far_ref; stack_t[sp] = t_far; ++sp; }
                cur = near_ref;
                continue;
            }
            if (ok_far) { cur = far_ref; continue; }            // far visited immediately: t_max unchanged since its test
        }
        // pop: re-apply `t_min < ray.t_max` with the current t_max (geometry.rs:749)
        bool got = false;
        while (sp > 0) {
            --sp;
            if (stack_t[sp] < t_max) { cur = stack_ref[sp]; got = true; break; }
        }
        if (!got) break;
    }
    return found;
}

}  // namespace pb2
