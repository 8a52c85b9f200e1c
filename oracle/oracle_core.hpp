// ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement (C++17, scalar f32, built with -ffp-contract=off) of the pbrt-rs hot path:
// BVHAccel (SAH build, flatten, closest-hit / any-hit traversal), Triangle::intersect_test,
// the slab test, PCG32 / RandomSampler and the pinhole PerspectiveCamera.
//
// PARITY UNPINNED: the reference (/root/reference, lazytiger/pbrt-rs) cannot be compiled here
// (no rustc/cargo), cannot render as written and its tests hold no
// golden vector for this path.  Authority of this file = line-by-line correspondence with the
// reference files cited at each function + the KEEP/FIX ledger of SURVEY.md Appendix A
// (FIX = pbrt-v3 semantics where the Rust code panics or destroys the result).  The one
// externally pinned piece is PCG32 (canonical PCG demo vector, tests/test_oracle_core.py).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.  Nothing under pbrt-rs_b200/ includes or links it.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace orc {

using Float = float;

// ---------------------------------------------------------------- src/core/pbrt.rs:16-91
static constexpr Float kMachineEpsilon = 0x1p-24f;           // pbrt.rs:26-27 (0.5 * f32::EPSILON)
static constexpr Float kOneMinusEpsilon = 1.0f - 0x1p-23f;   // pbrt.rs:28 (D31 KEEP: 1 - EPSILON)
static constexpr Float kShadowEpsilon = 0.0001f;             // pbrt.rs:25
static constexpr Float kInfinity = std::numeric_limits<Float>::infinity();
static constexpr Float kPi = 3.14159265358979323846f;

// pbrt.rs:89-91: n * eps / (1 - n * eps), every op rounded to f32
inline Float gamma(Float n) { return (n * kMachineEpsilon) / (1.0f - n * kMachineEpsilon); }

inline uint32_t float_to_bits(Float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline Float bits_to_float(uint32_t u) { Float f; std::memcpy(&f, &u, 4); return f; }

// pbrt.rs:43-58
inline Float next_float_up(Float n) {
    if (std::isinf(n) && n > 0.0f) return n;
    if (n == -0.0f) n = 0.0f;
    uint32_t u = float_to_bits(n);
    if (n >= 0.0f) u += 1; else u -= 1;
    return bits_to_float(u);
}
// pbrt.rs:61-77
inline Float next_float_down(Float n) {
    if (std::isinf(n) && n < 0.0f) return n;
    if (n == 0.0f) n = -0.0f;
    uint32_t u = float_to_bits(n);
    if (n > 0.0f) u -= 1; else u += 1;
    return bits_to_float(u);
}

// Rust f32::min / f32::max: return the non-NaN operand (same as fminf/fmaxf). D1 FIX: max is max.
inline Float fmin_(Float a, Float b) { return std::fmin(a, b); }
inline Float fmax_(Float a, Float b) { return std::fmax(a, b); }

// ---------------------------------------------------------------- src/core/geometry.rs:62-314
struct V3 {
    Float x, y, z;
    Float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    Float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, Float s) { return {a.x * s, a.y * s, a.z * s}; }      // geometry.rs:236-244
inline V3 operator/(V3 a, Float s) { return {a.x / s, a.y / s, a.z / s}; }      // geometry.rs:268-276 (three divides)
inline Float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }     // geometry.rs:228-234
inline Float length_squared(V3 a) { return (a.x * a.x + a.y * a.y) + a.z * a.z; }
inline Float length(V3 a) { return std::sqrt(length_squared(a)); }
inline V3 normalize(V3 a) { return a / length(a); }                               // geometry.rs:117-119
inline V3 vabs(V3 a) { return {std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)}; }
inline V3 cross(V3 a, V3 b) {                                                     // geometry.rs:361-373
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// geometry.rs:37-51,91-93: x>y && x>z -> 0; else y>z -> 1; else 2
inline int max_dimension(V3 a) { return (a.x > a.y && a.x > a.z) ? 0 : ((a.y > a.z) ? 1 : 2); }
// geometry.rs:53-60,95-97: x.max(y.max(z)) (right nested); D1 FIX
inline Float max_component(V3 a) { return fmax_(a.x, fmax_(a.y, a.z)); }
inline V3 permute(V3 a, int kx, int ky, int kz) { return {a[kx], a[ky], a[kz]}; }
inline V3 vmin(V3 a, V3 b) { return {fmin_(a.x, b.x), fmin_(a.y, b.y), fmin_(a.z, b.z)}; }
inline V3 vmax(V3 a, V3 b) { return {fmax_(a.x, b.x), fmax_(a.y, b.y), fmax_(a.z, b.z)}; }
// geometry.rs:375-383 coordinate_system: v2 built then normalize()d (divide by length)
inline void coordinate_system(V3 v1, V3* v2, V3* v3) {
    if (std::fabs(v1.x) > std::fabs(v1.y))
        *v2 = normalize(V3{-v1.z, 0.0f, v1.x});
    else
        *v2 = normalize(V3{0.0f, v1.z, -v1.y});
    *v3 = cross(v1, *v2);
}

// ---------------------------------------------------------------- geometry.rs:430-588, 657-670
struct Bounds3 {
    V3 mn{std::numeric_limits<Float>::max(), std::numeric_limits<Float>::max(), std::numeric_limits<Float>::max()};
    V3 mx{-std::numeric_limits<Float>::max(), -std::numeric_limits<Float>::max(), -std::numeric_limits<Float>::max()};
    V3 diagonal() const { return mx - mn; }
    int maximum_extent() const { return max_dimension(diagonal()); }                 // :482-485
    Float surface_area() const {                                                     // :667-670
        V3 d = diagonal();
        return 2.0f * ((d.x * d.y + d.x * d.z) + d.y * d.z);
    }
    V3 offset(V3 p) const {                                                          // :460-467
        V3 o = p - mn;
        if (mx.x > mn.x) o.x /= mx.x - mn.x;
        if (mx.y > mn.y) o.y /= mx.y - mn.y;
        if (mx.z > mn.z) o.z /= mx.z - mn.z;
        return o;
    }
};
inline Bounds3 bunion(const Bounds3& a, const Bounds3& b) { return {vmin(a.mn, b.mn), vmax(a.mx, b.mx)}; }  // :523-530 (D1 FIX)
inline Bounds3 bunion(const Bounds3& a, V3 p) { return {vmin(a.mn, p), vmax(a.mx, p)}; }                     // :532-539

// ---------------------------------------------------------------- geometry.rs:756-763
struct Ray {
    V3 o;
    Float t_max;
    V3 d;
    Float time;
};
static_assert(sizeof(Ray) == 32, "ray is 32 bytes");

// geometry.rs:709-751 slab test. D2 FIX: z far plane uses 1 + 2*gamma(3) like x and y.
// Returns the boolean of the reference; *t_entry gets the final t_min (test harness only).
inline bool slab_test(const Bounds3& b, const Ray& ray, V3 inv_dir, const int dir_is_neg[3], Float* t_entry = nullptr) {
    const V3 bb[2] = {b.mn, b.mx};
    Float t_min = (bb[dir_is_neg[0]].x - ray.o.x) * inv_dir.x;
    Float t_max = (bb[1 - dir_is_neg[0]].x - ray.o.x) * inv_dir.x;
    Float ty_min = (bb[dir_is_neg[1]].y - ray.o.y) * inv_dir.y;
    Float ty_max = (bb[1 - dir_is_neg[1]].y - ray.o.y) * inv_dir.y;
    const Float widen = 1.0f + 2.0f * gamma(3.0f);
    t_max *= widen;
    ty_max *= widen;
    if (t_min > ty_max || ty_min > t_max) return false;
    if (ty_min > t_min) t_min = ty_min;
    if (ty_max < t_max) t_max = ty_max;
    Float tz_min = (bb[dir_is_neg[2]].z - ray.o.z) * inv_dir.z;
    Float tz_max = (bb[1 - dir_is_neg[2]].z - ray.o.z) * inv_dir.z;
    tz_max *= widen;
    if (t_min > tz_max || tz_min > t_max) return false;
    if (tz_min > t_min) t_min = tz_min;
    if (tz_max < t_max) t_max = tz_max;
    if (t_entry) *t_entry = t_min;
    return t_min < ray.t_max && t_max > 0.0f;
}

// ---------------------------------------------------------------- src/shapes/triangle.rs:74-158
struct TriHit {
    bool hit;
    Float b0, b1, b2, t;
};

// Watertight ray/triangle test. D7 FIX (sy), D8 FIX (range-test precedence), D9 KEEP (edge
// functions always in f64), D10 KEEP (delta_e uses delta_y twice).
inline TriHit triangle_intersect_test(V3 p0, V3 p1, V3 p2, const Ray& ray) {
    const TriHit miss{false, 0, 0, 0, 0};
    V3 p0t = p0 - ray.o;                                     // :80-82
    V3 p1t = p1 - ray.o;
    V3 p2t = p2 - ray.o;
    int kz = max_dimension(vabs(ray.d));                     // :84
    int kx = kz + 1; if (kx == 3) kx = 0;
    int ky = kx + 1; if (ky == 3) ky = 0;
    V3 d = permute(ray.d, kx, ky, kz);
    p0t = permute(p0t, kx, ky, kz);
    p1t = permute(p1t, kx, ky, kz);
    p2t = permute(p2t, kx, ky, kz);
    Float sx = -d.x / d.z;                                   // :99-101
    Float sy = -d.y / d.z;
    Float sz = 1.0f / d.z;
    p0t.x += sx * p0t.z;  p0t.y += sy * p0t.z;               // :102-107 (D7)
    p1t.x += sx * p1t.z;  p1t.y += sy * p1t.z;
    p2t.x += sx * p2t.z;  p2t.y += sy * p2t.z;
    // :109-111 f64 products are exact, one f64 subtract rounds, then f64 -> f32 rounds again
    Float e0 = (Float)((double)p1t.x * (double)p2t.y - (double)p1t.y * (double)p2t.x);
    Float e1 = (Float)((double)p2t.x * (double)p0t.y - (double)p2t.y * (double)p0t.x);
    Float e2 = (Float)((double)p0t.x * (double)p1t.y - (double)p0t.y * (double)p1t.x);
    if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return miss;
    Float det = (e0 + e1) + e2;
    if (det == 0.0f) return miss;
    p0t.z *= sz;  p1t.z *= sz;  p2t.z *= sz;                 // :122-124
    Float t_scaled = (e0 * p0t.z + e1 * p1t.z) + e2 * p2t.z; // :126
    if (det < 0.0f && (t_scaled >= 0.0f || t_scaled < ray.t_max * det)) return miss;        // D8
    else if (det > 0.0f && (t_scaled <= 0.0f || t_scaled > ray.t_max * det)) return miss;
    Float inv_det = 1.0f / det;                              // :133-137
    Float b0 = e0 * inv_det, b1 = e1 * inv_det, b2 = e2 * inv_det;
    Float t = t_scaled * inv_det;
    Float max_zt = max_component(vabs(V3{p0t.z, p1t.z, p2t.z}));                             // :139-154
    Float delta_z = gamma(3.0f) * max_zt;
    Float max_xt = max_component(vabs(V3{p0t.x, p1t.x, p2t.x}));
    Float max_yt = max_component(vabs(V3{p0t.y, p1t.y, p2t.y}));
    Float delta_y = gamma(5.0f) * (max_yt + max_zt);
    Float delta_e = 2.0f * ((gamma(2.0f) * max_xt * max_yt + delta_y * max_xt) + delta_y * max_yt);   // D10
    Float max_e = max_component(vabs(V3{e0, e1, e2}));
    Float delta_t = 3.0f * ((gamma(3.0f) * max_e * max_zt + delta_e * max_zt) + delta_z * max_e) * std::fabs(inv_det);
    if (t <= delta_t) return miss;
    return {true, b0, b1, b2, t};
}

// triangle.rs:182-215: Triangle::intersect rejects an accepted candidate when dpdu x dpdv == 0 and
// the geometric normal is degenerate too.  uv = {u0, v0, u1, v1, u2, v2} of the mesh (Triangle::get_uvs, :60-72), or null for
// the default UVs (0,0),(1,0),(1,1) (determinant == 1).
inline bool triangle_frame(V3 p0, V3 p1, V3 p2, V3* dpdu, V3* dpdv, const Float* uv = nullptr) {
    V3 dp02 = p0 - p2, dp12 = p1 - p2;
    const Float duv[6] = {0.0f, 0.0f, 1.0f, 0.0f, 1.0f, 1.0f};
    if (!uv) uv = duv;
    const Float duv02[2] = {uv[0] - uv[4], uv[1] - uv[5]};
    const Float duv12[2] = {uv[2] - uv[4], uv[3] - uv[5]};
    Float determinant = duv02[0] * duv12[1] - duv02[1] * duv12[0];
    bool degenerate_uv = std::fabs(determinant) < 1e-8f;      // D11 FIX
    V3 du{0, 0, 0}, dv{0, 0, 0};
    if (!degenerate_uv) {
        Float inv_det = 1.0f / determinant;
        du = (dp02 * duv12[1] - dp12 * duv02[1]) * inv_det;
        dv = (dp02 * -duv12[0] + dp12 * duv02[0]) * inv_det;
    }
    if (degenerate_uv || length_squared(cross(du, dv)) == 0.0f) {
        V3 ng = cross(p2 - p0, p1 - p0);
        if (length_squared(ng) == 0.0f) return false;
        coordinate_system(normalize(ng), &du, &dv);
    }
    *dpdu = du;
    *dpdv = dv;
    return true;
}

#include "oracle_sphere.hpp"

// ---------------------------------------------------------------- src/accelerators/bvh.rs
struct LinearBVHNode {            // bvh.rs:129-135 (usize fields narrowed; D19)
    Bounds3 bounds;
    uint32_t primitive_or_second_child_offset;
    uint16_t n_primitives;
    uint8_t axis;
    uint8_t pad;
};
static_assert(sizeof(LinearBVHNode) == 32, "32-byte node");

struct BVHPrimitiveInfo {         // bvh.rs:26-41
    uint32_t primitive_number;
    Bounds3 bounds;
    V3 centroid;
};

struct BuildNode {                // bvh.rs:43-51
    Bounds3 bounds;
    int children[2] = {-1, -1};
    int split_axis = 0;
    uint32_t first_prim_offset = 0;
    uint32_t n_primitives = 0;
};

struct TraversalCounters {
    uint64_t nodes_tested = 0;    // boxes slab-tested, reference order
    uint64_t tris_tested = 0;     // Triangle::intersect_test calls, reference order
};

struct Hit {
    uint32_t prim_id;             // index into the caller's triangle list; 0xFFFFFFFF = miss
    Float t, b1, b2;
};

class BVHAccel {
public:
    std::vector<V3> verts;
    std::vector<uint32_t> indices;          // 3 per triangle, mesh order
    std::vector<uint32_t> ordered_prims;    // BVH leaf order -> mesh triangle id
    std::vector<LinearBVHNode> nodes;
    std::vector<Float> uvs;                 // TriangleMesh::uv (triangle.rs:21): 2 per vertex, empty = default UVs
    std::vector<Sphere> spheres;            // analytic spheres (shapes/sphere.rs): primitive ids n_tris() .. n_tris() + spheres.size() - 1; set before build()
    size_t n_tris() const { return indices.size() / 3; }
    bool is_sphere(uint32_t prim) const { return prim >= n_tris(); }
    const Sphere& sphere(uint32_t prim) const { return spheres[prim - n_tris()]; }
    int max_prims_in_node = 4;
    int max_depth_seen = 0;

    // Triangle::get_uvs (triangle.rs:60-72): the mesh's, or null for the defaults
    const Float* tri_uv(uint32_t prim, Float out[6]) const {
        if (uvs.empty()) return nullptr;
        for (int k = 0; k < 3; ++k) { out[2 * k] = uvs[2 * indices[3 * prim + k]]; out[2 * k + 1] = uvs[2 * indices[3 * prim + k] + 1]; }
        return out;
    }
    void tri(uint32_t prim, V3* p0, V3* p1, V3* p2) const {
        *p0 = verts[indices[3 * prim]];
        *p1 = verts[indices[3 * prim + 1]];
        *p2 = verts[indices[3 * prim + 2]];
    }
    Bounds3 tri_bound(uint32_t prim) const {                 // triangle.rs:175-180
        V3 p0, p1, p2;
        tri(prim, &p0, &p1, &p2);
        Bounds3 b{vmin(p0, p1), vmax(p0, p1)};
        return bunion(b, p2);
    }

    // bvh.rs:216-271
    void build(const Float* v, size_t nv, const uint32_t* idx, size_t nt, int max_prims, int split_method = 0) {
        verts.resize(nv);
        for (size_t i = 0; i < nv; ++i) verts[i] = {v[3 * i], v[3 * i + 1], v[3 * i + 2]};
        indices.assign(idx, idx + 3 * nt);
        max_prims_in_node = std::min(max_prims, 255);        // :222
        nodes.clear();
        ordered_prims.clear();
        const size_t n_mesh = nt;
        nt += spheres.size();                                 // the primitive list: the mesh's triangles, then the spheres
        if (nt == 0) return;
        std::vector<BVHPrimitiveInfo> info(nt);
        for (size_t i = 0; i < nt; ++i) {
            Bounds3 b = i < n_mesh ? tri_bound((uint32_t)i) : spheres[i - n_mesh].world_bound();
            info[i] = {(uint32_t)i, b, b.mn * 0.5f + b.mx * 0.5f};       // :38
        }
        build_nodes_.clear();
        build_nodes_.reserve(2 * nt);
        ordered_prims.reserve(nt);
        int root;
        split_method_ = split_method;
        if (split_method == 1) {                              // SplitMethod::HLBVH (:239-243)
            ordered_prims.resize(nt);
            root = hlbvh_build(info);
        } else {
            root = recursive_build(info, 0, nt);
        }
        nodes.resize(build_nodes_.size());
        uint32_t offset = 0;
        max_depth_seen = 0;
        flatten(root, &offset, 1);
        build_nodes_.clear();
        build_nodes_.shrink_to_fit();
    }

    Bounds3 world_bound() const { return nodes.empty() ? Bounds3{} : nodes[0].bounds; }   // :819-826

    // bvh.rs:828-879 + primitive.rs:65-78 + triangle.rs:182-215
    // ssi_out: the interaction of the accepted hit when that hit is a sphere (Sphere::intersect builds it at accept time)
    bool intersect(Ray& ray, Hit* out, Float* b0_out, TraversalCounters* c, SphereSI* ssi_out = nullptr) const {
        out->prim_id = 0xFFFFFFFFu; out->t = ray.t_max; out->b1 = 0; out->b2 = 0;
        if (b0_out) *b0_out = 0;
        if (nodes.empty()) return false;
        bool hit = false;
        V3 inv_dir{1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z};
        int dir_is_neg[3] = {inv_dir.x < 0.0f, inv_dir.y < 0.0f, inv_dir.z < 0.0f};
        int to_visit = 0;
        uint32_t current = 0;
        uint32_t stack[64];
        for (;;) {
            const LinearBVHNode& node = nodes[current];
            if (c) c->nodes_tested++;
            if (slab_test(node.bounds, ray, inv_dir, dir_is_neg)) {
                if (node.n_primitives > 0) {
                    for (uint32_t i = 0; i < node.n_primitives; ++i) {
                        uint32_t prim = ordered_prims[node.primitive_or_second_child_offset + i];
                        if (c) c->tris_tested++;
                        if (is_sphere(prim)) {                // GeometricPrimitive(Sphere)::intersect (primitive.rs:65-78, sphere.rs:38-93)
                            Float t;
                            SphereSI ssi;
                            if (!sphere(prim).intersect(ray, &t, &ssi)) continue;
                            ray.t_max = t;
                            out->prim_id = prim; out->t = t; out->b1 = ssi.u; out->b2 = ssi.v;
                            if (b0_out) *b0_out = 0.0f;
                            if (ssi_out) *ssi_out = ssi;
                            hit = true;
                            continue;
                        }
                        V3 p0, p1, p2;
                        tri(prim, &p0, &p1, &p2);
                        TriHit h = triangle_intersect_test(p0, p1, p2, ray);
                        if (!h.hit) continue;
                        V3 du, dv;
                        Float uvb[6];
                        if (!triangle_frame(p0, p1, p2, &du, &dv, tri_uv(prim, uvb))) continue;
                        ray.t_max = h.t;                      // primitive.rs:70
                        out->prim_id = prim; out->t = h.t; out->b1 = h.b1; out->b2 = h.b2;
                        if (b0_out) *b0_out = h.b0;
                        hit = true;
                    }
                    if (to_visit == 0) break;
                    current = stack[--to_visit];
                } else if (dir_is_neg[node.axis]) {
                    stack[to_visit++] = current + 1;
                    current = node.primitive_or_second_child_offset;
                } else {
                    stack[to_visit++] = node.primitive_or_second_child_offset;
                    current = current + 1;
                }
            } else {
                if (to_visit == 0) break;
                current = stack[--to_visit];
            }
        }
        return hit;
    }

    // bvh.rs:881-932 + triangle.rs:318-321
    bool intersect_p(const Ray& ray, TraversalCounters* c) const {
        if (nodes.empty()) return false;
        V3 inv_dir{1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z};
        int dir_is_neg[3] = {inv_dir.x < 0.0f, inv_dir.y < 0.0f, inv_dir.z < 0.0f};
        int to_visit = 0;
        uint32_t current = 0;
        uint32_t stack[64];
        for (;;) {
            const LinearBVHNode& node = nodes[current];
            if (c) c->nodes_tested++;
            if (slab_test(node.bounds, ray, inv_dir, dir_is_neg)) {
                if (node.n_primitives > 0) {
                    for (uint32_t i = 0; i < node.n_primitives; ++i) {
                        uint32_t prim = ordered_prims[node.primitive_or_second_child_offset + i];
                        if (c) c->tris_tested++;
                        if (is_sphere(prim)) {
                            if (sphere(prim).intersect_p(ray)) return true;
                            continue;
                        }
                        V3 p0, p1, p2;
                        tri(prim, &p0, &p1, &p2);
                        if (triangle_intersect_test(p0, p1, p2, ray).hit) return true;
                    }
                    if (to_visit == 0) break;
                    current = stack[--to_visit];
                } else if (dir_is_neg[node.axis]) {
                    stack[to_visit++] = current + 1;
                    current = node.primitive_or_second_child_offset;
                } else {
                    stack[to_visit++] = node.primitive_or_second_child_offset;
                    current = current + 1;
                }
            } else {
                if (to_visit == 0) break;
                current = stack[--to_visit];
            }
        }
        return false;
    }

    // Checker: loop over every triangle in BVH leaf order (SURVEY §4.3 property test).
    bool brute_force(Ray& ray, Hit* out) const {
        out->prim_id = 0xFFFFFFFFu; out->t = ray.t_max; out->b1 = 0; out->b2 = 0;
        bool hit = false;
        for (uint32_t prim : ordered_prims) {
            if (is_sphere(prim)) {
                Float t;
                SphereSI ssi;
                if (!sphere(prim).intersect(ray, &t, &ssi)) continue;
                ray.t_max = t;
                out->prim_id = prim; out->t = t; out->b1 = ssi.u; out->b2 = ssi.v;
                hit = true;
                continue;
            }
            V3 p0, p1, p2;
            tri(prim, &p0, &p1, &p2);
            TriHit h = triangle_intersect_test(p0, p1, p2, ray);
            if (!h.hit) continue;
            V3 du, dv;
            Float uvb[6];
            if (!triangle_frame(p0, p1, p2, &du, &dv, tri_uv(prim, uvb))) continue;
            ray.t_max = h.t;
            out->prim_id = prim; out->t = h.t; out->b1 = h.b1; out->b2 = h.b2;
            hit = true;
        }
        return hit;
    }

private:
    std::vector<BuildNode> build_nodes_;

    int make_leaf(int index, std::vector<BVHPrimitiveInfo>& info, size_t start, size_t end, const Bounds3& bounds) {
        BuildNode& n = build_nodes_[index];
        n.first_prim_offset = (uint32_t)ordered_prims.size();
        n.n_primitives = (uint32_t)(end - start);
        n.bounds = bounds;
        for (size_t i = start; i < end; ++i) ordered_prims.push_back(info[i].primitive_number);
        return index;
    }

    // bvh.rs:273-473: SplitMethod::SAH (0), ::Middle (2), ::EqualCounts (3), with D12-D17 fixed (pbrt-v3 semantics).
    // Middle as written (:331-349) partitions [start, end] and, when the partition is improper, sorts about a SHADOWED `mid`
    // while recursing on the improper one (unbounded recursion); pbrt-v3 falls through to EqualCounts, followed here (D60 FIX).
    int split_method_ = 0;
    int recursive_build(std::vector<BVHPrimitiveInfo>& info, size_t start, size_t end) {
        int index = (int)build_nodes_.size();
        build_nodes_.emplace_back();
        Bounds3 bounds;
        for (size_t i = start; i < end; ++i) bounds = bunion(bounds, info[i].bounds);
        size_t n_primitives = end - start;
        if (n_primitives == 1) return make_leaf(index, info, start, end, bounds);
        Bounds3 centroid_bounds;
        for (size_t i = start; i < end; ++i) centroid_bounds = bunion(centroid_bounds, info[i].centroid);   // D12
        int dim = centroid_bounds.maximum_extent();
        size_t mid = (start + end) / 2;
        if (centroid_bounds.mx[dim] == centroid_bounds.mn[dim]) return make_leaf(index, info, start, end, bounds);
        bool equal_counts = split_method_ == 3;
        if (split_method_ == 2) {                                                                         // :331-349 SplitMethod::Middle
            const Float p_mid = (centroid_bounds.mn[dim] + centroid_bounds.mx[dim]) / 2.0f;
            size_t lo = start, hi = end;                                                                  // partition_in_place, as below
            for (;;) {
                while (lo < hi && info[lo].centroid[dim] < p_mid) ++lo;
                if (lo == hi) break;
                do { --hi; } while (lo < hi && !(info[hi].centroid[dim] < p_mid));
                if (lo == hi) break;
                std::swap(info[lo], info[hi]);
                ++lo;
            }
            mid = lo;
            if (mid == start || mid == end) equal_counts = true;
        }
        if (split_method_ == 2 && !equal_counts) {
        } else if (equal_counts || n_primitives <= 2) {
            // :350-371 (D17): nth_element on [start,end) by centroid[dim]
            mid = (start + end) / 2;
            std::nth_element(info.begin() + start, info.begin() + mid, info.begin() + end,
                             [dim](const BVHPrimitiveInfo& a, const BVHPrimitiveInfo& b) { return a.centroid[dim] < b.centroid[dim]; });
        } else {
            constexpr int n_buckets = 12;
            struct Bucket { int count = 0; Bounds3 bounds; } buckets[n_buckets];
            auto bucket_of = [&](const BVHPrimitiveInfo& pi) {
                int b = (int)((Float)n_buckets * centroid_bounds.offset(pi.centroid)[dim]);                    // D14
                if (b == n_buckets) b = n_buckets - 1;
                return b;
            };
            for (size_t i = start; i < end; ++i) {
                int b = bucket_of(info[i]);
                buckets[b].count++;
                buckets[b].bounds = bunion(buckets[b].bounds, info[i].bounds);                                 // D12
            }
            Float cost[n_buckets - 1];
            for (int i = 0; i < n_buckets - 1; ++i) {
                Bounds3 b0, b1;
                int count0 = 0, count1 = 0;
                for (int j = 0; j <= i; ++j) { b0 = bunion(b0, buckets[j].bounds); count0 += buckets[j].count; }   // D15
                for (int j = i + 1; j < n_buckets; ++j) { b1 = bunion(b1, buckets[j].bounds); count1 += buckets[j].count; }
                cost[i] = 1.0f + ((Float)count0 * b0.surface_area() + (Float)count1 * b1.surface_area()) / bounds.surface_area();
            }
            Float min_cost = std::numeric_limits<Float>::max();
            int min_cost_split_bucket = 0;
            for (int i = 0; i < n_buckets - 1; ++i)
                if (cost[i] < min_cost) { min_cost = cost[i]; min_cost_split_bucket = i; }
            Float leaf_cost = (Float)n_primitives;
            if ((int)n_primitives > max_prims_in_node || min_cost < leaf_cost) {
                // :422-432 partition_in_place (Rust core::iter): find first false from the front, last true
                // from the back, swap.  D13 ([start,end)), D16 (b <= split bucket).
                size_t lo = start, hi = end;
                for (;;) {
                    while (lo < hi && bucket_of(info[lo]) <= min_cost_split_bucket) ++lo;
                    if (lo == hi) break;
                    do { --hi; } while (lo < hi && !(bucket_of(info[hi]) <= min_cost_split_bucket));
                    if (lo == hi) break;
                    std::swap(info[lo], info[hi]);
                    ++lo;
                }
                mid = lo;
            } else {
                return make_leaf(index, info, start, end, bounds);
            }
        }
        int c0 = recursive_build(info, start, mid);
        int c1 = recursive_build(info, mid, end);
        BuildNode& n = build_nodes_[index];                   // :53-92 init_interior
        n.bounds = bunion(build_nodes_[c0].bounds, build_nodes_[c1].bounds);
        n.children[0] = c0;
        n.children[1] = c1;
        n.split_axis = dim;
        n.n_primitives = 0;
        return index;
    }

    // ---- SplitMethod::HLBVH: bvh.rs:137-197 (Morton codes, radix sort), :475-568 hlbvh_build, :570-676 emit_lbvh,
    // :678-772 build_upper_sah.  The Rust port does not run as written (H1-H6 below); this follows pbrt-v3
    // (accelerators/bvh.cpp HLBVHBuild / emitLBVH / buildUpperSAH), which the reference declares itself a port of:
    //   H1 :482-485 the centroid bound discards Bounds3::union's result          -> FIX (accumulate)
    //   H2 :508     `start` is never advanced after a treelet is emitted          -> FIX (start = end)
    //   H3 :595     leaves are written into an empty ordered_prims Vec            -> FIX (sized up front)
    //   H4 :723-731,:755-763  `n_buckets * (x as usize)` casts before multiplying -> FIX ((n_buckets * x) as usize)
    //   H5 :752     partitions [start, end]                                       -> FIX ([start, end))
    //   H6 :765     keeps b < split bucket                                        -> FIX (b <= split bucket, pbrt-v3)
    //   H7 treelets are emitted sequentially in Morton order, so a leaf's first_prim_offset is its position in the sorted
    //      array (pbrt-v3 hands offsets out with an atomic; any order is a valid run of it)
    //   H8 pbrt-v3 CHECKs that the upper-level centroid extent is non-zero and that the partition is proper; here a
    //      zero extent or an improper partition splits the range in the middle instead of aborting
    struct MortonPrimitive { uint32_t primitive_index, morton_code; };
    static uint32_t left_shift3(uint32_t x) {                 // :138-152
        if (x == (1u << 10)) x -= 1;
        x = (x | (x << 16)) & 0b00000011000000000000000011111111u;
        x = (x | (x << 8)) & 0b00000011000000001111000000001111u;
        x = (x | (x << 4)) & 0b00000011000011000011000011000011u;
        x = (x | (x << 2)) & 0b00001001001001001001001001001001u;
        return x;
    }
    static uint32_t f2u_sat(Float v) {                        // Rust `as u32`: NaN -> 0, saturating
        if (!(v > 0.0f)) return 0u;
        if (v >= 4294967296.0f) return 0xFFFFFFFFu;
        return (uint32_t)v;
    }
    static uint32_t encode_morton3(V3 v) {                    // :154-157
        return (left_shift3(f2u_sat(v.z)) << 2) | (left_shift3(f2u_sat(v.y)) << 1) | left_shift3(f2u_sat(v.x));
    }
    static void radix_sort(std::vector<MortonPrimitive>& v) { // :159-196
        const int bits_per_pass = 6, n_bits = 30, n_passes = n_bits / bits_per_pass;
        std::vector<MortonPrimitive> temp(v.size());
        for (int pass = 0; pass < n_passes; ++pass) {
            const int low_bit = pass * bits_per_pass;
            std::vector<MortonPrimitive>& in = (pass & 1) ? temp : v;
            std::vector<MortonPrimitive>& out = (pass & 1) ? v : temp;
            const int n_buckets = 1 << bits_per_pass;
            const uint32_t bit_mask = (1u << bits_per_pass) - 1u;
            size_t bucket_count[64] = {0}, out_index[64];
            for (const MortonPrimitive& mp : in) bucket_count[(mp.morton_code >> low_bit) & bit_mask]++;
            out_index[0] = 0;
            for (int i = 1; i < n_buckets; ++i) out_index[i] = out_index[i - 1] + bucket_count[i - 1];
            for (const MortonPrimitive& mp : in) out[out_index[(mp.morton_code >> low_bit) & bit_mask]++] = mp;
        }
        if (n_passes & 1) std::swap(v, temp);
    }

    int hlbvh_build(const std::vector<BVHPrimitiveInfo>& info) {
        Bounds3 bounds;
        for (const BVHPrimitiveInfo& pi : info) bounds = bunion(bounds, pi.centroid);                            // H1
        std::vector<MortonPrimitive> morton(info.size());
        for (size_t i = 0; i < info.size(); ++i) {
            const int morton_bits = 10, morton_scale = 1 << morton_bits;
            const V3 centroid_offset = bounds.offset(info[i].centroid);
            morton[i].primitive_index = info[i].primitive_number;
            morton[i].morton_code = encode_morton3(centroid_offset * (Float)morton_scale);
        }
        radix_sort(morton);
        std::vector<int> roots;
        size_t ordered_offset = 0;
        for (size_t start = 0, end = 1; end <= morton.size(); ++end) {
            const uint32_t mask = 0b00111111111111000000000000000000u;
            if (end == morton.size() || (morton[start].morton_code & mask) != (morton[end].morton_code & mask)) {
                const int first_bit_index = 29 - 12;
                roots.push_back(emit_lbvh(info, morton.data() + start, end - start, &ordered_offset, first_bit_index));
                start = end;                                                                                     // H2
            }
        }
        return build_upper_sah(roots, 0, roots.size());
    }

    int emit_lbvh(const std::vector<BVHPrimitiveInfo>& info, const MortonPrimitive* mp, size_t n_primitives, size_t* ordered_offset,
                  int bit_index) {
        if (bit_index == -1 || n_primitives < (size_t)max_prims_in_node) {                                       // :583
            const int index = (int)build_nodes_.size();
            build_nodes_.emplace_back();
            Bounds3 bounds;
            const size_t first = *ordered_offset;
            *ordered_offset += n_primitives;                                                                     // H7
            for (size_t i = 0; i < n_primitives; ++i) {
                ordered_prims[first + i] = mp[i].primitive_index;                                                // H3
                bounds = bunion(bounds, info[mp[i].primitive_index].bounds);
            }
            BuildNode& n = build_nodes_[index];
            n.first_prim_offset = (uint32_t)first;
            n.n_primitives = (uint32_t)n_primitives;
            n.bounds = bounds;
            return index;
        }
        const uint32_t mask = 1u << bit_index;
        if ((mp[0].morton_code & mask) == (mp[n_primitives - 1].morton_code & mask))                              // :608-625
            return emit_lbvh(info, mp, n_primitives, ordered_offset, bit_index - 1);
        size_t search_start = 0, search_end = n_primitives - 1;
        while (search_start + 1 != search_end) {                                                                 // :627-638
            const size_t mid = (search_start + search_end) / 2;
            if ((mp[search_start].morton_code & mask) == (mp[mid].morton_code & mask)) search_start = mid;
            else search_end = mid;
        }
        const size_t split_offset = search_end;
        const int index = (int)build_nodes_.size();
        build_nodes_.emplace_back();
        const int c0 = emit_lbvh(info, mp, split_offset, ordered_offset, bit_index - 1);
        const int c1 = emit_lbvh(info, mp + split_offset, n_primitives - split_offset, ordered_offset, bit_index - 1);
        BuildNode& n = build_nodes_[index];
        n.bounds = bunion(build_nodes_[c0].bounds, build_nodes_[c1].bounds);
        n.children[0] = c0;
        n.children[1] = c1;
        n.split_axis = bit_index % 3;                                                                            // :671
        n.n_primitives = 0;
        return index;
    }

    int build_upper_sah(std::vector<int>& roots, size_t start, size_t end) {
        const size_t n_nodes = end - start;
        if (n_nodes == 1) return roots[start];
        const int index = (int)build_nodes_.size();
        build_nodes_.emplace_back();
        Bounds3 bounds, centroid_bounds;
        for (size_t i = start; i < end; ++i) bounds = bunion(bounds, build_nodes_[roots[i]].bounds);
        for (size_t i = start; i < end; ++i) {
            const Bounds3& b = build_nodes_[roots[i]].bounds;
            centroid_bounds = bunion(centroid_bounds, (b.mn + b.mx) * 0.5f);                                     // :703
        }
        const int dim = centroid_bounds.maximum_extent();
        size_t mid = (start + end) / 2;
        if (centroid_bounds.mx[dim] != centroid_bounds.mn[dim]) {                                                // H8
            constexpr int n_buckets = 12;
            struct Bucket { int count = 0; Bounds3 bounds; } buckets[n_buckets];
            auto bucket_of = [&](int root) {
                const Bounds3& nb = build_nodes_[root].bounds;
                const Float centroid = (nb.mn[dim] + nb.mx[dim]) * 0.5f;
                int b = (int)((Float)n_buckets * ((centroid - centroid_bounds.mn[dim]) / (centroid_bounds.mx[dim] - centroid_bounds.mn[dim])));   // H4
                if (b == n_buckets) b = n_buckets - 1;
                return b;
            };
            for (size_t i = start; i < end; ++i) {
                const int b = bucket_of(roots[i]);
                buckets[b].count++;
                buckets[b].bounds = bunion(buckets[b].bounds, build_nodes_[roots[i]].bounds);
            }
            Float cost[n_buckets - 1];
            for (int i = 0; i < n_buckets - 1; ++i) {
                Bounds3 b0, b1;
                int count0 = 0, count1 = 0;
                for (int j = 0; j <= i; ++j) { b0 = bunion(b0, buckets[j].bounds); count0 += buckets[j].count; }
                for (int j = i + 1; j < n_buckets; ++j) { b1 = bunion(b1, buckets[j].bounds); count1 += buckets[j].count; }
                cost[i] = 0.125f + ((Float)count0 * b0.surface_area() + (Float)count1 * b1.surface_area()) / bounds.surface_area();   // :736
            }
            Float min_cost = std::numeric_limits<Float>::max();
            int min_cost_split_bucket = 0;
            for (int i = 0; i < n_buckets - 1; ++i)
                if (cost[i] < min_cost) { min_cost = cost[i]; min_cost_split_bucket = i; }
            size_t lo = start, hi = end;                                                                         // partition_in_place, H5 H6
            for (;;) {
                while (lo < hi && bucket_of(roots[lo]) <= min_cost_split_bucket) ++lo;
                if (lo == hi) break;
                do { --hi; } while (lo < hi && !(bucket_of(roots[hi]) <= min_cost_split_bucket));
                if (lo == hi) break;
                std::swap(roots[lo], roots[hi]);
                ++lo;
            }
            if (lo > start && lo < end) mid = lo;                                                                // H8
        }
        const int c0 = build_upper_sah(roots, start, mid);
        const int c1 = build_upper_sah(roots, mid, end);
        BuildNode& n = build_nodes_[index];
        n.bounds = bunion(build_nodes_[c0].bounds, build_nodes_[c1].bounds);
        n.children[0] = c0;
        n.children[1] = c1;
        n.split_axis = dim;
        n.n_primitives = 0;
        return index;
    }

    // bvh.rs:774-811
    uint32_t flatten(int index, uint32_t* offset, int depth) {
        const BuildNode& node = build_nodes_[index];
        LinearBVHNode& ln = nodes[*offset];
        std::memset((void*)&ln, 0, sizeof ln);
        ln.bounds = node.bounds;
        uint32_t my_offset = (*offset)++;
        if (depth > max_depth_seen) max_depth_seen = depth;
        if (node.n_primitives > 0) {
            ln.primitive_or_second_child_offset = node.first_prim_offset;
            ln.n_primitives = (uint16_t)node.n_primitives;
        } else {
            ln.axis = (uint8_t)node.split_axis;
            ln.n_primitives = 0;
            flatten(node.children[0], offset, depth + 1);
            nodes[my_offset].primitive_or_second_child_offset = flatten(node.children[1], offset, depth + 1);
        }
        return my_offset;
    }
};

// ---------------------------------------------------------------- src/core/rng.rs:14-69
struct RNG {
    uint64_t state = 0x853c49e6748fea9bULL;   // PCG32_DEFAULT_STATE
    uint64_t inc = 0xda3e39cb94b95bdbULL;     // PCG32_DEFAULT_STREAM
    static constexpr uint64_t kMult = 0x5851f42d4c957f2dULL;
    // rng.rs:21-27.  init_state parameter exposes the canonical PCG seeding (reference hard-codes the default).
    void set_sequence(uint64_t sequence_index, uint64_t init_state = 0x853c49e6748fea9bULL) {
        state = 0;
        inc = (sequence_index << 1) | 1;
        uniform_u32();
        state += init_state;
        uniform_u32();
    }
    uint32_t uniform_u32() {                   // rng.rs:29-35 (D32: wrapping u64)
        uint64_t old = state;
        state = old * kMult + inc;
        uint32_t xs = (uint32_t)(((old >> 18) ^ old) >> 27);
        uint32_t rot = (uint32_t)(old >> 59);
        return (xs >> rot) | (xs << ((~rot + 1u) & 31));
    }
    Float uniform_float() {                    // rng.rs:46-48
        return fmin_(kOneMinusEpsilon, (Float)uniform_u32() * 2.3283064365386963e-10f);
    }
};

// ---------------------------------------------------------------- sin / cos
// The reference calls f32::sin / f32::cos (platform libm; last-bit results are platform dependent).  The numerics
// contract of this project fixes one definition shared by oracle and kernels: Cody-Waite reduction by
// pi/2 in three f32 steps + Cephes sinf/cosf minimax polynomials, every op separately rounded (<= 2 ulp).
inline void sincos_contract(Float x, Float* s_out, Float* c_out) {
    const Float q = std::rint(x * 0.636619772367581343f);
    const int k = (int)q;
    Float r = x - q * 1.5703125f;
    r = r - q * 4.837512969970703125e-4f;
    r = r - q * 7.549789948768648e-8f;
    const Float z = r * r;
    const Float sp = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * r + r;
    const Float cp = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z - 0.5f * z + 1.0f;
    switch (k & 3) {
        case 0: *s_out = sp; *c_out = cp; break;
        case 1: *s_out = cp; *c_out = -sp; break;
        case 2: *s_out = -sp; *c_out = -cp; break;
        default: *s_out = -cp; *c_out = sp; break;
    }
}
inline Float sin_c(Float x) { Float s, c; sincos_contract(x, &s, &c); return s; }
inline Float cos_c(Float x) { Float s, c; sincos_contract(x, &s, &c); return c; }

// ---------------------------------------------------------------- geometry.rs:1139-1154, interaction.rs:132-153
inline V3 offset_ray_origin(V3 p, V3 p_error, V3 n, V3 w) {
    Float d = dot(vabs(n), p_error);
    V3 offset = n * d;
    if (dot(w, n) < 0.0f) offset = -offset;
    V3 po = p + offset;
    for (int i = 0; i < 3; ++i) {
        if (offset[i] > 0.0f) po[i] = next_float_up(po[i]);
        else if (offset[i] < 0.0f) po[i] = next_float_down(po[i]);
    }
    return po;
}
struct Interaction {          // interaction.rs:100-123 BaseInteraction subset
    V3 p, error, n;
};
inline Ray spawn_ray(const Interaction& it, V3 d) {                       // :132-135
    return Ray{offset_ray_origin(it.p, it.error, it.n, d), kInfinity, d, 0.0f};
}
inline Ray spawn_ray_to(const Interaction& it, V3 p2) {                   // :138-144
    V3 d = p2 - it.p;
    return Ray{offset_ray_origin(it.p, it.error, it.n, d), 1.0f - kShadowEpsilon, d, 0.0f};
}
// triangle.rs:217-250: p_hit, p_error, geometric normal of an accepted hit
inline Interaction triangle_interaction(V3 p0, V3 p1, V3 p2, Float b0, Float b1, Float b2) {
    Interaction it;
    Float xs = (std::fabs(b0 * p0.x) + std::fabs(b1 * p1.x)) + std::fabs(b2 * p2.x);
    Float ys = (std::fabs(b0 * p0.y) + std::fabs(b1 * p1.y)) + std::fabs(b2 * p2.y);
    Float zs = (std::fabs(b0 * p0.z) + std::fabs(b1 * p1.z)) + std::fabs(b2 * p2.z);
    it.error = V3{xs, ys, zs} * gamma(7.0f);
    it.p = (p0 * b0 + p1 * b1) + p2 * b2;
    it.n = normalize(cross(p0 - p2, p1 - p2));
    return it;
}

// ---------------------------------------------------------------- sampling.rs:258-273, :289-294
inline void concentric_sample_disk(Float u0, Float u1, Float* x, Float* y) {
    Float ox = u0 * 2.0f - 1.0f, oy = u1 * 2.0f - 1.0f;
    if (ox == 0.0f && oy == 0.0f) { *x = 0; *y = 0; return; }
    Float r, theta;
    if (std::fabs(ox) > std::fabs(oy)) { r = ox; theta = (kPi / 4.0f) * (oy / ox); }
    else { r = oy; theta = (kPi / 2.0f) - (kPi / 4.0f) * (ox / oy); }
    *x = cos_c(theta) * r;
    *y = sin_c(theta) * r;
}
inline V3 cosine_sample_hemisphere(Float u0, Float u1) {                  // D30 FIX: sqrt
    Float x, y;
    concentric_sample_disk(u0, u1, &x, &y);
    Float z = std::sqrt(fmax_(0.0f, (1.0f - x * x) - y * y));
    return {x, y, z};
}

// ---------------------------------------------------------------- transform.rs + cameras/perspective.rs
struct M4 { Float m[4][4]; };
inline M4 m4_identity() { M4 r{}; for (int i = 0; i < 4; ++i) r.m[i][i] = 1.0f; return r; }
inline M4 m4_mul(const M4& a, const M4& b) {                                          // transform.rs:167-182
    M4 r{};
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            r.m[i][j] = ((a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j]) + a.m[i][2] * b.m[2][j]) + a.m[i][3] * b.m[3][j];
    return r;
}
// transform.rs:46-113 Gauss-Jordan with full pivoting.  D4 FIX: start from the matrix itself.
inline M4 m4_inverse(const M4& in) {
    int indxc[4], indxr[4], ipiv[4] = {0, 0, 0, 0};
    Float minv[4][4];
    std::memcpy(minv, in.m, sizeof minv);
    for (int i = 0; i < 4; i++) {
        int irow = 0, icol = 0;
        Float big = 0.0f;
        for (int j = 0; j < 4; j++) {
            if (ipiv[j] != 1) {
                for (int k = 0; k < 4; k++) {
                    if (ipiv[k] == 0) {
                        if (std::fabs(minv[j][k]) >= big) { big = std::fabs(minv[j][k]); irow = j; icol = k; }
                    }
                }
            }
        }
        ++ipiv[icol];
        if (irow != icol) for (int k = 0; k < 4; ++k) std::swap(minv[irow][k], minv[icol][k]);
        indxr[i] = irow;
        indxc[i] = icol;
        Float pivinv = 1.0f / minv[icol][icol];
        minv[icol][icol] = 1.0f;
        for (int j = 0; j < 4; j++) minv[icol][j] *= pivinv;
        for (int j = 0; j < 4; j++) {
            if (j != icol) {
                Float save = minv[j][icol];
                minv[j][icol] = 0.0f;
                for (int k = 0; k < 4; k++) minv[j][k] -= minv[icol][k] * save;
            }
        }
    }
    for (int j = 3; j >= 0; j--) {
        if (indxr[j] != indxc[j]) for (int k = 0; k < 4; k++) std::swap(minv[k][indxr[j]], minv[k][indxc[j]]);
    }
    M4 r;
    std::memcpy(r.m, minv, sizeof minv);
    return r;
}
struct Transform { M4 m, m_inv; };
inline Transform t_mul(const Transform& a, const Transform& b) {                     // transform.rs:609-617 (D5 FIX)
    return {m4_mul(a.m, b.m), m4_mul(b.m_inv, a.m_inv)};
}
inline Transform t_inverse(const Transform& a) { return {a.m_inv, a.m}; }            // transform.rs:208-213
inline Transform t_translate(V3 d) {                                                 // transform.rs:408-425
    Transform t{m4_identity(), m4_identity()};
    t.m.m[0][3] = d.x; t.m.m[1][3] = d.y; t.m.m[2][3] = d.z;
    t.m_inv.m[0][3] = -d.x; t.m_inv.m[1][3] = -d.y; t.m_inv.m[2][3] = -d.z;
    return t;
}
inline Transform t_scale(Float x, Float y, Float z) {                                // transform.rs:426-443 (m FIXED to a diagonal)
    Transform t{m4_identity(), m4_identity()};
    t.m.m[0][0] = x; t.m.m[1][1] = y; t.m.m[2][2] = z;
    t.m_inv.m[0][0] = 1.0f / x; t.m_inv.m[1][1] = 1.0f / y; t.m_inv.m[2][2] = 1.0f / z;
    return t;
}
inline Transform t_perspective(Float fov, Float n, Float f) {                        // transform.rs:555-566
    M4 persp = m4_identity();
    persp.m[2][2] = f / (f - n);
    persp.m[2][3] = -f * n / (f - n);
    persp.m[3][2] = 1.0f;
    persp.m[3][3] = 0.0f;
    Float inv_tan_ang = 1.0f / std::tan((kPi / 180.0f * fov) / 2.0f);
    Transform tp{persp, m4_inverse(persp)};
    return t_mul(t_scale(inv_tan_ang, inv_tan_ang, 1.0f), tp);
}
inline Transform t_look_at(V3 pos, V3 look, V3 up) {                                 // transform.rs:510-541
    M4 c2w = m4_identity();
    c2w.m[0][3] = pos.x; c2w.m[1][3] = pos.y; c2w.m[2][3] = pos.z; c2w.m[3][3] = 1.0f;
    V3 dir = normalize(look - pos);
    V3 right = normalize(cross(normalize(up), dir));
    V3 newup = cross(dir, right);
    c2w.m[0][0] = right.x; c2w.m[1][0] = right.y; c2w.m[2][0] = right.z; c2w.m[3][0] = 0.0f;
    c2w.m[0][1] = newup.x; c2w.m[1][1] = newup.y; c2w.m[2][1] = newup.z; c2w.m[3][1] = 0.0f;
    c2w.m[0][2] = dir.x; c2w.m[1][2] = dir.y; c2w.m[2][2] = dir.z; c2w.m[3][2] = 0.0f;
    return {m4_inverse(c2w), c2w};   // world -> camera
}
inline V3 t_point(const M4& m, V3 p) {                                               // transform.rs:351-369
    Float xp = ((m.m[0][0] * p.x + m.m[0][1] * p.y) + m.m[0][2] * p.z) + m.m[0][3];
    Float yp = ((m.m[1][0] * p.x + m.m[1][1] * p.y) + m.m[1][2] * p.z) + m.m[1][3];
    Float zp = ((m.m[2][0] * p.x + m.m[2][1] * p.y) + m.m[2][2] * p.z) + m.m[2][3];
    Float wp = ((m.m[3][0] * p.x + m.m[3][1] * p.y) + m.m[3][2] * p.z) + m.m[3][3];
    if (wp == 1.0f) return {xp, yp, zp};
    return V3{xp, yp, zp} / wp;
}
inline V3 t_point_err(const M4& m, V3 p, V3* err) {                                  // geometry.rs:898-936
    Float x = p.x, y = p.y, z = p.z;
    Float xp = ((m.m[0][0] * x + m.m[0][1] * y) + m.m[0][2] * z) + m.m[0][3];
    Float yp = ((m.m[1][0] * x + m.m[1][1] * y) + m.m[1][2] * z) + m.m[1][3];
    Float zp = ((m.m[2][0] * x + m.m[2][1] * y) + m.m[2][2] * z) + m.m[2][3];
    Float wp = ((m.m[3][0] * x + m.m[3][1] * y) + m.m[3][2] * z) + m.m[3][3];
    Float xs = ((std::fabs(m.m[0][0] * x) + std::fabs(m.m[0][1] * y)) + std::fabs(m.m[0][2] * z)) + std::fabs(m.m[0][3]);
    Float ys = ((std::fabs(m.m[1][0] * x) + std::fabs(m.m[1][1] * y)) + std::fabs(m.m[1][2] * z)) + std::fabs(m.m[1][3]);
    Float zs = ((std::fabs(m.m[2][0] * x) + std::fabs(m.m[2][1] * y)) + std::fabs(m.m[2][2] * z)) + std::fabs(m.m[2][3]);
    *err = V3{xs, ys, zs} * gamma(3.0f);
    if (wp == 1.0f) return {xp, yp, zp};
    return V3{xp, yp, zp} / wp;
}
inline V3 t_vector(const M4& m, V3 v) {                                              // transform.rs:371-386
    return {(m.m[0][0] * v.x + m.m[0][1] * v.y) + m.m[0][2] * v.z,
            (m.m[1][0] * v.x + m.m[1][1] * v.y) + m.m[1][2] * v.z,
            (m.m[2][0] * v.x + m.m[2][1] * v.y) + m.m[2][2] * v.z};
}

struct Camera {                    // cameras/perspective.rs:34-82
    M4 raster_to_camera;
    M4 camera_to_world;
    int res_x, res_y;
    Float lens_radius = 0.0f, focal_distance = 1e6f;        // perspective.rs:26-27 (pbrt's defaults)
    void init(V3 pos, V3 look, V3 up, Float fov, int rx, int ry) {
        res_x = rx; res_y = ry;
        // pbrt-v3 api.cpp default screen window: the shorter axis spans [-1,1]
        Float frame = (Float)rx / (Float)ry;
        Float sw_min_x, sw_max_x, sw_min_y, sw_max_y;
        if (frame > 1.0f) { sw_min_x = -frame; sw_max_x = frame; sw_min_y = -1.0f; sw_max_y = 1.0f; }
        else { sw_min_x = -1.0f; sw_max_x = 1.0f; sw_min_y = -1.0f / frame; sw_max_y = 1.0f / frame; }
        Transform camera_to_screen = t_perspective(fov, 1e-2f, 1000.0f);
        Transform screen_to_raster =
            t_mul(t_mul(t_scale((Float)rx, (Float)ry, 1.0f),
                        t_scale(1.0f / (sw_max_x - sw_min_x), 1.0f / (sw_min_y - sw_max_y), 1.0f)),
                  t_translate(V3{-sw_min_x, -sw_max_y, 0.0f}));
        Transform raster_to_screen = t_inverse(screen_to_raster);
        Transform r2c = t_mul(t_inverse(camera_to_screen), raster_to_screen);
        raster_to_camera = r2c.m;
        camera_to_world = t_look_at(pos, look, up).m_inv;
    }
    // perspective.rs:90-112 + geometry.rs:865-881; (lx, ly) = CameraSample::p_lens, used by the thin lens only
    Ray generate_ray(Float fx, Float fy, Float lx = 0.0f, Float ly = 0.0f) const {
        V3 p_camera = t_point(raster_to_camera, V3{fx, fy, 0.0f});
        V3 d_cam = normalize(p_camera);
        V3 o_cam{0, 0, 0};
        if (lens_radius > 0.0f) {                                                   // :101-107
            Float px, py;
            concentric_sample_disk(lx, ly, &px, &py);
            px = px * lens_radius; py = py * lens_radius;
            const Float ft = focal_distance / d_cam.z;
            const V3 p_focus = o_cam + d_cam * ft;                                  // Ray::point, geometry.rs:790-792
            o_cam = V3{px, py, 0.0f};
            d_cam = normalize(p_focus - o_cam);
        }
        V3 o_error;
        V3 o = t_point_err(camera_to_world, o_cam, &o_error);
        V3 d = t_vector(camera_to_world, d_cam);
        Float ls = length_squared(d);
        Float t_max = kInfinity;
        if (ls > 0.0f) {
            Float dt = dot(vabs(d), o_error) / ls;
            o = o + d * dt;
            t_max -= dt;
        }
        return Ray{o, t_max, d, 0.0f};
    }
};

}  // namespace orc
