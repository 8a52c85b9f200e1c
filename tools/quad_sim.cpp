// tools/quad_sim.cpp — development harness (not part of the product or of the test suite's product path): walks the
// QuadNode layout built by pbrt-rs_b200/csrc/bvh_build.cpp on the CPU with the kernel's rules (children ordered by the
// three split axes, entry distance re-checked at pop) and compares every ray with the reference-order binary walk over
// the same LinearNode array.  Uses the oracle's slab / triangle arithmetic.  Build: see tools/quad_sim.py.
#include <cstdint>
#include <cstdio>
#include <vector>
#include <thread>
#include <atomic>

#include "../oracle/oracle_core.hpp"
#include "../pbrt-rs_b200/csrc/bvh_build.hpp"

using namespace orc;

struct SimOut { uint32_t prim; float t; };

static bool plain_ray(const Ray& r, V3 inv) {
    auto fin = [](float v) { return v != 0.0f && std::fabs(v) < kInfinity; };
    return fin(inv.x) && fin(inv.y) && fin(inv.z);
}

// reference-order walk over LinearNode (closest hit), triangles from PackedTri in leaf order
static SimOut walk_binary(const pb2::HostBVH& b, Ray ray, uint64_t* nodes, uint64_t* tris) {
    SimOut o{0xFFFFFFFFu, ray.t_max};
    V3 inv{1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z};
    int neg[3] = {inv.x < 0.0f, inv.y < 0.0f, inv.z < 0.0f};
    uint32_t stack[64]; int sp = 0; uint32_t cur = 0;
    if (b.nodes.empty()) return o;
    for (;;) {
        const pb2::LinearNode& n = b.nodes[cur];
        ++*nodes;
        Bounds3 bb{{n.bmin[0], n.bmin[1], n.bmin[2]}, {n.bmax[0], n.bmax[1], n.bmax[2]}};
        if (slab_test(bb, ray, inv, neg)) {
            if (n.n_prims > 0) {
                for (uint32_t i = 0; i < n.n_prims; ++i) {
                    const pb2::PackedTri& t = b.tris[n.offset + i];
                    V3 p0{t.v0[0], t.v0[1], t.v0[2]}, p1{t.v1[0], t.v1[1], t.v1[2]}, p2{t.v2[0], t.v2[1], t.v2[2]};
                    ++*tris;
                    TriHit h = triangle_intersect_test(p0, p1, p2, ray);
                    if (!h.hit) continue;
                    V3 du, dv;
                    if (!triangle_frame(p0, p1, p2, &du, &dv)) continue;
                    ray.t_max = h.t; o.prim = t.prim_id; o.t = h.t;
                }
                if (!sp) break;
                cur = stack[--sp];
            } else if (neg[n.axis]) { stack[sp++] = cur + 1; cur = n.offset; }
            else { stack[sp++] = n.offset; cur = cur + 1; }
        } else { if (!sp) break; cur = stack[--sp]; }
    }
    return o;
}

static SimOut walk_quad(const pb2::HostBVH& b, Ray ray, uint64_t* steps, uint64_t* boxes, uint64_t* tris, int* max_sp) {
    SimOut o{0xFFFFFFFFu, ray.t_max};
    V3 inv{1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z};
    int neg[3] = {inv.x < 0.0f, inv.y < 0.0f, inv.z < 0.0f};
    if (b.nodes.empty()) return o;
    {   // root box
        const pb2::LinearNode& n = b.nodes[0];
        Bounds3 bb{{n.bmin[0], n.bmin[1], n.bmin[2]}, {n.bmax[0], n.bmax[1], n.bmax[2]}};
        if (!slab_test(bb, ray, inv, neg)) return o;
    }
    struct E { uint32_t ref; float t; };
    E stack[96]; int sp = 0;
    uint32_t cur = b.quad_root_ref;
    for (;;) {
        if (cur & pb2::kLeafBit) {
            uint32_t slot = cur & ~pb2::kLeafBit;
            for (;;) {
                const pb2::PackedTri& t = b.tris[slot];
                V3 p0{t.v0[0], t.v0[1], t.v0[2]}, p1{t.v1[0], t.v1[1], t.v1[2]}, p2{t.v2[0], t.v2[1], t.v2[2]};
                ++*tris;
                TriHit h = triangle_intersect_test(p0, p1, p2, ray);
                if (h.hit) {
                    V3 du, dv;
                    if (triangle_frame(p0, p1, p2, &du, &dv)) { ray.t_max = h.t; o.prim = t.prim_id; o.t = h.t; }
                }
                if (t.last) break;
                ++slot;
            }
        } else {
            const pb2::QuadNode& q = b.quads[cur];
            ++*steps;
            E e[4];
            for (int k = 0; k < 4; ++k) {
                e[k].ref = q.ref[k] == pb2::kQuadEmpty ? q.ref[k] : (q.ref[k] & pb2::kQuadRefMask);
                e[k].t = kInfinity;
                if (q.ref[k] == pb2::kQuadEmpty) continue;
                ++*boxes;
                Bounds3 bb{{q.lox[k], q.loy[k], q.loz[k]}, {q.hix[k], q.hiy[k], q.hiz[k]}};
                Float te;
                if (slab_test(bb, ray, inv, neg, &te)) e[k].t = te;      // includes te < t_max
            }
            const int aP = (q.ref[0] >> pb2::kQuadAxisShift) & 3, aA = (q.ref[1] >> pb2::kQuadAxisShift) & 3, aB = (q.ref[2] >> pb2::kQuadAxisShift) & 3;
            if (neg[aA]) std::swap(e[0], e[1]);
            if (neg[aB]) std::swap(e[2], e[3]);
            if (neg[aP]) { std::swap(e[0], e[2]); std::swap(e[1], e[3]); }
            int first = -1;
            for (int k = 0; k < 4; ++k) if (e[k].t < kInfinity) { first = k; break; }
            if (first >= 0) {
                for (int k = 3; k > first; --k) if (e[k].t < kInfinity) stack[sp++] = e[k];
                if (sp > *max_sp) *max_sp = sp;
                cur = e[first].ref;
                continue;
            }
        }
        bool got = false;
        while (sp > 0) { --sp; if (stack[sp].t < ray.t_max) { cur = stack[sp].ref; got = true; break; } }
        if (!got) break;
    }
    return o;
}


// Generic k-level collapse on the binary LinearNode tree (no layout built): one "step" at interior node X tests the boxes of X's
// descendant frontier `levels` levels down (a leaf met earlier stays in the frontier), in the reference's visiting order for this
// ray's signs, visits the first accepted one and pushes the others with their entry distances (re-checked at pop).  levels = 2
// is the shipped QuadNode walk; levels = 3 is the 8-wide record of the round-1 verdict's lever (ii).  Per-ray counters only.
struct WideStats { uint64_t steps = 0, boxes = 0, tris = 0, pushes = 0, pops = 0, leaf_visits = 0; };
static void frontier(const pb2::HostBVH& b, uint32_t node, int levels, const int* neg, uint32_t* out, int* n_out) {
    const pb2::LinearNode& n = b.nodes[node];
    if (levels == 0 || n.n_prims > 0) { out[(*n_out)++] = node; return; }
    const uint32_t first = neg[n.axis] ? n.offset : node + 1, second = neg[n.axis] ? node + 1 : n.offset;
    frontier(b, first, levels - 1, neg, out, n_out);
    frontier(b, second, levels - 1, neg, out, n_out);
}
static SimOut walk_wide(const pb2::HostBVH& b, Ray ray, int levels, WideStats* st) {
    SimOut o{0xFFFFFFFFu, ray.t_max};
    V3 inv{1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z};
    int neg[3] = {inv.x < 0.0f, inv.y < 0.0f, inv.z < 0.0f};
    if (b.nodes.empty()) return o;
    {
        const pb2::LinearNode& n = b.nodes[0];
        Bounds3 bb{{n.bmin[0], n.bmin[1], n.bmin[2]}, {n.bmax[0], n.bmax[1], n.bmax[2]}};
        if (!slab_test(bb, ray, inv, neg)) return o;
    }
    struct E { uint32_t node; float t; };
    E stack[256]; int sp = 0;
    uint32_t cur = 0;
    for (;;) {
        const pb2::LinearNode& n = b.nodes[cur];
        if (n.n_prims > 0) {
            ++st->leaf_visits;
            for (uint32_t i = 0; i < n.n_prims; ++i) {
                const pb2::PackedTri& t = b.tris[n.offset + i];
                V3 p0{t.v0[0], t.v0[1], t.v0[2]}, p1{t.v1[0], t.v1[1], t.v1[2]}, p2{t.v2[0], t.v2[1], t.v2[2]};
                ++st->tris;
                TriHit h = triangle_intersect_test(p0, p1, p2, ray);
                if (!h.hit) continue;
                V3 du, dv;
                if (!triangle_frame(p0, p1, p2, &du, &dv)) continue;
                ray.t_max = h.t; o.prim = t.prim_id; o.t = h.t;
            }
        } else {
            ++st->steps;
            uint32_t f[8]; int nf = 0;
            frontier(b, cur, levels, neg, f, &nf);
            E e[8];
            int first = -1;
            for (int k = 0; k < nf; ++k) {
                const pb2::LinearNode& c = b.nodes[f[k]];
                Bounds3 bb{{c.bmin[0], c.bmin[1], c.bmin[2]}, {c.bmax[0], c.bmax[1], c.bmax[2]}};
                ++st->boxes;
                Float te;
                e[k].node = f[k];
                e[k].t = slab_test(bb, ray, inv, neg, &te) ? te : kInfinity;
                if (first < 0 && e[k].t < kInfinity) first = k;
            }
            if (first >= 0) {
                for (int k = nf - 1; k > first; --k) if (e[k].t < kInfinity) { stack[sp++] = e[k]; ++st->pushes; }
                cur = e[first].node;
                continue;
            }
        }
        bool got = false;
        while (sp > 0) { --sp; ++st->pops; if (stack[sp].t < ray.t_max) { cur = stack[sp].node; got = true; break; } }
        if (!got) break;
    }
    return o;
}
// stats[0..5] = levels 2: steps, boxes, tris, pushes, pops, leaf visits; [6..11] = levels 3; [12] mismatches vs the binary walk
extern "C" int wide_sim(const float* verts, uint64_t nv, const uint32_t* idx, uint64_t nt, int max_prims, const float* rays, uint64_t n, uint64_t* stats) {
    pb2::HostBVH b;
    pb2::build_sah_bvh(verts, nv, idx, nt, max_prims, 0, &b);
    std::atomic<uint64_t> acc[13];
    for (auto& a : acc) a = 0;
    std::atomic<uint64_t> next{0};
    auto work = [&]() {
        WideStats s2, s3; uint64_t mm = 0, dn = 0, dt = 0;
        for (;;) {
            const uint64_t i0 = next.fetch_add(4096);
            if (i0 >= n) break;
            for (uint64_t i = i0; i < std::min(n, i0 + 4096); ++i) {
                Ray r; std::memcpy(&r, rays + 8 * i, 32);
                V3 inv{1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z};
                if (!plain_ray(r, inv)) continue;
                const SimOut a = walk_binary(b, r, &dn, &dt);
                const SimOut w2 = walk_wide(b, r, 2, &s2), w3 = walk_wide(b, r, 3, &s3);
                if (a.prim != w2.prim || float_to_bits(a.t) != float_to_bits(w2.t) || a.prim != w3.prim || float_to_bits(a.t) != float_to_bits(w3.t)) ++mm;
            }
        }
        const uint64_t v[13] = {s2.steps, s2.boxes, s2.tris, s2.pushes, s2.pops, s2.leaf_visits, s3.steps, s3.boxes, s3.tris, s3.pushes, s3.pops, s3.leaf_visits, mm};
        for (int k = 0; k < 13; ++k) acc[k] += v[k];
    };
    std::vector<std::thread> pool;
    const unsigned T = std::max(1u, std::thread::hardware_concurrency());
    for (unsigned t = 1; t < T; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    for (int k = 0; k < 13; ++k) stats[k] = acc[k];
    return 0;
}

extern "C" int quad_sim(const float* verts, uint64_t nv, const uint32_t* idx, uint64_t nt, int max_prims, const float* rays, uint64_t n,
                        uint64_t* stats /* [8] */) {
    pb2::HostBVH b;
    pb2::build_sah_bvh(verts, nv, idx, nt, max_prims, 0, &b);
    std::atomic<uint64_t> a_nodes{0}, a_tris{0}, q_steps{0}, q_boxes{0}, q_tris{0}, mism{0}, nonplain{0};
    std::atomic<int> max_sp{0};
    std::atomic<uint64_t> next{0};
    auto work = [&]() {
        uint64_t ln = 0, lt = 0, qs = 0, qb = 0, qt = 0, mm = 0, np = 0; int msp = 0;
        for (;;) {
            const uint64_t i0 = next.fetch_add(4096);
            if (i0 >= n) break;
            for (uint64_t i = i0; i < std::min(n, i0 + 4096); ++i) {
                Ray r; std::memcpy(&r, rays + 8 * i, 32);
                V3 inv{1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z};
                if (!plain_ray(r, inv)) { ++np; continue; }
                const SimOut a = walk_binary(b, r, &ln, &lt);
                const SimOut q = walk_quad(b, r, &qs, &qb, &qt, &msp);
                if (a.prim != q.prim || float_to_bits(a.t) != float_to_bits(q.t)) ++mm;
            }
        }
        a_nodes += ln; a_tris += lt; q_steps += qs; q_boxes += qb; q_tris += qt; mism += mm; nonplain += np;
        int cur = max_sp.load(); while (msp > cur && !max_sp.compare_exchange_weak(cur, msp)) {}
    };
    std::vector<std::thread> pool;
    const unsigned T = std::max(1u, std::thread::hardware_concurrency());
    for (unsigned t = 1; t < T; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    stats[0] = a_nodes; stats[1] = a_tris; stats[2] = q_steps; stats[3] = q_boxes; stats[4] = q_tris; stats[5] = mism; stats[6] = nonplain;
    stats[7] = ((uint64_t)b.quads.size() << 8) | (uint64_t)max_sp.load();
    std::printf("pairs %zu quads %zu nodes %zu depth %d\n", (size_t)0, b.quads.size(), b.nodes.size(), b.max_depth);
    return 0;
}
