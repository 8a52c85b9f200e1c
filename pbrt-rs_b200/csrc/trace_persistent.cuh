// trace_persistent.cuh — the persistent-warp traversal loop shared by k_closest_hit, k_any_hit and the wavefront
// extend / shadow stages (sm_100a).  bvh.rs:828-879 / :881-932 semantics; see traverse.cuh for the exactness notes.
//
// One grid sized to the machine; every lane owns one ray pulled from a global counter.
//  * Refill: when fewer than `refill_below` lanes of a warp still have work, the idle lanes fetch new rays (warp
//    ballot + one atomicAdd by the leader lane), so lanes stay busy despite very uneven walk lengths.
//  * Phase scheduling: a lane is either at an interior node or holding a leaf.  Each warp iteration takes a vote and
//    runs ONE of the two code paths for the whole warp — the leaf step once `leaf_quorum` lanes hold a leaf or fewer
//    than `node_quorum` lanes still want a node step, the node step otherwise — so each path runs with most of its
//    lanes enabled instead of every iteration paying for both.  (The tree is deep and leaves hold ~1.3 triangles: a
//    per-lane "while-while" loop idles 3 of 4 lanes waiting for the slowest lane to reach its next leaf.)
//  * Register diet: throughput follows resident warps (profiles/r01_tuning.md: loading a node's record early into
//    registers lost 20-35 % because it cost a CTA per SM), so the per-ray state carried between iterations is 15
//    registers — origin, inverse direction, shear, t_max, current reference, stack depth, ray index and one word of
//    flags (direction signs, shear axis, slab fast-path, hit-found) — and everything else is rebuilt inside the
//    step that needs it.  An accepted closest-hit candidate is written to the result at once (the last one written
//    wins, as primitive.rs:70 keeps the last accepted hit), so no hit record is carried.
//  * The traversal stack (64 entries, bvh.rs:839) is one uint2 {reference, entry distance} array in local memory.
// Scheduling has no influence on results: every ray still sees the reference's node and triangle order.
#pragma once
#include "traverse.cuh"

namespace pb2 {

constexpr uint32_t kDone = 0xFFFFFFFFu;     // sentinel reference: leaf bit set, never a valid slot
constexpr unsigned kFullMask = 0xFFFFFFFFu;

struct TraceTuning {
    int refill_below;
    int node_quorum;
    int leaf_quorum;
    int prefetch;
};

#ifndef PB2_MIN_BLOCKS
#define PB2_MIN_BLOCKS 8   /* 64 registers: 8 CTAs of 4 warps per SM (the quad step spills below that; profiles/r01_tuning.md) */
#endif
// Node steps read QuadNode records: two tree levels per fetch, half the dependent fetches and votes per ray of the
// one-level PairNode walk this kernel used first (profiles/r01_tuning.md); leaves are still visited in the reference's order.
constexpr int kQuadStackDepth = 96;         // <= 3 pushes per two levels of a tree at most 64 levels deep (bvh.rs:839)

PB2_D void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

PB2_D bool finite_nonzero(float v) { return v != 0.0f && fabsf(v) < __int_as_float(0x7f800000); }

// Slab test for rays whose inverse direction is finite and non-zero on all three axes (every product below is then an
// ordinary number or +-inf, never 0 * inf = NaN).  Without NaNs the reference's sequence of rejections and conditional
// updates (geometry.rs:709-751) reduces to entry = max of the three near values, exit = min of the three widened far
// values, accept iff entry <= exit, exit > 0 and entry < t_max: the two early rejections are exactly the six cross-axis
// comparisons, and the three same-axis comparisons hold automatically whenever exit > 0 (far >= near before widening, and
// widening a positive value never decreases it); entry equals the reference's final t_min.  Arguments: the near and far
// planes of the box for this ray's direction signs.  Returns the entry distance, +inf when the box is missed or starts
// beyond t_max.  The widening factor is applied once, to the smallest far value: x -> round(x * widen) is non-decreasing
// (widen > 0, rounding is monotone), so min(round(a*w), round(b*w), round(c*w)) = round(min(a, b, c) * w) bit for bit — two
// multiplies per box fewer than widening each far plane as geometry.rs:722-738 does.  (Only the sign of a zero can differ, and
// exit enters two comparisons only.)
PB2_D float quad_child_entry(float nx, float ny, float nz, float fx, float fy, float fz, vec3 o, vec3 inv, float t_max) {
    const float widen = 1.0f + 2.0f * gammaf_(3.0f);
    const float tn = fmaxf(fmaxf((nx - o.x) * inv.x, (ny - o.y) * inv.y), (nz - o.z) * inv.z);
    const float tf = fminf(fminf((fx - o.x) * inv.x, (fy - o.y) * inv.y), (fz - o.z) * inv.z) * widen;
    // one select on the conjunction of the three ordered comparisons (the compiler's own code selects +inf once per comparison:
    // 6 instead of 4 instructions per box, 8 more per QuadNode step)
    float r;
    asm("{\n\t.reg .pred p;\n\tsetp.le.f32 p, %1, %2;\n\tsetp.gt.and.f32 p, %2, 0f00000000, p;\n\tsetp.lt.and.f32 p, %1, %3, p;\n\t"
        "selp.f32 %0, %1, 0f7F800000, p;\n\t}"
        : "=f"(r)
        : "f"(tn), "f"(tf), "f"(t_max));
    return r;
}
PB2_D void swap_if(bool c, uint32_t& ra, float& ta, uint32_t& rb, float& tb) {
    const uint32_t r = c ? rb : ra; rb = c ? ra : rb; ra = r;
    const float t = c ? tb : ta; tb = c ? ta : tb; ta = t;
}

// Flag word of a ray: bits 0-2 direction signs (bvh.rs:832-836), bits 3-4 shear axis kz (triangle.rs:84-92),
// bit 5 slab fast path allowed, bit 6 a closest-hit candidate has been accepted.
// bit 7: the walk is over and its result has not been handed to the sink yet (closest hit: finish(); any hit: occluded(),
// with bit 6 = "occluded").  The hand-off waits for the warp's next refill, where the ~20 lanes that ended since the last one
// make their sink calls together; made at the point where a single walk ends it ran with 2.9 of 32 lanes enabled and took
// 5 % of the warp instructions and 12 % of the stall samples of k_extend (profiles/r02_tuning.md).
constexpr uint32_t kFlagPlain = 32u, kFlagFound = 64u, kFlagFinish = 128u;

// The part of RayCtx a step needs, rebuilt from the carried registers (make_ray_ctx computed them once per ray).
PB2_D RayCtx ctx_of(vec3 o, vec3 inv, vec3 sh, uint32_t flags) {
    RayCtx r;
    r.o = o;
    r.d = mk(0.f, 0.f, 0.f);                 // not used by slab_entry / tri_test
    r.inv = inv;
    r.nx = (flags & 1u) != 0u;
    r.ny = (flags & 2u) != 0u;
    r.nz = (flags & 4u) != 0u;
    r.kz = (int)((flags >> 3) & 3u);
    r.kx = r.kz == 2 ? 0 : r.kz + 1;
    r.ky = r.kx == 2 ? 0 : r.kx + 1;
    r.sx = sh.x; r.sy = sh.y; r.sz = sh.z;
    return r;
}

// Sink = where rays come from and where results go (batch arrays, wavefront queues).  i < n < 2^32.
//   bool load(uint32_t i, vec3* o, vec3* d, float* t_max)                                — false: slot i carries no ray
//   void accept(uint32_t i, uint32_t prim, float t, float b0, float b1, float b2)       — closest hit: a candidate was
//                                                     accepted (called in test order; the last call is the closest hit)
//   void accept_sphere(uint32_t i, uint32_t prim, float t, float u, float v)            — the same for an analytic sphere (SPH only)
//   void finish(uint32_t i, bool found, float t_max)                                     — closest hit: walk over
//   void occluded(uint32_t i, bool occ)                                                  — any hit: walk over
// SPH: the scene holds analytic spheres (sphere.cuh); a separate instantiation, so that triangle-only scenes keep their
// register budget.  A sphere test needs the ray's direction itself (the walk carries its inverse only), so it reloads the ray
// from the sink; a sphere hit is handed to accept() with b0 = t (Sphere::intersect has no barycentrics; b1, b2 = u, v).
template <bool ANY, bool SPH, class Sink>
__device__ __forceinline__ void trace_persistent(const SceneView& s, uint32_t n, unsigned long long* __restrict__ counter,
                                                 const Sink& sink, const TraceTuning tune) {
    uint2 stack[kQuadStackDepth];            // {reference, entry distance bits}
    int sp = 0;
    uint2 top = make_uint2(0u, 0u);          // register copy of stack[sp - 1]: a pop never waits for local memory
    uint32_t cur = kDone;
    uint32_t ray_idx = 0;
    uint32_t flags = 0;
    vec3 o = mk(0.f, 0.f, 0.f), inv = mk(1.f, 1.f, 1.f), sh = mk(0.f, 0.f, 1.f);
    float t_max = 0.0f;
    bool exhausted = false;                  // warp-uniform: the ray counter ran past n
#ifndef PB2_DEFER_FINISH
#define PB2_DEFER_FINISH 1      /* 0: hand the result over where the walk ends (the round-1 kernel; tuning builds only) */
#endif
    auto hand_off = [&]() {
        if (flags & kFlagFinish) {
            flags &= ~kFlagFinish;
            if (ANY) sink.occluded(ray_idx, (flags & kFlagFound) != 0u);
            else sink.finish(ray_idx, (flags & kFlagFound) != 0u, t_max);
        }
    };

    for (;;) {
        // ---- refill idle lanes ----
        if (!exhausted) {
            const unsigned idle = __ballot_sync(kFullMask, cur == kDone);
            if (idle) {
                const unsigned lane = threadIdx.x & 31u;
                const int leader = __ffs(idle) - 1;
                unsigned long long base = 0;
                if ((int)lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(idle));
                base = __shfl_sync(kFullMask, base, leader);
                exhausted = base + (unsigned)__popc(idle) >= n;
                if (cur == kDone) {
                    hand_off();
                    const unsigned long long mine = base + (unsigned)__popc(idle & ((1u << lane) - 1u));
                    if (mine < n) {
                        ray_idx = (uint32_t)mine;
                        vec3 d;
                        if (sink.load(ray_idx, &o, &d, &t_max)) {
                            const RayCtx r = make_ray_ctx(o, d);
                            inv = r.inv;
                            sh = mk(r.sx, r.sy, r.sz);
                            flags = (r.nx ? 1u : 0u) | (r.ny ? 2u : 0u) | (r.nz ? 4u : 0u) | ((uint32_t)r.kz << 3);
                            const float inf = __int_as_float(0x7f800000);
                            if (finite_nonzero(inv.x) && finite_nonzero(inv.y) && finite_nonzero(inv.z) && fabsf(o.x) < inf &&
                                fabsf(o.y) < inf && fabsf(o.z) < inf)
                                flags |= kFlagPlain;
                            sp = 0;
                            if (!(flags & kFlagPlain)) {
                                // A zero direction component makes 0 * inf = NaN possible in the slab test, and a NaN can
                                // reject a box whose child it accepts; folding two levels relies on "child hit => parent
                                // hit", so these (rare) rays take the literal one-level walk right here.
                                HitRec h;
                                const bool found = traverse<ANY, SPH>(s, o, d, t_max, &h);
                                if (ANY) sink.occluded(ray_idx, found);
                                else {
                                    if (found) {
                                        if (SPH && h.sphere) sink.accept_sphere(ray_idx, h.prim, h.t, h.b1, h.b2);
                                        else sink.accept(ray_idx, h.prim, h.t, h.b0, h.b1, h.b2);
                                    }
                                    sink.finish(ray_idx, found, found ? h.t : t_max);
                                }
                            } else
                            {
                                float te;
                                const bool enter = s.n_tris != 0 &&
                                    slab_entry(r, s.root_lo[0], s.root_lo[1], s.root_lo[2], s.root_hi[0], s.root_hi[1], s.root_hi[2], &te) && te < t_max;
                                if (enter) cur = s.quad_root_ref;
                                else if (ANY) sink.occluded(ray_idx, false);
                                else sink.finish(ray_idx, false, t_max);
                            }
                        }
                    }
                }
            }
        }
        // ---- walk until the warp wants a refill (or is finished) ----
        for (;;) {
            const bool at_node = !(cur & kLeafFlag);
            const unsigned node_mask = __ballot_sync(kFullMask, at_node);
            const unsigned work_mask = __ballot_sync(kFullMask, cur != kDone);
            if (work_mask == 0u) {
                if (exhausted) { hand_off(); return; }
                break;
            }
            if (!exhausted && __popc(work_mask) < tune.refill_below) break;
            const int n_node = __popc(node_mask), n_leaf = __popc(work_mask & ~node_mask);
            bool need_pop = false;
            // keep the flag word opaque so the per-step decoding below is not hoisted into loop-carried registers
            asm volatile("" : "+r"(flags));
            if (n_leaf == 0 || (n_node >= tune.node_quorum && n_leaf < tune.leaf_quorum)) {
                if (at_node) {
                    // One QuadNode = interior node P, its children A, B and their children.  bvh.rs:856-866 visits A's
                    // subtree before B's unless the ray is negative on axis(P), and inside A (B) the first child before
                    // the second unless negative on axis(A) (axis(B)); a box is tested when its node is visited.  Here
                    // the four grandchild boxes are tested at once, ordered the same way, the nearest accepted one is
                    // visited and the others wait on the stack with their entry distance, which is re-checked against
                    // the then-current t_max when popped.  A's and B's own boxes need no test: on the fast path a box
                    // that accepts the ray implies its parent does (rounded subtract / multiply are monotone), and the
                    // parent's t_max at its visit is never smaller than the child's.
                    const float4* qp = s.quads + 8ull * cur;
                    float4 lox, loy, loz, hix, hiy, hiz;
                    ldg8(qp, &lox, &loy);
                    ldg8(qp + 2, &loz, &hix);
                    ldg8(qp + 4, &hiy, &hiz);
                    const uint4 ref = __ldg(reinterpret_cast<const uint4*>(qp + 6));
                    const bool nx = (flags & 1u) != 0u, ny = (flags & 2u) != 0u, nz = (flags & 4u) != 0u;
                    float t0 = quad_child_entry(nx ? hix.x : lox.x, ny ? hiy.x : loy.x, nz ? hiz.x : loz.x,
                                                nx ? lox.x : hix.x, ny ? loy.x : hiy.x, nz ? loz.x : hiz.x, o, inv, t_max);
                    float t1 = quad_child_entry(nx ? hix.y : lox.y, ny ? hiy.y : loy.y, nz ? hiz.y : loz.y,
                                                nx ? lox.y : hix.y, ny ? loy.y : hiy.y, nz ? loz.y : hiz.y, o, inv, t_max);
                    float t2 = quad_child_entry(nx ? hix.z : lox.z, ny ? hiy.z : loy.z, nz ? hiz.z : loz.z,
                                                nx ? lox.z : hix.z, ny ? loy.z : hiy.z, nz ? loz.z : hiz.z, o, inv, t_max);
                    float t3 = quad_child_entry(nx ? hix.w : lox.w, ny ? hiy.w : loy.w, nz ? hiz.w : loz.w,
                                                nx ? lox.w : hix.w, ny ? loy.w : hiy.w, nz ? loz.w : hiz.w, o, inv, t_max);
                    // bits 29-30 of the first three references: axis(P), axis(A), axis(B)
                    const bool swap_groups = ((flags >> ((ref.x >> 29) & 3u)) & 1u) != 0u;   // A before B, or B before A
                    const bool swap_a = ((flags >> ((ref.y >> 29) & 3u)) & 1u) != 0u, swap_b = ((flags >> ((ref.z >> 29) & 3u)) & 1u) != 0u;
                    uint32_t r0 = ref.x & 0x9FFFFFFFu, r1 = ref.y & 0x9FFFFFFFu, r2 = ref.z & 0x9FFFFFFFu, r3 = ref.w;
                    if (!ANY) {
                        swap_if(swap_a, r0, t0, r1, t1);                                      // inside A
                        swap_if(swap_b, r2, t2, r3, t3);                                      // inside B
                        swap_if(swap_groups, r0, t0, r2, t2);
                        swap_if(swap_groups, r1, t1, r3, t3);
                    }
                    // (intersect_p's answer does not depend on the visiting order: t_max never shrinks, so the set of leaves whose
                    // boxes the ray passes — and with it "some triangle in them is hit" — is the same in any order)
                    const float inf = __int_as_float(0x7f800000);
                    const bool h0 = t0 < inf, h1 = t1 < inf, h2 = t2 < inf, h3 = t3 < inf;
                    // later candidates first, so the next one in visiting order is popped first
                    if (h3 && (h0 || h1 || h2)) { top = make_uint2(r3, __float_as_uint(t3)); stack[sp] = top; ++sp; }
                    if (h2 && (h0 || h1)) { top = make_uint2(r2, __float_as_uint(t2)); stack[sp] = top; ++sp; }
                    if (h1 && h0) { top = make_uint2(r1, __float_as_uint(t1)); stack[sp] = top; ++sp; }
                    if (h0 || h1 || h2 || h3) cur = h0 ? r0 : (h1 ? r1 : (h2 ? r2 : r3));
                    else need_pop = true;
                }
            } else if (cur != kDone && !at_node) {
                // leaf triangles, in leaf order
                const RayCtx r = ctx_of(o, inv, sh, flags);
                uint32_t slot = cur & ~kLeafFlag;
                bool occluded = false;
                for (;;) {
                    const float4 a = ldg4(s.tris + 3ull * slot);
                    const float4 b = ldg4(s.tris + 3ull * slot + 1);
                    const float4 c = ldg4(s.tris + 3ull * slot + 2);
                    const vec3 p0 = mk(a.x, a.y, a.z), p1 = mk(b.x, b.y, b.z), p2 = mk(c.x, c.y, c.z);
                    float t, b0, b1, b2;
                    if (SPH && (__float_as_uint(c.w) & 2u)) {
                        vec3 ro, rd, p_hit, od;
                        float tm, phi;
                        sink.load(ray_idx, &ro, &rd, &tm);
                        if (sphere_test(sphere_of(s, a), ro, rd, t_max, &p_hit, &phi, &od, &t)) {
                            if (ANY) { occluded = true; break; }
                            const SphereVertex sv = sphere_vertex(sphere_of(s, a), p_hit, phi, od);
                            t_max = t;
                            flags |= kFlagFound;
                            sink.accept_sphere(ray_idx, __float_as_uint(a.w), t, sv.u, sv.v);
                        }
                    } else
                    if (tri_test(r, t_max, p0, p1, p2, &t, &b0, &b1, &b2)) {
                        if (ANY) { occluded = true; break; }
                        // c.w: Triangle::intersect bails out on this triangle's degenerate frame (triangle.rs:193-215);
                        // a property of the triangle alone, evaluated once by k_mark_degenerate with tri_frame()
                        if (__float_as_uint(c.w) == 0u) {
                            t_max = t;                                  // primitive.rs:70
                            flags |= kFlagFound;
                            sink.accept(ray_idx, __float_as_uint(a.w), t, b0, b1, b2);
                        }
                    }
                    if (__float_as_uint(b.w) != 0u) break;              // last triangle of the leaf
                    ++slot;
                }
                if (ANY && occluded) {
                    cur = kDone;
                    flags |= kFlagFinish | kFlagFound;
                    if (!PB2_DEFER_FINISH) hand_off();
                } else {
                    need_pop = true;
                }
            }
            if (need_pop) {
                cur = kDone;
                while (sp > 0) {                                        // geometry.rs:749 re-applied at pop time
                    --sp;
                    const uint2 e = top;
                    if (sp > 0) top = stack[sp - 1];                    // needed at the next pop, not now
                    if (ANY || __uint_as_float(e.y) < t_max) { cur = e.x; break; }   // (any hit: t_max is the ray's own, checked at push time)
                }
                if (cur == kDone) {
                    flags |= kFlagFinish;
                    if (!PB2_DEFER_FINISH) hand_off();
                }
            }
        }
    }
}

}  // namespace pb2
