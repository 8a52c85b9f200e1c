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
//  * The traversal stack (64 entries, bvh.rs:839) lives in local memory with its top entry cached in registers: a pop
//    consumes the register copy and issues the reload of the next entry, which is not needed before the next pop,
//    so the local-memory latency stays off the critical path.  Far children are prefetched into L2 when pushed.
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

PB2_D void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

#ifndef PB2_TOPCACHE
#define PB2_TOPCACHE 1
#endif
#ifndef PB2_FASTSLAB
#define PB2_FASTSLAB 1
#endif

// Slab test for rays whose inverse direction is finite and non-zero on all three axes (every product below is then
// an ordinary number or +-inf, never 0 * inf = NaN).  Without NaNs the reference's sequence of rejections and
// conditional updates (geometry.rs:709-751) reduces to entry = max of the three near values, exit = min of the three
// widened far values, accept iff entry <= exit and exit > 0: the two early rejections are exactly the six cross-axis
// comparisons, and the three same-axis comparisons hold automatically whenever exit > 0 (far >= near before widening,
// and widening a positive value never decreases it).  entry equals the reference's final t_min.
PB2_D bool slab_entry_fast(const RayCtx& r, float lx, float ly, float lz, float hx, float hy, float hz, float* t_entry) {
    const float widen = 1.0f + 2.0f * gammaf_(3.0f);
    const float txn = ((r.nx ? hx : lx) - r.o.x) * r.inv.x;
    const float txf = (((r.nx ? lx : hx) - r.o.x) * r.inv.x) * widen;
    const float tyn = ((r.ny ? hy : ly) - r.o.y) * r.inv.y;
    const float tyf = (((r.ny ? ly : hy) - r.o.y) * r.inv.y) * widen;
    const float tzn = ((r.nz ? hz : lz) - r.o.z) * r.inv.z;
    const float tzf = (((r.nz ? lz : hz) - r.o.z) * r.inv.z) * widen;
    const float tn = fmaxf(fmaxf(txn, tyn), tzn);
    const float tf = fminf(fminf(txf, tyf), tzf);
    *t_entry = tn;
    return tn <= tf && tf > 0.0f;
}
PB2_D bool finite_nonzero(float v) { return v != 0.0f && fabsf(v) < __int_as_float(0x7f800000); }

// Sink = where rays come from and where results go (batch arrays, wavefront queues).
//   bool load(uint64_t i, vec3* o, vec3* d, float* t_max)   — false: slot i carries no ray
//   void miss_or_hit(uint64_t i, uint32_t prim, float t, float b0, float b1, float b2)   (closest hit)
//   void occluded(uint64_t i, bool occ)                                                   (any hit)
template <bool ANY, class Sink>
__device__ __forceinline__ void trace_persistent(const SceneView& s, uint64_t n, unsigned long long* __restrict__ counter,
                                                 const Sink& sink, const TraceTuning tune) {
    const unsigned lane = threadIdx.x & 31u;
    uint32_t stack_ref[kStackDepth];
    float stack_t[kStackDepth];
    int sp = 0;                              // entries on the stack, the newest one held in (top_ref, top_t)
    uint32_t top_ref = 0;
    float top_t = 0.0f;
    uint32_t cur = kDone;
    uint64_t ray_idx = 0;
    RayCtx r = make_ray_ctx(mk(0.f, 0.f, 0.f), mk(0.f, 0.f, 1.f));
    float t_max = 0.0f;
    uint32_t h_prim = 0xFFFFFFFFu;
    float h_b0 = 0.f, h_b1 = 0.f, h_b2 = 0.f;
    bool exhausted = false;                  // warp-uniform: the ray counter ran past n
    bool plain = true;                       // this lane's ray qualifies for slab_entry_fast

    for (;;) {
        // ---- refill idle lanes ----
        if (!exhausted) {
            const unsigned idle = __ballot_sync(kFullMask, cur == kDone);
            if (idle) {
                const int leader = __ffs(idle) - 1;
                unsigned long long base = 0;
                if ((int)lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(idle));
                base = __shfl_sync(kFullMask, base, leader);
                exhausted = base + (unsigned)__popc(idle) >= n;
                if (cur == kDone) {
                    ray_idx = base + (unsigned)__popc(idle & ((1u << lane) - 1u));
                    if (ray_idx < n) {
                        vec3 o, d;
                        if (sink.load(ray_idx, &o, &d, &t_max)) {
                            r = make_ray_ctx(o, d);
                            plain = finite_nonzero(r.inv.x) && finite_nonzero(r.inv.y) && finite_nonzero(r.inv.z);
                            h_prim = 0xFFFFFFFFu; h_b0 = 0.f; h_b1 = 0.f; h_b2 = 0.f;
                            sp = 0;
                            float te;
                            const bool enter = s.n_tris != 0 &&
                                slab_entry(r, s.root_lo[0], s.root_lo[1], s.root_lo[2], s.root_hi[0], s.root_hi[1], s.root_hi[2], &te) && te < t_max;
                            if (enter) cur = s.root_ref;
                            else if (ANY) sink.occluded(ray_idx, false);
                            else sink.closest(ray_idx, 0xFFFFFFFFu, t_max, 0.f, 0.f, 0.f);
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
            if (work_mask == 0u) { if (exhausted) return; break; }
            if (!exhausted && __popc(work_mask) < tune.refill_below) break;
            const int n_node = __popc(node_mask), n_leaf = __popc(work_mask & ~node_mask);
            bool need_pop = false;
#if PB2_FASTSLAB
            const bool all_plain = __ballot_sync(kFullMask, at_node && !plain) == 0u;
#endif
            if (n_leaf == 0 || (n_node >= tune.node_quorum && n_leaf < tune.leaf_quorum)) {
                if (at_node) {
                    const float4* np = s.pairs + 4ull * cur;
                    const float4 a = ldg4(np), b = ldg4(np + 1), c = ldg4(np + 2);
                    const uint4 m = __ldg(reinterpret_cast<const uint4*>(np + 3));
                    float tl, tr;
                    bool okl, okr;
#if PB2_FASTSLAB
                    if (all_plain) {
                        okl = slab_entry_fast(r, a.x, a.y, a.z, a.w, b.x, b.y, &tl) && (tl < t_max);
                        okr = slab_entry_fast(r, b.z, b.w, c.x, c.y, c.z, c.w, &tr) && (tr < t_max);
                    } else
#endif
                    {
                        okl = slab_entry(r, a.x, a.y, a.z, a.w, b.x, b.y, &tl) && (tl < t_max);
                        okr = slab_entry(r, b.z, b.w, c.x, c.y, c.z, c.w, &tr) && (tr < t_max);
                    }
                    // bvh.rs:856-866: near child first, by the sign of the direction on the split axis
                    const bool neg = (m.z == 0u) ? r.nx : ((m.z == 1u) ? r.ny : r.nz);
                    const uint32_t near_ref = neg ? m.y : m.x, far_ref = neg ? m.x : m.y;
                    const bool ok_near = neg ? okr : okl, ok_far = neg ? okl : okr;
                    if (ok_near) {
                        if (ok_far) {
#if PB2_TOPCACHE
                            if (sp > 0) { stack_ref[sp - 1] = top_ref; stack_t[sp - 1] = top_t; }
                            top_ref = far_ref;
                            top_t = neg ? tl : tr;
#else
                            stack_ref[sp] = far_ref;
                            stack_t[sp] = neg ? tl : tr;
#endif
                            ++sp;
                            if (tune.prefetch)
                                prefetch_l2((far_ref & kLeafFlag) ? (const void*)(s.tris + 3ull * (far_ref & ~kLeafFlag))
                                                                  : (const void*)(s.pairs + 4ull * far_ref));
                        }
                        cur = near_ref;
                    } else if (ok_far) {
                        cur = far_ref;
                    } else {
                        need_pop = true;
                    }
                }
            } else if (cur != kDone && !at_node) {
                // leaf triangles, in leaf order
                uint32_t slot = cur & ~kLeafFlag;
                bool occluded = false;
                for (;;) {
                    const float4 a = ldg4(s.tris + 3ull * slot);
                    const float4 b = ldg4(s.tris + 3ull * slot + 1);
                    const float4 c = ldg4(s.tris + 3ull * slot + 2);
                    const vec3 p0 = mk(a.x, a.y, a.z), p1 = mk(b.x, b.y, b.z), p2 = mk(c.x, c.y, c.z);
                    float t, b0, b1, b2;
                    if (tri_test(r, t_max, p0, p1, p2, &t, &b0, &b1, &b2)) {
                        if (ANY) { occluded = true; break; }
                        vec3 du, dv;
                        if (tri_frame(p0, p1, p2, &du, &dv)) {
                            t_max = t;                                  // primitive.rs:70
                            h_prim = __float_as_uint(a.w);
                            h_b0 = b0; h_b1 = b1; h_b2 = b2;
                        }
                    }
                    if (__float_as_uint(b.w) != 0u) break;              // last triangle of the leaf
                    ++slot;
                }
                if (ANY && occluded) {
                    cur = kDone;
                    sink.occluded(ray_idx, true);
                } else {
                    need_pop = true;
                }
            }
            if (need_pop) {
                cur = kDone;
                while (sp > 0) {                                        // geometry.rs:749 re-applied at pop time
#if PB2_TOPCACHE
                    const uint32_t e_ref = top_ref;
                    const float e_t = top_t;
                    --sp;
                    if (sp > 0) { top_ref = stack_ref[sp - 1]; top_t = stack_t[sp - 1]; }
#else
                    --sp;
                    const uint32_t e_ref = stack_ref[sp];
                    const float e_t = stack_t[sp];
#endif
                    if (e_t < t_max) { cur = e_ref; break; }
                }
                if (cur == kDone) {
                    if (ANY) sink.occluded(ray_idx, false);
                    else sink.closest(ray_idx, h_prim, t_max, h_b0, h_b1, h_b2);
                }
            }
        }
    }
}

}  // namespace pb2
