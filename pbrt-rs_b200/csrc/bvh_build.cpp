// bvh_build.cpp — host SAH build; see bvh_build.hpp.  Compiled with -ffp-contract=off.
//
// Arithmetic that decides the tree (centroids, bucket index, SAH cost, split choice, partition order) follows
// src/accelerators/bvh.rs:273-473 with the Appendix-A fixes D12-D17 (pbrt-v3 semantics), so the flattened array
// equals what the reference algorithm produces for the same primitive list.
#include "bvh_build.hpp"

#include <limits>

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cfloat>
#include <cstring>
#include <thread>

namespace pb2 {
namespace {

struct Box {
    float lo[3], hi[3];
    void reset() {
        lo[0] = lo[1] = lo[2] = FLT_MAX;            // Bounds3::new(): geometry.rs:439-448
        hi[0] = hi[1] = hi[2] = -FLT_MAX;
    }
    void grow(const float* blo, const float* bhi) {
        for (int k = 0; k < 3; ++k) {
            if (blo[k] < lo[k]) lo[k] = blo[k];
            if (bhi[k] > hi[k]) hi[k] = bhi[k];
        }
    }
    void grow_point(const float* p) { grow(p, p); }
    void merge(const Box& o) { grow(o.lo, o.hi); }
    float area() const {                            // geometry.rs:667-670
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return 2.0f * ((dx * dy + dx * dz) + dy * dz);
    }
    int widest() const {                            // geometry.rs:482-485 + :91-93
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return (dx > dy && dx > dz) ? 0 : ((dy > dz) ? 1 : 2);
    }
};

struct PrimRef {            // BVHPrimitiveInfo, bvh.rs:26-41
    float lo[3], hi[3];
    float c[3];
    uint32_t id;
    uint32_t bucket;        // scratch: SAH bucket of the current split
};

constexpr int kBuckets = 12;
constexpr int kMaxThreads = 64;

struct Split {
    Box bounds;
    int dim = 0;
    size_t mid = 0;
    bool leaf = false;
};

class Builder {
public:
    Builder(std::vector<PrimRef>& refs, int max_prims, int threads, int split_method)
        : refs_(refs), max_prims_(max_prims), threads_(threads), split_method_(split_method) {}

    // Decide what happens to [start, end): leaf, or partition about `mid` along `dim`.
    Split split_range(size_t start, size_t end, bool parallel) {
        Split s;
        const size_t n = end - start;
        Box cb;
        range_bounds(start, end, parallel, &s.bounds, &cb);
        if (n == 1) { s.leaf = true; return s; }
        s.dim = cb.widest();
        const int dim = s.dim;
        if (cb.hi[dim] == cb.lo[dim]) { s.leaf = true; return s; }
        if (split_method_ == 2) {
            // SplitMethod::Middle (bvh.rs:331-349): partition about the centroid-bound midpoint; an improper partition
            // falls through to EqualCounts as in pbrt-v3 (the port recurses on the improper split)
            const float p_mid = (cb.lo[dim] + cb.hi[dim]) / 2.0f;
            size_t lo = start, hi = end;
            for (;;) {
                while (lo < hi && refs_[lo].c[dim] < p_mid) ++lo;
                if (lo == hi) break;
                do { --hi; } while (lo < hi && !(refs_[hi].c[dim] < p_mid));
                if (lo == hi) break;
                std::swap(refs_[lo], refs_[hi]);
                ++lo;
            }
            if (lo != start && lo != end) { s.mid = lo; return s; }
        }
        if (split_method_ == 2 || split_method_ == 3) {
            // SplitMethod::EqualCounts (bvh.rs:350-360): nth_element about the middle by centroid[dim]
            s.mid = (start + end) / 2;
            std::nth_element(refs_.begin() + start, refs_.begin() + s.mid, refs_.begin() + end,
                             [dim](const PrimRef& a, const PrimRef& b) { return a.c[dim] < b.c[dim]; });
            return s;
        }
        if (n <= 2) {
            // bvh.rs:361-371: nth_element about the middle by centroid[dim]
            if (refs_[start + 1].c[dim] < refs_[start].c[dim]) std::swap(refs_[start], refs_[start + 1]);
            s.mid = (start + end) / 2;
            return s;
        }
        int count[kBuckets];
        Box bb[kBuckets];
        bucket_pass(start, end, parallel, cb, dim, count, bb);
        float cost[kBuckets - 1];
        const float total_area = s.bounds.area();
        for (int i = 0; i < kBuckets - 1; ++i) {
            Box b0, b1;
            b0.reset();
            b1.reset();
            int c0 = 0, c1 = 0;
            for (int j = 0; j <= i; ++j) { b0.merge(bb[j]); c0 += count[j]; }
            for (int j = i + 1; j < kBuckets; ++j) { b1.merge(bb[j]); c1 += count[j]; }
            cost[i] = 1.0f + ((float)c0 * b0.area() + (float)c1 * b1.area()) / total_area;    // bvh.rs:403
        }
        float min_cost = FLT_MAX;
        int min_bucket = 0;
        for (int i = 0; i < kBuckets - 1; ++i)
            if (cost[i] < min_cost) { min_cost = cost[i]; min_bucket = i; }
        const float leaf_cost = (float)n;
        if (!((int)n > max_prims_ || min_cost < leaf_cost)) { s.leaf = true; return s; }
        // Unstable two-pointer partition (Rust partition_in_place == libstdc++ bidirectional std::partition):
        // first element failing the predicate from the front is swapped with the last one passing it from the back.
        size_t lo = start, hi = end;
        const uint32_t mb = (uint32_t)min_bucket;
        for (;;) {
            while (lo < hi && refs_[lo].bucket <= mb) ++lo;
            if (lo == hi) break;
            do { --hi; } while (lo < hi && refs_[hi].bucket > mb);
            if (lo == hi) break;
            std::swap(refs_[lo], refs_[hi]);
            ++lo;
        }
        s.mid = lo;
        return s;
    }

    // Sequential subtree build, appending to `out` in depth-first order with indices local to `out`.
    void build_subtree(size_t start, size_t end, std::vector<LinearNode>& out, int depth, int* max_depth) {
        Split s = split_range(start, end, false);
        const uint32_t me = (uint32_t)out.size();
        out.emplace_back();
        write_bounds(out[me], s.bounds);
        if (depth > *max_depth) *max_depth = depth;
        if (s.leaf) {
            out[me].offset = (uint32_t)start;      // leaves are emitted in order, so first_prim_offset == start
            out[me].n_prims = (uint16_t)(end - start);
            if (end - start > 65535) leaf_overflow_.store(true);      // n_primitives is 16 bits in the 32-byte node
            out[me].axis = 0;
            out[me].pad = 0;
            return;
        }
        out[me].n_prims = 0;
        out[me].axis = (uint8_t)s.dim;
        out[me].pad = 0;
        build_subtree(start, s.mid, out, depth + 1, max_depth);
        out[me].offset = (uint32_t)out.size();
        build_subtree(s.mid, end, out, depth + 1, max_depth);
    }

    struct TopNode {
        LinearNode node;
        int child[2] = {-1, -1};
        int task = -1;
    };
    struct Task {
        size_t start, end;
        int depth;
        std::vector<LinearNode> nodes;
        int max_depth = 0;
    };

    int build_top(size_t start, size_t end, int depth, size_t grain) {
        if (end - start <= grain) return make_task(start, end, depth);
        Split s = split_range(start, end, true);
        if (s.leaf) {
            // oversized leaf (coincident centroids): the task re-derives it
            return make_task(start, end, depth);
        }
        int me = (int)top_.size();
        top_.emplace_back();
        write_bounds(top_[me].node, s.bounds);
        top_[me].node.n_prims = 0;
        top_[me].node.axis = (uint8_t)s.dim;
        top_[me].node.pad = 0;
        int c0 = build_top(start, s.mid, depth + 1, grain);
        int c1 = build_top(s.mid, end, depth + 1, grain);
        top_[me].child[0] = c0;
        top_[me].child[1] = c1;
        return me;
    }

    void run(HostBVH* out) {
        const size_t n = refs_.size();
        const size_t grain = std::max<size_t>(4096, n / (size_t)(threads_ * 8));
        int root = build_top(0, n, 1, threads_ > 1 ? grain : n);
        // phase 2: subtrees in parallel
        std::atomic<size_t> next{0};
        auto work = [&]() {
            for (;;) {
                size_t t = next.fetch_add(1);
                if (t >= tasks_.size()) break;
                Task& tk = tasks_[t];
                tk.nodes.reserve(2 * (tk.end - tk.start));
                tk.max_depth = tk.depth;
                build_subtree(tk.start, tk.end, tk.nodes, tk.depth, &tk.max_depth);
            }
        };
        std::vector<std::thread> pool;
        for (int i = 1; i < threads_; ++i) pool.emplace_back(work);
        work();
        for (auto& t : pool) t.join();
        // phase 3: concatenate in depth-first order
        size_t total = top_.size() - tasks_.size();
        for (auto& t : tasks_) total += t.nodes.size();
        out->nodes.clear();
        out->nodes.reserve(total);
        out->max_depth = 0;
        for (auto& t : tasks_) out->max_depth = std::max(out->max_depth, t.max_depth);
        emit(root, out->nodes);
        out->ordered_prims.resize(n);
        for (size_t i = 0; i < n; ++i) out->ordered_prims[i] = refs_[i].id;
        out->leaf_overflow = leaf_overflow_.load();
    }

private:
    std::atomic<bool> leaf_overflow_{false};
    std::vector<PrimRef>& refs_;
    int max_prims_;
    int threads_;
    int split_method_;          // bvh.rs:199-204: 0 SAH, 2 Middle, 3 EqualCounts (HLBVH is built on the device)
    std::vector<TopNode> top_;
    std::vector<Task> tasks_;

    static void write_bounds(LinearNode& n, const Box& b) {
        for (int k = 0; k < 3; ++k) { n.bmin[k] = b.lo[k]; n.bmax[k] = b.hi[k]; }
    }

    int make_task(size_t start, size_t end, int depth) {
        int me = (int)top_.size();
        top_.emplace_back();
        top_[me].task = (int)tasks_.size();
        tasks_.push_back(Task{start, end, depth, {}, 0});
        return me;
    }

    void emit(int t, std::vector<LinearNode>& out) {
        const TopNode& tn = top_[t];
        if (tn.task >= 0) {
            const Task& tk = tasks_[tn.task];
            const uint32_t base = (uint32_t)out.size();
            for (const LinearNode& ln : tk.nodes) {
                LinearNode c = ln;
                if (c.n_prims == 0) c.offset += base;
                out.push_back(c);
            }
            return;
        }
        const uint32_t me = (uint32_t)out.size();
        out.push_back(tn.node);
        emit(tn.child[0], out);
        out[me].offset = (uint32_t)out.size();
        emit(tn.child[1], out);
    }

    int slots_for(size_t n, bool parallel) const { return (parallel && n > (1u << 18)) ? threads_ : 1; }

    template <class F>
    void for_chunks(size_t start, size_t end, int nt, F&& fn) {
        const size_t n = end - start;
        if (nt <= 1) { fn(0, start, end); return; }
        std::vector<std::thread> pool;
        const size_t per = (n + nt - 1) / nt;
        for (int i = 0; i < nt; ++i) {
            size_t s = start + per * i, e = std::min(end, s + per);
            if (s >= e) break;
            pool.emplace_back([&fn, i, s, e]() { fn(i, s, e); });
        }
        for (auto& t : pool) t.join();
    }

    // union of primitive bounds (bvh.rs:283-286) and of centroids (bvh.rs:305-308, D12)
    void range_bounds(size_t start, size_t end, bool parallel, Box* bounds, Box* cbounds) {
        const int nt = slots_for(end - start, parallel);
        Box pb[kMaxThreads], pc[kMaxThreads];
        for (int i = 0; i < nt; ++i) { pb[i].reset(); pc[i].reset(); }
        for_chunks(start, end, nt, [&](int slot, size_t s, size_t e) {
            Box b, c;
            b.reset();
            c.reset();
            for (size_t i = s; i < e; ++i) {
                b.grow(refs_[i].lo, refs_[i].hi);
                c.grow_point(refs_[i].c);
            }
            pb[slot] = b;
            pc[slot] = c;
        });
        *bounds = pb[0];
        *cbounds = pc[0];
        for (int i = 1; i < nt; ++i) { bounds->merge(pb[i]); cbounds->merge(pc[i]); }
    }

    // bvh.rs:373-386 with D14: b = (int)(nBuckets * offset[dim]); offset = (c - min) / (max - min)  (geometry.rs:460-467)
    void bucket_pass(size_t start, size_t end, bool parallel, const Box& cb, int dim, int* count, Box* bb) {
        const int nt = slots_for(end - start, parallel);
        const float lo = cb.lo[dim], extent = cb.hi[dim] - cb.lo[dim];
        auto pass = [&](size_t s, size_t e, int* cnt, Box* bx) {
            for (int j = 0; j < kBuckets; ++j) { cnt[j] = 0; bx[j].reset(); }
            for (size_t i = s; i < e; ++i) {
                float o = refs_[i].c[dim] - lo;
                o /= extent;                              // cb.hi > cb.lo on this axis (checked by caller)
                int b = (int)((float)kBuckets * o);
                if (b == kBuckets) b = kBuckets - 1;
                refs_[i].bucket = (uint32_t)b;
                cnt[b]++;
                bx[b].grow(refs_[i].lo, refs_[i].hi);
            }
        };
        if (nt <= 1) { pass(start, end, count, bb); return; }
        std::vector<int> pcnt((size_t)nt * kBuckets);
        std::vector<Box> pbb((size_t)nt * kBuckets);
        for (auto& b : pbb) b.reset();
        for_chunks(start, end, nt, [&](int slot, size_t s, size_t e) { pass(s, e, &pcnt[(size_t)slot * kBuckets], &pbb[(size_t)slot * kBuckets]); });
        for (int j = 0; j < kBuckets; ++j) {
            count[j] = 0;
            bb[j].reset();
            for (int t = 0; t < nt; ++t) { count[j] += pcnt[(size_t)t * kBuckets + j]; bb[j].merge(pbb[(size_t)t * kBuckets + j]); }
        }
    }
};

}  // namespace

void build_sah_bvh(const float* verts, uint64_t n_verts, const uint32_t* indices, uint64_t n_tris,
                   int max_prims_in_node, int threads, HostBVH* out, int split_method, const float* extra_bounds, uint64_t n_extra) {
    (void)n_verts;
    *out = HostBVH();
    const uint64_t n_mesh = n_tris;
    n_tris += n_extra;                                           // the primitive list: the mesh's triangles, then the extra shapes
    if (n_tris == 0) return;
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    if (threads > kMaxThreads) threads = kMaxThreads;
    const int max_prims = std::min(max_prims_in_node, 255);      // bvh.rs:222
    std::vector<PrimRef> refs(n_tris);
    {
        // Triangle::world_bound (triangle.rs:175-180) and centroid = min*0.5 + max*0.5 (bvh.rs:38)
        std::atomic<uint64_t> next{0};
        auto work = [&]() {
            const uint64_t chunk = 1 << 16;
            for (;;) {
                uint64_t s = next.fetch_add(chunk);
                if (s >= n_tris) break;
                uint64_t e = std::min<uint64_t>(n_tris, s + chunk);
                for (uint64_t t = s; t < e; ++t) {
                    if (t >= n_mesh) {                           // Shape::world_bound of an analytic shape, handed in by the caller
                        const float* b = extra_bounds + 6ull * (t - n_mesh);
                        PrimRef& r = refs[t];
                        for (int k = 0; k < 3; ++k) { r.lo[k] = b[k]; r.hi[k] = b[3 + k]; r.c[k] = b[k] * 0.5f + b[3 + k] * 0.5f; }
                        r.id = (uint32_t)t;
                        r.bucket = 0;
                        continue;
                    }
                    const float* p0 = verts + 3ull * indices[3 * t];
                    const float* p1 = verts + 3ull * indices[3 * t + 1];
                    const float* p2 = verts + 3ull * indices[3 * t + 2];
                    PrimRef& r = refs[t];
                    for (int k = 0; k < 3; ++k) {
                        float lo = p0[k] < p1[k] ? p0[k] : p1[k];
                        float hi = p0[k] > p1[k] ? p0[k] : p1[k];
                        if (p2[k] < lo) lo = p2[k];
                        if (p2[k] > hi) hi = p2[k];
                        r.lo[k] = lo;
                        r.hi[k] = hi;
                        r.c[k] = lo * 0.5f + hi * 0.5f;
                    }
                    r.id = (uint32_t)t;
                    r.bucket = 0;
                }
            }
        };
        std::vector<std::thread> pool;
        for (int i = 1; i < threads; ++i) pool.emplace_back(work);
        work();
        for (auto& t : pool) t.join();
    }
    const bool timing = getenv("PB2_BUILD_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    Builder b(refs, max_prims, threads, split_method);
    b.run(out);
    const double t1 = now();
    repack_device_layout(verts, indices, n_mesh, out, n_extra);
    if (timing) fprintf(stderr, "[pb2] host build: tree %.3f s, device-layout repack %.3f s (%d threads)\n", t1 - t0, now() - t1, threads);
}

// Runs fn(begin, end) over [0, n) on the host's hardware threads.
template <class F>
static void parallel_ranges(size_t n, F&& fn) {
    const size_t T = std::max<size_t>(1, std::min<size_t>(std::thread::hardware_concurrency(), n / 65536 + 1));
    if (T == 1) { fn((size_t)0, n); return; }
    std::vector<std::thread> pool;
    const size_t step = (n + T - 1) / T;
    for (size_t t = 0; t < T; ++t) {
        const size_t lo = t * step, hi = std::min(n, lo + step);
        if (lo < hi) pool.emplace_back([&fn, lo, hi] { fn(lo, hi); });
    }
    for (auto& th : pool) th.join();
}

// LinearNode array + ordered_prims -> PairNode / QuadNode / PackedTri (shared by the SAH and HLBVH builders).  The node
// array is in depth-first order, so the numbering of the device records follows from two prefix counts over it (interior
// nodes -> pair index; interior nodes at even depth -> quad index, depth-first like the pairs) and every record can then
// be filled independently of the others.
void repack_device_layout(const float* verts, const uint32_t* indices, uint64_t n_tris, HostBVH* out, uint64_t n_extra) {
    const uint64_t n_mesh = n_tris;
    n_tris += n_extra;
    const size_t n_nodes = out->nodes.size();
    const std::vector<LinearNode>& nodes = out->nodes;
    std::vector<uint32_t> pair_of(n_nodes, 0), quad_of(n_nodes, 0);
    std::vector<uint8_t> odd(n_nodes, 0);                               // depth parity (parents precede their children)
    uint32_t n_pairs = 0, n_quads = 0;
    for (size_t i = 0; i < n_nodes; ++i) {
        if (nodes[i].n_prims != 0) continue;
        pair_of[i] = n_pairs++;
        if (!odd[i]) quad_of[i] = n_quads++;
        odd[i + 1] = odd[nodes[i].offset] = (uint8_t)(odd[i] ^ 1);
    }
    out->pairs.resize(n_pairs);
    out->quads.resize(n_quads);
    auto ref_of = [&](uint32_t node) -> uint32_t {
        const LinearNode& ln = nodes[node];
        return ln.n_prims > 0 ? (kLeafBit | ln.offset) : pair_of[node];
    };
    parallel_ranges(n_nodes, [&](size_t lo, size_t hi) {
        const float inf = std::numeric_limits<float>::infinity();
        for (size_t i = lo; i < hi; ++i) {
            const LinearNode& ln = nodes[i];
            if (ln.n_prims != 0) continue;
            const LinearNode& L = nodes[i + 1];
            const LinearNode& R = nodes[ln.offset];
            PairNode& p = out->pairs[pair_of[i]];
            p.a[0] = L.bmin[0]; p.a[1] = L.bmin[1]; p.a[2] = L.bmin[2]; p.a[3] = L.bmax[0];
            p.b[0] = L.bmax[1]; p.b[1] = L.bmax[2]; p.b[2] = R.bmin[0]; p.b[3] = R.bmin[1];
            p.c[0] = R.bmin[2]; p.c[1] = R.bmax[0]; p.c[2] = R.bmax[1]; p.c[3] = R.bmax[2];
            p.left = ref_of((uint32_t)i + 1);
            p.right = ref_of(ln.offset);
            p.axis = ln.axis;
            p.pad = 0;
            if (odd[i]) continue;
            // two levels folded into one record (see QuadNode in bvh_build.hpp)
            QuadNode& Q = out->quads[quad_of[i]];
            for (int k = 0; k < 4; ++k) {
                Q.lox[k] = Q.loy[k] = Q.loz[k] = inf;
                Q.hix[k] = Q.hiy[k] = Q.hiz[k] = -inf;
                Q.ref[k] = kQuadEmpty;
                Q.pad[k] = 0;
            }
            uint32_t axes[3] = {ln.axis, 0u, 0u};
            const uint32_t kids[2] = {(uint32_t)i + 1, ln.offset};
            for (int g = 0; g < 2; ++g) {
                const LinearNode& X = nodes[kids[g]];
                uint32_t members[2];
                int n_members;
                if (X.n_prims > 0) { members[0] = kids[g]; n_members = 1; }
                else { members[0] = kids[g] + 1; members[1] = X.offset; n_members = 2; axes[1 + g] = X.axis; }
                for (int j = 0; j < n_members; ++j) {
                    const LinearNode& Y = nodes[members[j]];
                    const int k = 2 * g + j;
                    Q.lox[k] = Y.bmin[0]; Q.loy[k] = Y.bmin[1]; Q.loz[k] = Y.bmin[2];
                    Q.hix[k] = Y.bmax[0]; Q.hiy[k] = Y.bmax[1]; Q.hiz[k] = Y.bmax[2];
                    Q.ref[k] = Y.n_prims > 0 ? (kLeafBit | Y.offset) : quad_of[members[j]];
                }
            }
            for (int k = 0; k < 3; ++k) Q.ref[k] = (Q.ref[k] & kQuadRefMask) | (axes[k] << kQuadAxisShift);
        }
    });
    if (n_nodes) {
        out->root_ref = ref_of(0);
        out->quad_root_ref = nodes[0].n_prims == 0 ? 0u : out->root_ref;      // a single leaf has no record of either kind
        for (int k = 0; k < 3; ++k) { out->root_bounds[k] = nodes[0].bmin[k]; out->root_bounds[3 + k] = nodes[0].bmax[k]; }
    } else {
        out->root_ref = out->quad_root_ref = 0;
    }

    out->tris.resize(n_tris);
    parallel_ranges(n_tris, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            const uint32_t t = out->ordered_prims[i];
            PackedTri& pt = out->tris[i];
            if (t >= n_mesh) {                                   // an analytic sphere: its index in the sphere table rides in v0[0]
                const uint32_t k = (uint32_t)(t - n_mesh);
                std::memset(&pt, 0, sizeof pt);
                std::memcpy(&pt.v0[0], &k, 4);
                pt.prim_id = t;
                pt.pad = kPrimSphere;
                continue;
            }
            const float* p0 = verts + 3ull * indices[3ull * t];
            const float* p1 = verts + 3ull * indices[3ull * t + 1];
            const float* p2 = verts + 3ull * indices[3ull * t + 2];
            for (int k = 0; k < 3; ++k) { pt.v0[k] = p0[k]; pt.v1[k] = p1[k]; pt.v2[k] = p2[k]; }
            pt.prim_id = t;
            pt.last = 0;
            pt.pad = 0;
        }
    });
    parallel_ranges(n_nodes, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            const LinearNode& ln = nodes[i];
            if (ln.n_prims > 0) out->tris[(size_t)ln.offset + ln.n_prims - 1].last = 1;
        }
    });
}

}  // namespace pb2
