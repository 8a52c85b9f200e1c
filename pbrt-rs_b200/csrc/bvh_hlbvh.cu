// bvh_hlbvh.cu — SplitMethod::HLBVH built on the GPU (sm_100a).
//
// Replaces BVHAccel::hlbvh_build / emit_lbvh / build_upper_sah (src/accelerators/bvh.rs:475-772) and its helpers
// left_shift3 / encode_morton3 / radix_sort (:137-197).  The Rust port does not run as written (its treelet loop never
// advances, leaves are written into an empty Vec, the bucket index is cast before it is scaled); the semantics here are
// pbrt-v3's HLBVHBuild, the algorithm the reference declares itself a port of, with the decisions H1-H8 of DESIGN.md §7
// (centroid bound accumulated, treelet loop advances, ordered prims sized up front, bucket = (12 * x) as usize, partition
// over [start, end) keeping b <= split bucket, leaf offsets = positions in Morton order, a zero centroid extent or an
// improper partition splits the range in the middle).  tests/ compare the result node for node with the CPU checker.
//
//   k_tri_bounds      triangle bounds + centroid bounds (BVHPrimitiveInfo, bvh.rs:26-41; Bounds3::union)
//   k_morton          10 bits per axis of the centroid's offset in the centroid bounds (:489-501)
//   cub radix sort    30-bit keys, stable — the same permutation as the reference's 5 x 6-bit LSD passes (:159-196)
//   k_treelet_flags   a treelet starts where the top 12 Morton bits change (:506-528); cub::DeviceSelect compacts them
//   k_emit_treelets   one thread per treelet walks its range top-down by Morton bit (:570-676) and writes the subtree in
//                     depth-first order (first child right behind its parent, as flatten_bvh_tree lays nodes out), then
//                     fills the boxes bottom-up
//   host              SAH over the <= 4096 treelet roots (:678-772), then the depth-first position of every treelet
//   k_place_*         treelet blocks and upper nodes copied to their final LinearBVHNode positions (:774-811)
//
// Every float operation is a separately rounded IEEE op in the reference's order (-fmad=false).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <limits>
#include <vector>

#include "bvh_build.hpp"
#include "pb2_math.cuh"

namespace pb2 {

namespace {

constexpr uint32_t kTreeletMask = 0x3FFC0000u;      // top 12 of the 30 Morton bits (bvh.rs:507)
constexpr int kFirstBit = 29 - 12;                  // bvh.rs:536

// ---- order-preserving float <-> uint encoding for atomicMin / atomicMax ----
__device__ __forceinline__ uint32_t f_enc(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
inline float f_dec_host(uint32_t e) {
    const uint32_t u = (e & 0x80000000u) ? (e & 0x7FFFFFFFu) : ~e;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

__global__ void __launch_bounds__(256) k_tri_bounds(const float* __restrict__ verts, const uint32_t* __restrict__ idx, uint32_t n,
                                                    float4* __restrict__ lo, float4* __restrict__ hi, uint32_t* __restrict__ cb /* 6 encoded */) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float c[3] = {0.f, 0.f, 0.f};
    const bool live = i < n;
    if (live) {
        const uint32_t i0 = idx[3ull * i], i1 = idx[3ull * i + 1], i2 = idx[3ull * i + 2];
        float l[3], h[3];
        for (int k = 0; k < 3; ++k) {
            const float a = verts[3ull * i0 + k], b = verts[3ull * i1 + k], d = verts[3ull * i2 + k];
            l[k] = fminf(fminf(a, b), d);                      // triangle.rs:175-180: Bounds3(p0, p1).union(p2)
            h[k] = fmaxf(fmaxf(a, b), d);
            c[k] = l[k] * 0.5f + h[k] * 0.5f;                  // bvh.rs:38
        }
        lo[i] = make_float4(l[0], l[1], l[2], 0.f);
        hi[i] = make_float4(h[0], h[1], h[2], 0.f);
    }
    // centroid bounds: warp reduce, then one atomic per CTA and component (one per warp serialised ~240 K same-address atomics
    // in L2 for 1.3 M triangles: 160 us at 4 % issue utilisation in ncu)
    __shared__ uint32_t s_mn[3][8], s_mx[3][8];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (int k = 0; k < 3; ++k) {
        uint32_t mn = live ? f_enc(c[k]) : 0xFFFFFFFFu, mx = live ? f_enc(c[k]) : 0u;
        for (int o = 16; o > 0; o >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, o));
            mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
        }
        if (lane == 0u) { s_mn[k][warp] = mn; s_mx[k][warp] = mx; }
    }
    __syncthreads();
    if (threadIdx.x < 3u) {
        const int k = (int)threadIdx.x;
        uint32_t mn = 0xFFFFFFFFu, mx = 0u;
        for (int w = 0; w < 8; ++w) { mn = min(mn, s_mn[k][w]); mx = max(mx, s_mx[k][w]); }
        atomicMin(&cb[k], mn);
        atomicMax(&cb[3 + k], mx);
    }
}

__device__ __forceinline__ uint32_t left_shift3(uint32_t x) {         // bvh.rs:138-152
    if (x == (1u << 10)) x -= 1;
    x = (x | (x << 16)) & 0x030000FFu;
    x = (x | (x << 8)) & 0x0300F00Fu;
    x = (x | (x << 4)) & 0x030C30C3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}
__device__ __forceinline__ uint32_t f2u_sat(float v) {                // Rust `as u32`: NaN -> 0, saturating
    if (!(v > 0.0f)) return 0u;
    if (v >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)v;
}

__global__ void __launch_bounds__(256) k_morton(const float4* __restrict__ lo, const float4* __restrict__ hi, uint32_t n, float mnx, float mny,
                                                float mnz, float mxx, float mxy, float mxz, uint32_t* __restrict__ codes,
                                                uint32_t* __restrict__ prim) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 l = lo[i], h = hi[i];
    float ox = (l.x * 0.5f + h.x * 0.5f) - mnx, oy = (l.y * 0.5f + h.y * 0.5f) - mny, oz = (l.z * 0.5f + h.z * 0.5f) - mnz;
    if (mxx > mnx) ox = ox / (mxx - mnx);                               // Bounds3::offset, geometry.rs:460-467
    if (mxy > mny) oy = oy / (mxy - mny);
    if (mxz > mnz) oz = oz / (mxz - mnz);
    const float scale = 1024.0f;                                        // 1 << morton_bits
    codes[i] = (left_shift3(f2u_sat(oz * scale)) << 2) | (left_shift3(f2u_sat(oy * scale)) << 1) | left_shift3(f2u_sat(ox * scale));
    prim[i] = i;
}

__global__ void __launch_bounds__(256) k_treelet_flags(const uint32_t* __restrict__ codes, uint32_t n, uint8_t* __restrict__ flags) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flags[i] = (i == 0u || ((codes[i] ^ codes[i - 1]) & kTreeletMask) != 0u) ? 1 : 0;
}

// emit_lbvh (bvh.rs:570-676) for all treelets at once.  pbrt's recursion is a deterministic function of the sorted Morton
// codes, so it is evaluated here level by level instead of depth first (one thread per treelet walking its whole subtree
// left 97 % of the machine idle: ncu, 2.7 % occupancy):
//   k_lbvh_split    one thread per node of the current level: skip the bits the range agrees on (:608-625), make a leaf
//                   (:583) or find the split by binary search (:627-638) and append its two children to the next level;
//   k_lbvh_up       one thread per leaf: box of the leaf, then up the tree — the second child to arrive at a node (atomic
//                   arrival counter) combines sizes and boxes and carries on;
//   k_lbvh_place    one thread per node: its position in the treelet's depth-first order is its depth plus the sizes of the
//                   first-child subtrees it hangs to the right of (walk up to the root), which is where emit_lbvh's
//                   recursion would have written it; writes the LinearNode there (in local[2 * s0 ...], as before).
// meta[t] = {node count, deepest level, error flag}.
struct TNode {
    uint32_t s, n;          // sorted range
    uint32_t parent;        // index into the TNode array, 0xFFFFFFFF for a treelet root (roots are nodes [0, n_treelets))
    uint32_t child;         // first child (second = child + 1); 0 for a leaf
    uint32_t size;          // nodes in the subtree
    int8_t bit;             // split bit (after skipping), -1..17
    uint8_t second;         // this node is its parent's second child
    uint8_t level;          // 0 = treelet root
    uint8_t leaf;
    float bmin[3], bmax[3];
};

__global__ void __launch_bounds__(256) k_lbvh_roots(const uint32_t* __restrict__ starts, uint32_t n_treelets, uint32_t n_prims, TNode* __restrict__ nodes) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_treelets) return;
    TNode nd;
    nd.s = starts[t];
    nd.n = ((t + 1 < n_treelets) ? starts[t + 1] : n_prims) - nd.s;
    nd.parent = 0xFFFFFFFFu; nd.child = 0u; nd.size = 0u;
    nd.bit = (int8_t)kFirstBit; nd.second = 0; nd.level = 0; nd.leaf = 0;
    for (int d = 0; d < 3; ++d) { nd.bmin[d] = 0.f; nd.bmax[d] = 0.f; }
    nodes[t] = nd;
}

// counters: [0] nodes appended so far (next level grows from `next_begin`), [1] error flag
__global__ void __launch_bounds__(256) k_lbvh_split(TNode* __restrict__ nodes, uint32_t level_begin, uint32_t level_count, uint32_t next_begin,
                                                    const uint32_t* __restrict__ codes, int max_prims, uint32_t* __restrict__ counters,
                                                    uint32_t* __restrict__ arrivals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= level_count) return;
    const uint32_t q = level_begin + i;
    TNode nd = nodes[q];
    int bit = nd.bit;
    // bits on which the whole range agrees make no node (:608-625); a short range or no bits left makes a leaf (:583)
    while (bit >= 0 && nd.n >= (uint32_t)max_prims && ((codes[nd.s] ^ codes[nd.s + nd.n - 1]) & (1u << bit)) == 0u) --bit;
    nd.bit = (int8_t)bit;
    arrivals[q] = 0u;
    if (bit < 0 || nd.n < (uint32_t)max_prims) {
        if (nd.n > 65535u) atomicOr(&counters[1], 1u);                  // LinearBVHNode::n_primitives is 16 bits (pbrt-v3 CHECKs)
        nd.leaf = 1;
        nd.child = 0u;
        nodes[q] = nd;
        return;
    }
    const uint32_t mask = 1u << bit;
    uint32_t a = 0, b = nd.n - 1;                                       // :627-638
    while (a + 1 != b) {
        const uint32_t mid = (a + b) / 2;
        if (((codes[nd.s + a] ^ codes[nd.s + mid]) & mask) == 0u) a = mid; else b = mid;
    }
    const uint32_t c = next_begin + atomicAdd(&counters[0], 2u);
    nd.leaf = 0;
    nd.child = c;
    nodes[q] = nd;
    TNode ch;
    ch.parent = q; ch.child = 0u; ch.size = 0u; ch.bit = (int8_t)(bit - 1); ch.level = (uint8_t)(nd.level + 1); ch.leaf = 0;
    for (int d = 0; d < 3; ++d) { ch.bmin[d] = 0.f; ch.bmax[d] = 0.f; }
    ch.s = nd.s; ch.n = b; ch.second = 0;
    nodes[c] = ch;
    ch.s = nd.s + b; ch.n = nd.n - b; ch.second = 1;
    nodes[c + 1] = ch;
}

__global__ void __launch_bounds__(256) k_lbvh_up(TNode* __restrict__ nodes, uint32_t n_nodes, const uint32_t* __restrict__ prim,
                                                 const float4* __restrict__ lo, const float4* __restrict__ hi, uint32_t* __restrict__ arrivals) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_nodes || !nodes[q].leaf) return;
    float l[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f}, h[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    {
        const uint32_t s = nodes[q].s, n = nodes[q].n;
        for (uint32_t j = 0; j < n; ++j) {
            const uint32_t p = prim[s + j];
            const float4 pl = lo[p], ph = hi[p];
            l[0] = fminf(l[0], pl.x); l[1] = fminf(l[1], pl.y); l[2] = fminf(l[2], pl.z);
            h[0] = fmaxf(h[0], ph.x); h[1] = fmaxf(h[1], ph.y); h[2] = fmaxf(h[2], ph.z);
        }
    }
    uint32_t size = 1u;
    for (;;) {
        TNode& nd = nodes[q];
        for (int d = 0; d < 3; ++d) { nd.bmin[d] = l[d]; nd.bmax[d] = h[d]; }
        nd.size = size;
        const uint32_t parent = nd.parent;
        if (parent == 0xFFFFFFFFu) return;
        __threadfence();
        if (atomicAdd(&arrivals[parent], 1u) == 0u) return;             // the sibling's thread finishes the parent
        __threadfence();
        const uint32_t sib = nd.second ? q - 1u : q + 1u;
        const volatile TNode& sb = nodes[sib];
        for (int d = 0; d < 3; ++d) { l[d] = fminf(l[d], sb.bmin[d]); h[d] = fmaxf(h[d], sb.bmax[d]); }
        size = size + sb.size + 1u;
        q = parent;
    }
}

__global__ void __launch_bounds__(256) k_lbvh_place(const TNode* __restrict__ nodes, uint32_t n_nodes, LinearNode* __restrict__ local,
                                                    uint4* __restrict__ meta) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_nodes) return;
    const TNode nd = nodes[q];
    // depth-first position inside the treelet: one step per level, plus every first-child subtree passed on the right
    uint32_t dfs = nd.level, up = q;
    for (uint32_t k = nd.level; k > 0; --k) {
        const TNode& cur = nodes[up];
        if (cur.second) dfs += nodes[up - 1u].size;
        up = cur.parent;
    }
    const uint32_t t = up;                                              // treelet roots are nodes [0, n_treelets)
    LinearNode out;
    for (int d = 0; d < 3; ++d) { out.bmin[d] = nd.bmin[d]; out.bmax[d] = nd.bmax[d]; }
    out.pad = (uint8_t)(nd.level & 1u);                                 // depth parity inside the treelet
    if (nd.leaf) {
        out.offset = nd.s;                                              // H7: position in the sorted order
        out.n_prims = (uint16_t)nd.n;
        out.axis = 0;
    } else {
        out.offset = dfs + 1u + nodes[nd.child].size;                   // second child, local index
        out.n_prims = 0;
        out.axis = (uint8_t)(nd.bit % 3);                               // :671
    }
    local[2ull * nodes[t].s + dfs] = out;
    atomicMax(&meta[t].y, (uint32_t)nd.level + 1u);
    if (q == t) meta[t].x = nd.size;
}

// Treelet t's block -> final[base[t] ...]; interior second-child indices become global.
__global__ void __launch_bounds__(256) k_place_treelets(const uint32_t* __restrict__ starts, const uint4* __restrict__ meta,
                                                       const uint32_t* __restrict__ base, const uint8_t* __restrict__ root_parity,
                                                       const LinearNode* __restrict__ local, LinearNode* __restrict__ final_nodes) {
    const uint32_t t = blockIdx.x;
    const LinearNode* src = local + 2ull * starts[t];
    const uint32_t count = meta[t].x, b = base[t];
    const uint8_t par = root_parity[t];
    for (uint32_t k = threadIdx.x; k < count; k += blockDim.x) {
        LinearNode nd = src[k];
        if (nd.n_prims == 0) nd.offset += b;
        nd.pad ^= par;                                                  // depth parity in the whole tree
        final_nodes[b + k] = nd;
    }
}
__global__ void __launch_bounds__(256) k_gather_roots(const uint32_t* __restrict__ starts, uint32_t n_treelets, const LinearNode* __restrict__ local,
                                                      LinearNode* __restrict__ roots) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_treelets) roots[t] = local[2ull * starts[t]];
}
// ---- device-layout repack: the host's repack_device_layout (bvh_build.cpp) as kernels over the flattened array --------------
__global__ void __launch_bounds__(256) k_record_flags(const LinearNode* __restrict__ nodes, uint32_t n, uint32_t* __restrict__ is_pair,
                                                      uint32_t* __restrict__ is_quad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool interior = nodes[i].n_prims == 0;
    is_pair[i] = interior ? 1u : 0u;
    is_quad[i] = (interior && nodes[i].pad == 0) ? 1u : 0u;             // interior node at even depth
}
__device__ __forceinline__ uint32_t node_ref(const LinearNode& ln, uint32_t index_of) {
    return ln.n_prims > 0 ? (kLeafBit | ln.offset) : index_of;
}
__global__ void __launch_bounds__(256) k_fill_records(const LinearNode* __restrict__ nodes, uint32_t n, const uint32_t* __restrict__ pair_of,
                                                      const uint32_t* __restrict__ quad_of, PairNode* __restrict__ pairs,
                                                      QuadNode* __restrict__ quads) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const LinearNode ln = nodes[i];
    if (ln.n_prims != 0) return;
    const LinearNode L = nodes[i + 1], R = nodes[ln.offset];
    PairNode p;
    p.a[0] = L.bmin[0]; p.a[1] = L.bmin[1]; p.a[2] = L.bmin[2]; p.a[3] = L.bmax[0];
    p.b[0] = L.bmax[1]; p.b[1] = L.bmax[2]; p.b[2] = R.bmin[0]; p.b[3] = R.bmin[1];
    p.c[0] = R.bmin[2]; p.c[1] = R.bmax[0]; p.c[2] = R.bmax[1]; p.c[3] = R.bmax[2];
    p.left = node_ref(L, pair_of[i + 1]);
    p.right = node_ref(R, pair_of[ln.offset]);
    p.axis = ln.axis;
    p.pad = 0;
    pairs[pair_of[i]] = p;
    if (ln.pad != 0) return;
    QuadNode Q;
    const float inf = __int_as_float(0x7f800000);
    for (int k = 0; k < 4; ++k) {
        Q.lox[k] = Q.loy[k] = Q.loz[k] = inf;
        Q.hix[k] = Q.hiy[k] = Q.hiz[k] = -inf;
        Q.ref[k] = kQuadEmpty;
        Q.pad[k] = 0;
    }
    uint32_t axes[3] = {ln.axis, 0u, 0u};
    const uint32_t kids[2] = {i + 1, ln.offset};
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        const LinearNode X = g == 0 ? L : R;
        if (X.n_prims > 0) {
            const int k = 2 * g;
            Q.lox[k] = X.bmin[0]; Q.loy[k] = X.bmin[1]; Q.loz[k] = X.bmin[2];
            Q.hix[k] = X.bmax[0]; Q.hiy[k] = X.bmax[1]; Q.hiz[k] = X.bmax[2];
            Q.ref[k] = kLeafBit | X.offset;
        } else {
            axes[1 + g] = X.axis;
            const uint32_t members[2] = {kids[g] + 1, X.offset};
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const LinearNode Y = nodes[members[j]];
                const int k = 2 * g + j;
                Q.lox[k] = Y.bmin[0]; Q.loy[k] = Y.bmin[1]; Q.loz[k] = Y.bmin[2];
                Q.hix[k] = Y.bmax[0]; Q.hiy[k] = Y.bmax[1]; Q.hiz[k] = Y.bmax[2];
                Q.ref[k] = node_ref(Y, quad_of[members[j]]);
            }
        }
    }
    for (int k = 0; k < 3; ++k) Q.ref[k] = (Q.ref[k] & kQuadRefMask) | (axes[k] << kQuadAxisShift);
    quads[quad_of[i]] = Q;
}
__global__ void __launch_bounds__(256) k_fill_tris(const float* __restrict__ verts, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ prim,
                                                   uint32_t n, PackedTri* __restrict__ tris, uint32_t* __restrict__ slot_of_prim) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t t = prim[i];
    PackedTri pt;
    const uint32_t i0 = idx[3ull * t], i1 = idx[3ull * t + 1], i2 = idx[3ull * t + 2];
    for (int k = 0; k < 3; ++k) { pt.v0[k] = verts[3ull * i0 + k]; pt.v1[k] = verts[3ull * i1 + k]; pt.v2[k] = verts[3ull * i2 + k]; }
    pt.prim_id = t;
    pt.last = 0;
    pt.pad = 0;
    tris[i] = pt;
    slot_of_prim[t] = i;
}
__global__ void __launch_bounds__(256) k_mark_last(const LinearNode* __restrict__ nodes, uint32_t n, PackedTri* __restrict__ tris) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const LinearNode ln = nodes[i];
    if (ln.n_prims > 0) tris[(size_t)ln.offset + ln.n_prims - 1].last = 1;
}

struct UpperNode {
    uint32_t position;
    LinearNode node;
};
__global__ void k_place_upper(const UpperNode* __restrict__ up, uint32_t n, LinearNode* __restrict__ final_nodes) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) final_nodes[up[i].position] = up[i].node;
}

// ---- host: SAH over the treelet roots (bvh.rs:678-772, pbrt-v3 buildUpperSAH; H4-H6, H8) ------------------------------
struct HBox {
    float mn[3], mx[3];
};
inline HBox empty_box() {
    const float m = std::numeric_limits<float>::max();
    return {{m, m, m}, {-m, -m, -m}};
}
inline HBox join(const HBox& a, const HBox& b) {
    HBox r;
    for (int k = 0; k < 3; ++k) { r.mn[k] = std::fmin(a.mn[k], b.mn[k]); r.mx[k] = std::fmax(a.mx[k], b.mx[k]); }
    return r;
}
inline float area(const HBox& b) {                                     // geometry.rs:667-670
    const float dx = b.mx[0] - b.mn[0], dy = b.mx[1] - b.mn[1], dz = b.mx[2] - b.mn[2];
    return 2.0f * ((dx * dy + dx * dz) + dy * dz);
}
struct UpperTree {
    struct N { HBox box; int child[2]; int axis; };                    // child >= 0: upper node, < 0: treelet ~child
    std::vector<N> nodes;
    const std::vector<HBox>* roots;
    const HBox& box_of(int ref) const { return ref >= 0 ? nodes[ref].box : (*roots)[~ref]; }
    int build(std::vector<int>& order, size_t start, size_t end) {     // order: treelet ids
        if (end - start == 1) return ~order[start];
        const int index = (int)nodes.size();
        nodes.emplace_back();
        HBox bounds = empty_box(), cb = empty_box();
        for (size_t i = start; i < end; ++i) bounds = join(bounds, (*roots)[order[i]]);
        for (size_t i = start; i < end; ++i) {
            const HBox& b = (*roots)[order[i]];
            HBox p;
            for (int k = 0; k < 3; ++k) p.mn[k] = p.mx[k] = (b.mn[k] + b.mx[k]) * 0.5f;          // :703
            cb = join(cb, p);
        }
        const float dx = cb.mx[0] - cb.mn[0], dy = cb.mx[1] - cb.mn[1], dz = cb.mx[2] - cb.mn[2];
        const int dim = (dx > dy && dx > dz) ? 0 : ((dy > dz) ? 1 : 2);                          // geometry.rs:482-485
        size_t mid = (start + end) / 2;
        if (cb.mx[dim] != cb.mn[dim]) {
            constexpr int kBuckets = 12;
            int count[kBuckets] = {0};
            HBox bb[kBuckets];
            for (auto& b : bb) b = empty_box();
            auto bucket_of = [&](int treelet) {
                const HBox& b = (*roots)[treelet];
                const float c = (b.mn[dim] + b.mx[dim]) * 0.5f;
                int k = (int)((float)kBuckets * ((c - cb.mn[dim]) / (cb.mx[dim] - cb.mn[dim])));
                if (k == kBuckets) k = kBuckets - 1;
                return k;
            };
            for (size_t i = start; i < end; ++i) {
                const int k = bucket_of(order[i]);
                count[k]++;
                bb[k] = join(bb[k], (*roots)[order[i]]);
            }
            float cost[kBuckets - 1];
            for (int i = 0; i < kBuckets - 1; ++i) {
                HBox b0 = empty_box(), b1 = empty_box();
                int c0 = 0, c1 = 0;
                for (int j = 0; j <= i; ++j) { b0 = join(b0, bb[j]); c0 += count[j]; }
                for (int j = i + 1; j < kBuckets; ++j) { b1 = join(b1, bb[j]); c1 += count[j]; }
                cost[i] = 0.125f + ((float)c0 * area(b0) + (float)c1 * area(b1)) / area(bounds);   // :736
            }
            float best = std::numeric_limits<float>::max();
            int split = 0;
            for (int i = 0; i < kBuckets - 1; ++i)
                if (cost[i] < best) { best = cost[i]; split = i; }
            size_t lo = start, hi = end;                                // Rust partition_in_place over [start, end)
            for (;;) {
                while (lo < hi && bucket_of(order[lo]) <= split) ++lo;
                if (lo == hi) break;
                do { --hi; } while (lo < hi && !(bucket_of(order[hi]) <= split));
                if (lo == hi) break;
                std::swap(order[lo], order[hi]);
                ++lo;
            }
            if (lo > start && lo < end) mid = lo;
        }
        const int c0 = build(order, start, mid);
        const int c1 = build(order, mid, end);
        N& n = nodes[index];
        n.box = join(box_of(c0), box_of(c1));
        n.child[0] = c0;
        n.child[1] = c1;
        n.axis = dim;
        return index;
    }
};

struct Flattener {
    const UpperTree& ut;
    const std::vector<uint32_t>& counts;
    const std::vector<uint32_t>& depths;
    std::vector<uint32_t> base;
    std::vector<uint8_t> root_parity;                                   // depth parity of every treelet root
    std::vector<UpperNode> upper;
    uint32_t next = 0;
    int deepest = 0;
    uint32_t visit(int ref, int level) {                               // flatten_bvh_tree, bvh.rs:774-811
        if (ref < 0) {
            const int t = ~ref;
            const uint32_t my = next;
            base[t] = my;
            root_parity[t] = (uint8_t)((level - 1) & 1);
            next += counts[t];
            deepest = std::max(deepest, level - 1 + (int)depths[t]);
            return my;
        }
        const uint32_t my = next++;
        visit(ut.nodes[ref].child[0], level + 1);
        const uint32_t second = visit(ut.nodes[ref].child[1], level + 1);
        UpperNode u;
        u.position = my;
        for (int k = 0; k < 3; ++k) { u.node.bmin[k] = ut.nodes[ref].box.mn[k]; u.node.bmax[k] = ut.nodes[ref].box.mx[k]; }
        u.node.offset = second;
        u.node.n_prims = 0;
        u.node.axis = (uint8_t)ut.nodes[ref].axis;
        u.node.pad = (uint8_t)((level - 1) & 1);
        upper.push_back(u);
        return my;
    }
};

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
    template <class T> T* as() { return (T*)p; }
};

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

#define HL_CUDA(call)                                                                                       \
    do {                                                                                                    \
        cudaError_t e_ = (call);                                                                            \
        if (e_ != cudaSuccess) {                                                                            \
            snprintf(err, err_len, "CUDA error %s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return -1;                                                                                      \
        }                                                                                                   \
    } while (0)

int build_hlbvh_gpu(const float* verts, uint64_t n_verts, const uint32_t* indices, uint64_t n_tris, int max_prims_in_node, DeviceBVH* out,
                    char* err, int err_len, double* timing_ms) {
    *out = DeviceBVH();
    double tm[6] = {0, 0, 0, 0, 0, 0};
    if (n_tris == 0) return 0;
    const int max_prims = std::min(max_prims_in_node, 255);             // bvh.rs:222
    const uint32_t n = (uint32_t)n_tris;
    const unsigned grid = (n + 255) / 256;
    double t0 = now_ms();
    DevBuf d_verts, d_idx, d_lo, d_hi, d_cb, d_codes, d_prim, d_codes2, d_prim2, d_flags, d_starts, d_nsel, d_tmp, d_local, d_meta;
    HL_CUDA(d_verts.alloc(n_verts * 12));
    HL_CUDA(d_idx.alloc(n_tris * 12));
    HL_CUDA(d_lo.alloc(n_tris * 16));
    HL_CUDA(d_hi.alloc(n_tris * 16));
    HL_CUDA(d_cb.alloc(24));
    HL_CUDA(d_codes.alloc(n_tris * 4));
    HL_CUDA(d_prim.alloc(n_tris * 4));
    HL_CUDA(d_codes2.alloc(n_tris * 4));
    HL_CUDA(d_prim2.alloc(n_tris * 4));
    HL_CUDA(d_flags.alloc(n_tris));
    HL_CUDA(d_starts.alloc(n_tris < 4096 ? n_tris * 4 : 4096 * 4));
    HL_CUDA(d_nsel.alloc(4));
    HL_CUDA(cudaMemcpy(d_verts.p, verts, n_verts * 12, cudaMemcpyHostToDevice));
    HL_CUDA(cudaMemcpy(d_idx.p, indices, n_tris * 12, cudaMemcpyHostToDevice));
    const uint32_t cb_init[6] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u};
    HL_CUDA(cudaMemcpy(d_cb.p, cb_init, 24, cudaMemcpyHostToDevice));
    k_tri_bounds<<<grid, 256>>>(d_verts.as<float>(), d_idx.as<uint32_t>(), n, d_lo.as<float4>(), d_hi.as<float4>(), d_cb.as<uint32_t>());
    uint32_t cb_enc[6];
    HL_CUDA(cudaMemcpy(cb_enc, d_cb.p, 24, cudaMemcpyDeviceToHost));
    float cb[6];
    for (int k = 0; k < 6; ++k) cb[k] = f_dec_host(cb_enc[k]);
    k_morton<<<grid, 256>>>(d_lo.as<float4>(), d_hi.as<float4>(), n, cb[0], cb[1], cb[2], cb[3], cb[4], cb[5], d_codes.as<uint32_t>(),
                            d_prim.as<uint32_t>());
    HL_CUDA(cudaDeviceSynchronize());
    tm[0] = now_ms() - t0;

    t0 = now_ms();
    size_t tmp_sort = 0, tmp_sel = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, d_codes.as<uint32_t>(), d_codes2.as<uint32_t>(), d_prim.as<uint32_t>(),
                                    d_prim2.as<uint32_t>(), (int)n, 0, 30);
    thrust::counting_iterator<uint32_t> counting(0u);
    cub::DeviceSelect::Flagged(nullptr, tmp_sel, counting, d_flags.as<uint8_t>(), d_starts.as<uint32_t>(), d_nsel.as<uint32_t>(), (int)n);
    HL_CUDA(d_tmp.alloc(std::max(tmp_sort, tmp_sel)));
    HL_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_sort, d_codes.as<uint32_t>(), d_codes2.as<uint32_t>(), d_prim.as<uint32_t>(),
                                            d_prim2.as<uint32_t>(), (int)n, 0, 30));
    HL_CUDA(cudaDeviceSynchronize());
    tm[1] = now_ms() - t0;

    t0 = now_ms();
    const uint32_t* codes = d_codes2.as<uint32_t>();
    const uint32_t* prim = d_prim2.as<uint32_t>();
    k_treelet_flags<<<grid, 256>>>(codes, n, d_flags.as<uint8_t>());
    HL_CUDA(cub::DeviceSelect::Flagged(d_tmp.p, tmp_sel, counting, d_flags.as<uint8_t>(), d_starts.as<uint32_t>(), d_nsel.as<uint32_t>(), (int)n));
    uint32_t n_treelets = 0;
    HL_CUDA(cudaMemcpy(&n_treelets, d_nsel.p, 4, cudaMemcpyDeviceToHost));
    if (n_treelets == 0 || n_treelets > 4096) { snprintf(err, err_len, "HLBVH: %u treelets (expected 1..4096)", n_treelets); return -1; }
    HL_CUDA(d_local.alloc(2ull * n_tris * sizeof(LinearNode)));
    HL_CUDA(d_meta.alloc(n_treelets * sizeof(uint4)));
    HL_CUDA(cudaMemset(d_meta.p, 0, n_treelets * sizeof(uint4)));
    {
        // level-synchronous emit_lbvh: at most 2 * n - 1 nodes per treelet; a level holds the children appended by the one before
        DevBuf d_tn, d_arr, d_cnt;
        const uint64_t cap = 2ull * n_tris + n_treelets;
        HL_CUDA(d_tn.alloc(cap * sizeof(TNode)));
        HL_CUDA(d_arr.alloc(cap * sizeof(uint32_t)));
        HL_CUDA(d_cnt.alloc(2 * sizeof(uint32_t)));
        HL_CUDA(cudaMemset(d_cnt.p, 0, 2 * sizeof(uint32_t)));
        k_lbvh_roots<<<(n_treelets + 255) / 256, 256>>>(d_starts.as<uint32_t>(), n_treelets, n, d_tn.as<TNode>());
        uint32_t level_begin = 0, level_count = n_treelets, total = n_treelets;
        while (level_count > 0) {
            k_lbvh_split<<<(level_count + 255) / 256, 256>>>(d_tn.as<TNode>(), level_begin, level_count, n_treelets, codes, max_prims,
                                                           d_cnt.as<uint32_t>(), d_arr.as<uint32_t>());
            uint32_t appended = 0;
            HL_CUDA(cudaMemcpy(&appended, d_cnt.p, 4, cudaMemcpyDeviceToHost));          // nodes appended after the roots so far
            level_begin += level_count;
            level_count = n_treelets + appended - total;
            total = n_treelets + appended;
        }
        uint32_t flag = 0;
        HL_CUDA(cudaMemcpy(&flag, (const char*)d_cnt.p + 4, 4, cudaMemcpyDeviceToHost));
        if (flag) { snprintf(err, err_len, "HLBVH: a leaf holds more than 65535 primitives with identical Morton codes"); return -1; }
        k_lbvh_up<<<(total + 255) / 256, 256>>>(d_tn.as<TNode>(), total, prim, d_lo.as<float4>(), d_hi.as<float4>(), d_arr.as<uint32_t>());
        k_lbvh_place<<<(total + 255) / 256, 256>>>(d_tn.as<TNode>(), total, d_local.as<LinearNode>(), d_meta.as<uint4>());
        HL_CUDA(cudaGetLastError());
        HL_CUDA(cudaDeviceSynchronize());
    }
    std::vector<uint4> meta(n_treelets);
    HL_CUDA(cudaMemcpy(meta.data(), d_meta.p, n_treelets * sizeof(uint4), cudaMemcpyDeviceToHost));
    DevBuf d_roots;
    HL_CUDA(d_roots.alloc(n_treelets * sizeof(LinearNode)));
    k_gather_roots<<<(n_treelets + 255) / 256, 256>>>(d_starts.as<uint32_t>(), n_treelets, d_local.as<LinearNode>(), d_roots.as<LinearNode>());
    std::vector<LinearNode> root_nodes(n_treelets);
    HL_CUDA(cudaMemcpy(root_nodes.data(), d_roots.p, n_treelets * sizeof(LinearNode), cudaMemcpyDeviceToHost));
    std::vector<HBox> roots(n_treelets);
    std::vector<uint32_t> counts(n_treelets), depths(n_treelets);
    for (uint32_t t = 0; t < n_treelets; ++t) {
        if (meta[t].z) { snprintf(err, err_len, "HLBVH: a leaf holds more than 65535 primitives with identical Morton codes"); return -1; }
        counts[t] = meta[t].x;
        depths[t] = meta[t].y;
        for (int k = 0; k < 3; ++k) { roots[t].mn[k] = root_nodes[t].bmin[k]; roots[t].mx[k] = root_nodes[t].bmax[k]; }
    }
    tm[2] = now_ms() - t0;

    t0 = now_ms();
    UpperTree ut;
    ut.roots = &roots;
    ut.nodes.reserve(n_treelets);
    std::vector<int> order(n_treelets);
    for (uint32_t t = 0; t < n_treelets; ++t) order[t] = (int)t;
    const int root_ref = ut.build(order, 0, n_treelets);
    Flattener fl{ut, counts, depths, std::vector<uint32_t>(n_treelets, 0), std::vector<uint8_t>(n_treelets, 0), {}, 0, 0};
    fl.visit(root_ref, 1);
    const uint32_t total = fl.next;
    tm[3] = now_ms() - t0;

    t0 = now_ms();
    DevBuf d_final, d_base, d_par, d_upper;
    HL_CUDA(d_final.alloc((size_t)total * sizeof(LinearNode)));
    HL_CUDA(d_base.alloc(n_treelets * 4));
    HL_CUDA(d_par.alloc(n_treelets));
    HL_CUDA(cudaMemcpy(d_base.p, fl.base.data(), n_treelets * 4, cudaMemcpyHostToDevice));
    HL_CUDA(cudaMemcpy(d_par.p, fl.root_parity.data(), n_treelets, cudaMemcpyHostToDevice));
    k_place_treelets<<<n_treelets, 256>>>(d_starts.as<uint32_t>(), d_meta.as<uint4>(), d_base.as<uint32_t>(), d_par.as<uint8_t>(),
                                         d_local.as<LinearNode>(), d_final.as<LinearNode>());
    if (!fl.upper.empty()) {
        HL_CUDA(d_upper.alloc(fl.upper.size() * sizeof(UpperNode)));
        HL_CUDA(cudaMemcpy(d_upper.p, fl.upper.data(), fl.upper.size() * sizeof(UpperNode), cudaMemcpyHostToDevice));
        k_place_upper<<<((unsigned)fl.upper.size() + 255) / 256, 256>>>(d_upper.as<UpperNode>(), (uint32_t)fl.upper.size(), d_final.as<LinearNode>());
    }
    HL_CUDA(cudaGetLastError());
    HL_CUDA(cudaDeviceSynchronize());
    // the big scratch buffers are not needed any more
    cudaFree(d_local.p); d_local.p = nullptr;
    cudaFree(d_lo.p); d_lo.p = nullptr;
    cudaFree(d_hi.p); d_hi.p = nullptr;
    tm[4] = now_ms() - t0;

    // ---- traversal layout, on the device ----
    t0 = now_ms();
    const unsigned ngrid = (total + 255) / 256;
    DevBuf d_is_pair, d_is_quad, d_pair_of, d_quad_of, d_scan_tmp;
    HL_CUDA(d_is_pair.alloc((size_t)total * 4));
    HL_CUDA(d_is_quad.alloc((size_t)total * 4));
    HL_CUDA(d_pair_of.alloc((size_t)total * 4));
    HL_CUDA(d_quad_of.alloc((size_t)total * 4));
    k_record_flags<<<ngrid, 256>>>(d_final.as<LinearNode>(), total, d_is_pair.as<uint32_t>(), d_is_quad.as<uint32_t>());
    size_t tmp_scan = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, d_is_pair.as<uint32_t>(), d_pair_of.as<uint32_t>(), (int)total);
    HL_CUDA(d_scan_tmp.alloc(tmp_scan));
    HL_CUDA(cub::DeviceScan::ExclusiveSum(d_scan_tmp.p, tmp_scan, d_is_pair.as<uint32_t>(), d_pair_of.as<uint32_t>(), (int)total));
    HL_CUDA(cub::DeviceScan::ExclusiveSum(d_scan_tmp.p, tmp_scan, d_is_quad.as<uint32_t>(), d_quad_of.as<uint32_t>(), (int)total));
    uint32_t last[4];
    HL_CUDA(cudaMemcpy(&last[0], d_is_pair.as<uint32_t>() + (total - 1), 4, cudaMemcpyDeviceToHost));
    HL_CUDA(cudaMemcpy(&last[1], d_pair_of.as<uint32_t>() + (total - 1), 4, cudaMemcpyDeviceToHost));
    HL_CUDA(cudaMemcpy(&last[2], d_is_quad.as<uint32_t>() + (total - 1), 4, cudaMemcpyDeviceToHost));
    HL_CUDA(cudaMemcpy(&last[3], d_quad_of.as<uint32_t>() + (total - 1), 4, cudaMemcpyDeviceToHost));
    out->n_pairs = last[0] + last[1];
    out->n_quads = last[2] + last[3];
    out->n_nodes = total;
    out->n_tris = n_tris;
    HL_CUDA(cudaMalloc(&out->d_pairs, std::max<size_t>(64, out->n_pairs * sizeof(PairNode))));
    HL_CUDA(cudaMalloc(&out->d_quads, std::max<size_t>(128, out->n_quads * sizeof(QuadNode))));
    HL_CUDA(cudaMalloc(&out->d_tris, n_tris * sizeof(PackedTri)));
    HL_CUDA(cudaMalloc(&out->d_slot_of_prim, n_tris * 4));
    k_fill_records<<<ngrid, 256>>>(d_final.as<LinearNode>(), total, d_pair_of.as<uint32_t>(), d_quad_of.as<uint32_t>(), (PairNode*)out->d_pairs,
                                   (QuadNode*)out->d_quads);
    k_fill_tris<<<grid, 256>>>(d_verts.as<float>(), d_idx.as<uint32_t>(), prim, n, (PackedTri*)out->d_tris, (uint32_t*)out->d_slot_of_prim);
    k_mark_last<<<ngrid, 256>>>(d_final.as<LinearNode>(), total, (PackedTri*)out->d_tris);
    HL_CUDA(cudaGetLastError());
    LinearNode root;
    HL_CUDA(cudaMemcpy(&root, d_final.p, sizeof root, cudaMemcpyDeviceToHost));
    for (int k = 0; k < 3; ++k) { out->root_bounds[k] = root.bmin[k]; out->root_bounds[3 + k] = root.bmax[k]; }
    out->root_ref = root.n_prims > 0 ? (kLeafBit | root.offset) : 0u;
    out->quad_root_ref = out->root_ref;                                 // both numberings start at the root
    out->max_depth = fl.deepest;
    HL_CUDA(cudaDeviceSynchronize());
    // the flattened nodes and the primitive order stay on the device for pb2_bvh_export
    out->d_nodes = d_final.p; d_final.p = nullptr;
    out->d_ordered_prims = d_prim2.p; d_prim2.p = nullptr;
    tm[5] = now_ms() - t0;
    if (timing_ms) for (int k = 0; k < 6; ++k) timing_ms[k] = tm[k];
    return 0;
}

}  // namespace pb2
