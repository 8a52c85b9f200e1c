// bvh_build.hpp — host-side SAH BVH build for the B200 backend.
//
// Replaces BVHAccel::new / recursive_build / flatten_bvh_tree (src/accelerators/bvh.rs:216-473, :774-811) for
// SplitMethod::SAH.  The build stays on the host (north-star); it emits the flattened depth-first node array
// directly (first child at i+1, second-child offset patched after the first subtree), runs independent
// subtrees on worker threads, and then repacks into the device layout:
//   * PairNode (64 B): one record per INTERIOR node holding both children's boxes, so one 64-byte fetch feeds
//     two slab tests and leaves need no node record at all;
//   * QuadNode (128 B, one cache line): two levels of the binary tree folded into one record — the boxes of the (up to
//     four) grandchildren of an interior node as six float4 (SoA), their references and the three split axes that
//     fix the reference's visiting order; a ray takes half as many dependent fetches to reach a leaf;
//   * PackedTri (48 B): triangles gathered in BVH leaf order as three float4, w-lanes carry the caller's
//     primitive id and an end-of-leaf flag.
#pragma once
#include <cstdint>
#include <vector>

namespace pb2 {

// The reference's LinearBVHNode (bvh.rs:129-135) narrowed to pbrt-v3's 32 bytes.
struct LinearNode {
    float bmin[3];
    float bmax[3];
    uint32_t offset;      // leaf: first index into ordered prims; interior: second child
    uint16_t n_prims;     // 0 = interior
    uint8_t axis;
    uint8_t pad;
};
static_assert(sizeof(LinearNode) == 32, "LinearNode must be 32 bytes");

constexpr uint32_t kLeafBit = 0x80000000u;

// Device node: boxes of the left (i+1) and right (second child) children of one interior node.
//   a = {L.min.x, L.min.y, L.min.z, L.max.x}
//   b = {L.max.y, L.max.z, R.min.x, R.min.y}
//   c = {R.min.z, R.max.x, R.max.y, R.max.z}
//   d = {L.ref, R.ref, axis, 0}      ref = pair index, or kLeafBit | first triangle slot
struct alignas(64) PairNode {
    float a[4], b[4], c[4];
    uint32_t left, right, axis, pad;
};
static_assert(sizeof(PairNode) == 64, "PairNode must be 64 bytes");

// Device node of the traversal kernels: interior node P with children A (= P+1) and B (= second child) folded with
// A's and B's own children.  Slots 0,1 belong to A, slots 2,3 to B: an interior child contributes its two children
// (first child in the lower slot), a leaf child occupies the lower slot of its group and leaves the other one empty
// (inverted box, ref = kQuadEmpty).  ref = quad index, or kLeafBit | first triangle slot, in bits 0-28 and 31; bits 29-30
// of ref[0], ref[1], ref[2] carry axis(P), axis(A), axis(B) (bvh.rs:856-866 decides near/far by the sign of the ray
// direction on these axes), so a visit reads 112 bytes: three 256-bit loads (boxes) and one 128-bit load (references).
struct alignas(128) QuadNode {
    float lox[4], loy[4], loz[4], hix[4], hiy[4], hiz[4];
    uint32_t ref[4];
    uint32_t pad[4];
};
static_assert(sizeof(QuadNode) == 128, "QuadNode must be 128 bytes");
constexpr uint32_t kQuadEmpty = 0xFFFFFFFFu;
constexpr uint32_t kQuadAxisShift = 29, kQuadRefMask = 0x9FFFFFFFu;      // index space: 2^29 quads / triangle slots

struct alignas(16) PackedTri {
    float v0[3];
    uint32_t prim_id;     // index into the caller's triangle list
    float v1[3];
    uint32_t last;        // 1 = last triangle of its leaf
    float v2[3];
    uint32_t pad;
};
static_assert(sizeof(PackedTri) == 48, "PackedTri must be 48 bytes");

struct HostBVH {
    std::vector<LinearNode> nodes;          // reference layout (parity export)
    std::vector<uint32_t> ordered_prims;    // leaf order -> caller triangle id
    int max_depth = 0;                      // nodes on the longest root-to-leaf path
    bool leaf_overflow = false;             // a leaf holds more primitives than LinearNode::n_prims (16 bits) can count
    // device layout
    std::vector<PairNode> pairs;
    std::vector<QuadNode> quads;
    std::vector<PackedTri> tris;
    uint32_t root_ref = 0;                  // into pairs
    uint32_t quad_root_ref = 0;             // into quads
    float root_bounds[6] = {0, 0, 0, 0, 0, 0};
};

// verts: 3 floats per vertex; indices: 3 per triangle.  threads <= 0 -> hardware concurrency.  split_method
// (bvh.rs:199-204): 0 = SplitMethod::SAH, 2 = ::Middle, 3 = ::EqualCounts (the top-down recursive_build, bvh.rs:273-473).
// extra_bounds / n_extra: primitives that are not triangles (analytic spheres, shapes/sphere.rs), given by their world bounds
// (6 floats each: min.xyz, max.xyz = Shape::world_bound); they follow the triangles in the primitive list (ids n_tris ..
// n_tris + n_extra - 1) and take one PackedTri slot each with pad bit 1 set and v0[0] = bits of their index (kPrimSphere).
void build_sah_bvh(const float* verts, uint64_t n_verts, const uint32_t* indices, uint64_t n_tris,
                   int max_prims_in_node, int threads, HostBVH* out, int split_method = 0, const float* extra_bounds = nullptr,
                   uint64_t n_extra = 0);

// Fills pairs / quads / tris / root refs / root bounds of `out` from its nodes + ordered_prims.
void repack_device_layout(const float* verts, const uint32_t* indices, uint64_t n_tris, HostBVH* out, uint64_t n_extra = 0);
constexpr uint32_t kPrimDegenerate = 1u, kPrimSphere = 2u;      // PackedTri::pad bits

// What a device-side build hands over: the traversal layout already in device memory (ownership passes to the caller,
// cudaFree each pointer) plus the flattened reference-layout nodes and primitive order, kept on the device and only
// copied to the host when someone asks for them (pb2_bvh_export).
struct DeviceBVH {
    void* d_pairs = nullptr;          // PairNode[n_pairs]
    void* d_quads = nullptr;          // QuadNode[n_quads]
    void* d_tris = nullptr;           // PackedTri[n_tris]
    void* d_slot_of_prim = nullptr;   // uint32[n_tris]
    void* d_nodes = nullptr;          // LinearNode[n_nodes], depth-first (pad = depth parity)
    void* d_ordered_prims = nullptr;  // uint32[n_tris]
    uint64_t n_pairs = 0, n_quads = 0, n_nodes = 0, n_tris = 0;
    uint32_t root_ref = 0, quad_root_ref = 0;
    float root_bounds[6] = {0, 0, 0, 0, 0, 0};
    int max_depth = 0;
};

// SplitMethod::HLBVH (bvh.rs:475-772) on the GPU: Morton codes, radix sort, one LBVH treelet per 12-bit Morton prefix,
// SAH over the treelet roots, flatten, repack into the traversal layout — bvh_hlbvh.cu.  Nothing but the <= 4096 treelet
// roots visits the host.  Returns 0, or -1 with a message in *err.  timing_ms (optional, 6 doubles): upload + bounds +
// Morton, sort, treelets, upper SAH (host), flatten, device-layout repack.
int build_hlbvh_gpu(const float* verts, uint64_t n_verts, const uint32_t* indices, uint64_t n_tris, int max_prims_in_node,
                    DeviceBVH* out, char* err, int err_len, double* timing_ms);

}  // namespace pb2
