// api_internal.hpp — state behind the opaque handles of include/pbrt_b200.h.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/pbrt_b200.h"
#include "bvh_build.hpp"
#include "kernels.hpp"
#include "traverse.cuh"
#include "sphere.cuh"

namespace pb2 {

int set_error(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
int make_camera_view(const pb2_camera* cam, CameraView* out);

#define PB2_CUDA(call)                                                                  \
    do {                                                                                \
        cudaError_t pb2_e_ = (call);                                                    \
        if (pb2_e_ != cudaSuccess) return pb2::cuda_fail(pb2_e_, #call, __FILE__, __LINE__); \
    } while (0)

// Host-buffer batches stream through a ring of device slots on three streams — H2D copies, kernels, D2H copies — so the
// two copy engines and the SMs each run their own in-order queue and all three overlap.
struct Stage {                          // one ring slot
    void* d_in = nullptr;
    void* d_out = nullptr;
    void* d_aux = nullptr;
    cudaEvent_t in_ready = nullptr;     // H2D of the chunk in this slot finished
    cudaEvent_t done = nullptr;         // kernel finished
    cudaEvent_t drained = nullptr;      // D2H finished: the slot may be overwritten
    size_t cap = 0;
};
constexpr int kStages = 4;
struct Pipe {
    cudaStream_t h2d = nullptr, d2h = nullptr;
    cudaStream_t compute[2] = {nullptr, nullptr};   // alternate chunks: the tail of one chunk's kernel overlaps the next one
    Stage slot[kStages];
    uint64_t chunks = 0;                // chunks enqueued so far, over all calls: picks the slot and the compute stream
    bool pending = false;               // work enqueued by an _async entry point that pb2_scene_wait has not yet waited for
    // one event per _async batch, recorded behind its last D2H copy (pb2_scene_wait_until); a ring: before an event is reused
    // its previous batch is waited for, so at most kBatchRing batches are ever un-waited
    static constexpr int kBatchRing = 64;
    cudaEvent_t batch_done[kBatchRing] = {};
    uint64_t batches = 0;               // _async batches enqueued so far
};

struct Wavefront;
void wavefront_destroy(Wavefront* wf);

// Orders successive uses of one resource across CUDA streams (a film's accumulators; a scene's wavefront arena, counters and
// cached sampler tables): every entry point that touches the resource on stream `s` calls enter(s) before its first
// kernel / copy and leave(s) after its last one, so work enqueued by a later call — on whatever stream, the legacy default
// stream included — starts after the earlier call's work has finished.  (cudaStreamNonBlocking streams do not synchronise
// with the legacy stream, which the film's cudaMemset / cudaMemcpy calls use.)  Calls are enqueued under the owner's mutex,
// so enqueue order is call order.
struct UseChain {
    cudaEvent_t ev = nullptr;
    bool used = false;
    cudaError_t enter(cudaStream_t s) {
        if (!ev) {
            cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
        return used ? cudaStreamWaitEvent(s, ev, 0) : cudaSuccess;
    }
    cudaError_t leave(cudaStream_t s) {
        used = true;
        return cudaEventRecord(ev, s);
    }
    void destroy() {
        if (ev) cudaEventDestroy(ev);
        ev = nullptr;
        used = false;
    }
};

}  // namespace pb2

struct pb2_scene {
    std::vector<float> verts;
    std::vector<uint32_t> indices;
    std::vector<uint32_t> tri_material;
    std::vector<pb2_material> materials;
    std::vector<pb2_light> lights;
    // TriangleMesh's optional per-vertex normals / tangents / UVs (triangle.rs:17-26), pb2_scene_set_shading_geometry
    std::vector<float> normals, tangents, uvs;
    // analytic spheres (shapes/sphere.rs), pb2_scene_add_spheres: primitive ids n_tris .. n_tris + spheres.size() - 1
    std::vector<pb2_sphere> spheres;
    std::vector<pb2::DSphere> sphere_records;     // device layout (sphere.cuh), filled at build time
    void* d_spheres = nullptr;
    // participating media (pb2_scene_set_media)
    std::vector<pb2_medium> media;
    std::vector<int32_t> prim_inside, prim_outside;
    int32_t camera_medium = -1;
    bool has_material_less = false;     // some primitive carries PB2_NO_MATERIAL
    void* d_media = nullptr;
    void* d_prim_inside = nullptr;
    void* d_prim_outside = nullptr;
    uint64_t n_tris() const { return indices.size() / 3; }
    uint64_t n_primitives() const { return indices.size() / 3 + spheres.size(); }
    void* d_indices = nullptr;
    void* d_normals = nullptr;
    void* d_tangents = nullptr;
    void* d_uvs = nullptr;
    pb2::HostBVH bvh;
    double build_ms[6] = {0, 0, 0, 0, 0, 0};   // HLBVH stage times (bvh_build.hpp: build_hlbvh_gpu)
    bool built_host = false;   // LinearNode array + ordered prims valid
    bool built = false;        // device copies valid
    int device = 0;
    std::mutex mu;
    // device copies
    // device-built trees (HLBVH): the flattened reference-layout nodes / primitive order stay on the device until exported
    void* d_lin_nodes = nullptr;
    void* d_lin_prims = nullptr;
    uint64_t n_nodes = 0, n_prims = 0;
    int tree_depth = 0;
    float root_bounds[6] = {0, 0, 0, 0, 0, 0};
    void* d_pairs = nullptr;
    void* d_quads = nullptr;
    void* d_tris = nullptr;
    void* d_slot_of_prim = nullptr;
    void* d_tris_prim = nullptr;     // PackedTri + shading frame in primitive order (shading scenes; SceneView::tris_prim)
    void* d_tri_material = nullptr;
    void* d_tri_light = nullptr;
    void* d_materials = nullptr;
    void* d_lights = nullptr;
    void* d_light_cdf = nullptr;
    // work counters of the persistent traversal kernels: one slot per launch, handed out round-robin
    static constexpr unsigned kCounters = 256;
    unsigned long long* d_counters = nullptr;
    std::atomic<unsigned> counter_cursor{0};
    unsigned long long* next_counter() { return d_counters + (counter_cursor.fetch_add(1) % kCounters); }
    // Distribution1D of the light-selection strategies: [0] uniform, [1] power (lightdistrib.rs:26-69)
    std::vector<float> light_func[2], light_cdf[2];
    float light_func_int[2] = {0.0f, 0.0f};
    // SpatialLightDistribution tables (lightdistrib.rs:71-220), filled on the device the first time "spatial" is asked for:
    // [func n_vox * n | cdf n_vox * (n + 1) | func_int n_vox] floats in one allocation
    unsigned shading_class_mask = 7u;   // bit c: a material of shading class c exists (api_path.cu: upload_shading_tables)
    void* d_spatial = nullptr;
    int spatial_nv[3] = {0, 0, 0};
    pb2::SceneView view;
    pb2::Pipe pipe;
    pb2::Wavefront* wf = nullptr;
    pb2::UseChain path_chain;           // wf + d_tab1 / d_tab2 + d_spatial: shared by every pb2_render_path / pb2_path_li call
    // HaltonSampler tables on the device (built on first use; the scales depend on the sample-bounds extent)
    void* d_halton_perms = nullptr;
    void* d_halton_primes = nullptr;
    void* d_halton_sums = nullptr;
    void* d_sobol = nullptr;            // SobolSampler generator matrices: [m32 1024 x 52 u32 | vdc 25 x 52 u64 | vdc_inv 26 x 52 u64]
    // PixelSampler tables (stratified / (0,2)) and the parameters they were generated for
    void* d_tab1 = nullptr;
    void* d_tab2 = nullptr;
    long long tab_key[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint64_t counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    void free_device();
};

namespace pb2 {
int upload_shading_tables(pb2_scene* scene);
}
