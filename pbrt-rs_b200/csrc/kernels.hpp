// kernels.hpp — host-visible launchers of the device kernels (internal; the public surface is include/pbrt_b200.h).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "pb2_math.cuh"

namespace pb2 {

struct SceneView;

// PerspectiveCamera state (perspective.rs:23-31): raster_to_camera and camera_to_world, computed on the host.
struct CameraView {
    mat4 raster_to_camera;
    mat4 camera_to_world;
    int res_x, res_y;
    float lens_radius, focal_distance;    // thin lens (perspective.rs:101-107); lens_radius <= 0: pinhole
};

void launch_closest_hit(const SceneView& s, const void* d_rays, uint64_t n, void* d_hits, void* d_b0, unsigned long long* d_counter,
                        cudaStream_t st);
void launch_any_hit(const SceneView& s, const void* d_rays, uint64_t n, void* d_out, unsigned long long* d_counter, cudaStream_t st);
void launch_camera_rays(const CameraView& cam, const void* d_pfilm, const void* d_plens, uint64_t n, void* d_rays, cudaStream_t st);
void launch_spawn_shadow(const SceneView& s, const void* d_rays, const void* d_hits, uint64_t n, const float light[3],
                         void* d_out, cudaStream_t st);
void launch_spawn_bounce(const SceneView& s, const void* d_rays, const void* d_hits, uint64_t n, void* d_out, cudaStream_t st);
void launch_spawn_shadow_bounce(const SceneView& s, const void* d_rays, const void* d_hits, uint64_t n, const float light[3], void* d_out_shadow,
                                void* d_out_bounce, cudaStream_t st);
void launch_tris_by_prim(const void* d_tris, uint64_t n, void* d_out, cudaStream_t st);
void launch_mark_degenerate(void* d_tris, uint64_t n, const void* d_indices, const void* d_uvs, cudaStream_t st);
// Matrix4x4::inverse (transform.rs:46-113; camera_host.cpp)
mat4 invert_mat4(const mat4& m);
void set_trace_tuning(int refill_below, int node_quorum, int leaf_quorum, int prefetch);
void launch_rng_floats(uint64_t first_seq, uint32_t n_seq, uint32_t n_per, float* d_out, cudaStream_t st);

}  // namespace pb2
