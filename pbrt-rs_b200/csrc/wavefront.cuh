// wavefront.cuh — data layout of the wavefront PathIntegrator (SoA path state + stage queues in HBM).
//
// One path slot per (pixel, sample) of the current batch; slot i = s_local * n_pixels + pixel, so a warp holds
// neighbouring pixels of one sample index.  Every array below is indexed by slot (coalesced float4 / uint4 accesses);
// stages talk to each other only through these arrays and the uint32 slot queues.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.hpp"
#include "shade.cuh"
#include "traverse.cuh"

namespace pb2 {

// SpatialLightDistribution (lightdistrib.rs:77-81) as filled by light_distrib.cu; func == nullptr: not selected.
struct SpatialView {
    const float* func;            // [voxel][n_lights], voxel = (z * nv[1] + y) * nv[0] + x
    const float* cdf;             // [voxel][n_lights + 1]
    const float* func_int;        // [voxel]
    int nv[3];                    // n_voxel
    float lo[3], hi[3];           // scene.world_bound()
};
// SpatialLightDistribution::lookup (lightdistrib.rs:165-175): Bounds3::offset (geometry.rs:460-467), `as i32` (saturating,
// NaN -> 0: cvt.rzi.s32.f32 does the same), clamp to the grid.
__device__ __forceinline__ size_t spatial_voxel(const SpatialView& g, vec3 p) {
    float ox = p.x - g.lo[0], oy = p.y - g.lo[1], oz = p.z - g.lo[2];
    if (g.hi[0] > g.lo[0]) ox = ox / (g.hi[0] - g.lo[0]);
    if (g.hi[1] > g.lo[1]) oy = oy / (g.hi[1] - g.lo[1]);
    if (g.hi[2] > g.lo[2]) oz = oz / (g.hi[2] - g.lo[2]);
    const int ix = min(max(__float2int_rz(ox * (float)g.nv[0]), 0), g.nv[0] - 1);
    const int iy = min(max(__float2int_rz(oy * (float)g.nv[1]), 0), g.nv[1] - 1);
    const int iz = min(max(__float2int_rz(oz * (float)g.nv[2]), 0), g.nv[2] - 1);
    return ((size_t)iz * (size_t)g.nv[1] + (size_t)iy) * (size_t)g.nv[0] + (size_t)ix;
}

// Device-side scene tables for shading (indexed by the caller's primitive id).
struct ShadeView {
    const uint32_t* tri_material;
    const int32_t* tri_light;
    const DMaterial* mats;
    const DLight* lights;
    int n_lights;
    unsigned class_mask;          // bit c: some material of the scene has shading class c (k_shade<c> is launched only then)
    const float* light_func;      // Distribution1D func[n_lights]
    const float* light_cdf;       // cdf[n_lights + 1]
    float light_func_int;
    SpatialView spatial;          // "spatial" strategy with more than one light: the vertex's voxel replaces the three above
    // TriangleMesh's optional attributes (triangle.rs:17-26), all null for a plain mesh
    const uint32_t* indices;      // 3 vertex ids per caller triangle
    const float* normals;         // 3 per vertex
    const float* tangents;        // 3 per vertex
    const float2* uvs;            // per vertex
    // participating media (media/homogeneous.rs) and every primitive's MediumInterface as indices into `media` (-1 = none);
    // all null / -1 unless pb2_scene_set_media was called
    const DMedium* media;
    const int32_t* prim_inside;
    const int32_t* prim_outside;
    int camera_medium;
    int has_interfaces;           // some primitive has no material (a medium interface): paths and transmittance rays cross it
};

// Film geometry: image, sample bounds (film.rs:76-81 with D42), filter radius and its 16x16 table (film.rs:53-63).
struct FilmView {
    int px0, py0, px1, py1;       // cropped_pixel_bounds (film.rs:41-50): the pixels the film stores, row-major
    int sb_x0, sb_y0, sb_w, sb_h;
    unsigned long long sb_w_magic;   // fast_div_magic(sb_w): pixel -> row without a 32-bit division (slot_info runs per path vertex)
    float max_lum;                // max_sample_luminance (film.rs:259-261), +inf when unset
    float radius_x, radius_y;
    int exact;                    // box filter, r = 0.5: ordered accumulation (bit-reproducible)
    const float* table;           // 256 floats
    float4* acc;                  // per image pixel: RGB sum + filter weight sum of the current render call
    float4* xyzw;                 // per image pixel: X, Y, Z, weight accumulators (the Film itself)
    float4* splat;                // per image pixel: splat_xyz (film.rs:9-15), null until the first Film::add_splat
    unsigned long long* stray_keys;
    float4* stray_vals;
    uint32_t stray_capacity;
    __host__ __device__ int width() const { return px1 - px0; }
    __host__ __device__ size_t n_pixels() const { return (size_t)(px1 - px0) * (size_t)(py1 - py0); }
    __host__ __device__ size_t index(int x, int y) const { return (size_t)(y - py0) * (size_t)(px1 - px0) + (size_t)(x - px0); }
};

// HaltonSampler state shared by all paths (samplers/halton.rs:24-37 + the permutation table of lowdiscrepancy.rs:333-349);
// perms == nullptr selects the RandomSampler streams.
struct SamplerView {
    int kind;                     // PB2_SAMPLER_*: 0 random, 1 Halton, 2 stratified, 3 (0,2)-sequence, 4 Sobol'
    // PixelSampler tables (sampler.rs:257-322) of kinds 2 / 3, written by k_pixel_tables: value of tabulated dimension d,
    // sample s, pixel p at [(d * spp + s) * tab_n_pix + p] (a warp = neighbouring pixels of one sample: coalesced)
    int n_dims;
    int spp_tab;                  // samples per pixel of the tables
    uint32_t tab_n_pix;
    const float* t1;
    const float2* t2;
    const uint16_t* perms;        // RADICAL_INVERSE_PERMUTATIONS
    const uint32_t* primes;       // first 1000 primes
    const uint32_t* prime_sums;   // offset of every base's permutation
    int base_scales[2], base_exponents[2];
    unsigned long long sample_stride, mult_inverse[2];
    // SobolSampler (samplers/sobol.rs:13-37): generator matrices (core/sobolmatrices.rs; 52 entries per dimension / resolution),
    // sample_bounds.min, resolution = the sample-bounds extent rounded up to a power of two
    const uint32_t* sobol_m32;
    const unsigned long long* sobol_vdc;
    const unsigned long long* sobol_vdc_inv;
    int sobol_min[2];
    int sobol_resolution, sobol_log2_resolution;
};

// Exact n / d for 32-bit n, d through one 64 x 64 -> high 64 multiply: m = ceil(2^64 / d), n / d = hi64(n * m) (the error term
// n * (m d - 2^64) / (d 2^64) is below 2^-32 <= 1 / d, so the floor is unchanged); d = 1 has no 64-bit m and is flagged by m = 0.
inline unsigned long long fast_div_magic(uint32_t d) { return d <= 1u ? 0ull : 0xFFFFFFFFFFFFFFFFull / d + 1ull; }
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t fast_div(uint32_t n, unsigned long long magic) {
    return magic ? (uint32_t)__umul64hi((unsigned long long)n, magic) : n;
}
#endif

// slot -> (pixel, sample index)
struct PathMap {
    uint32_t n_pix;               // sample-bounds pixels per sample index
    int spp;
    int sample0;
    const int32_t* explicit_xy;   // pb2_path_li: explicit pixel coordinates, else null
    const uint32_t* explicit_s;
    SamplerView smp;
    unsigned long long n_pix_magic;   // fast_div_magic(n_pix)
};

enum Counter : int {
    C_ACTIVE_A = 0, C_ACTIVE_B = 1, C_MAT0 = 2, C_MAT1 = 3, C_MAT2 = 4, C_SHADOW = 5, C_MIS = 6, C_MIS_PREV = 7,
    C_WORK_EXTEND = 8, C_WORK_SHADOW = 9, C_WORK_MIS = 10, C_STRAYS = 11, C_STRAY_OVERFLOW = 12,
    T_CAMERA = 16, T_EXTEND = 17, T_SHADOW = 18, T_MIS = 19, T_LAUNCHES = 20,
    // wavefront VolPathIntegrator: the transmittance rays that cross an interface and go on (two buffers each), unused select outputs
    C_VOL_S_ALT = 24, C_VOL_M_ALT = 25, C_VOL_SCRATCH = 26, C_COUNT = 32
};

struct PathBuffers {
    float4* ray_o;        // xyz origin, w = t_max
    float4* ray_d;        // xyz direction
    float4* beta;         // rgb throughput, w = eta_scale
    float4* L;            // rgb radiance, w = bits: bounces | specular_bounce << 16
    unsigned long long* rng;
    uint4* hit;           // prim id, b0, b1, b2 (float bits)
    // next-event estimation record of the current bounce
    float4* sh_o;         // shadow ray origin, w = t_max
    float4* sh_d;         // shadow ray direction, w = light pick pdf
    float4* t1;           // light-sampling term (rgb), w = bits: 1 = shadow ray pending, 2 = MIS ray pending
    float4* mis_o;        // MIS ray origin
    float4* mis_d;        // MIS ray direction
    float4* t2;           // BSDF-sampling term (rgb), w = light primitive id bits
    float4* beta_nee;     // throughput before the bounce
    uint8_t* occluded;
    uint32_t* mis_prim;
    uint32_t* q_active[2];
    uint32_t* q_mat[3];
    uint32_t* q_shadow;
    uint32_t* q_mis;
    unsigned long long* counters;
    // order-preserving queues: what became of the path in `slot` this bounce — bits 0-1 shading class of its hit (3 = no hit),
    // bit 2 continues, bit 3 shadow ray pending, bit 4 MIS ray pending — and the tile descriptors of the single-pass select that rebuilds the queues
    uint8_t* state;
    unsigned long long* select_status;     // k_select3: per tile and output, {epoch, flag, count or inclusive prefix}
    unsigned* select_tickets;              // two tile counters, used by alternate launches
};

// What the wavefront VolPathIntegrator (wavefront_volpath.cu) keeps per path slot beside PathBuffers; allocated at its first use.
// In that mode the NEE record holds the factors of estimate_direct's two terms unmultiplied (the transmittance enters the
// product before f, integrator.rs:172-173, 259-261): t1 = {li, w1}, f1 = {f |cos|, light_pdf (negated: delta light)},
// t2 = {lmis * f2, light primitive}, mis_o.w = scattering_pdf, mis_d.w = w2, sh_d.w = pick pdf, beta_nee = beta at the vertex.
struct VolBuffers {
    int32_t* med;         // medium of the path ray (-1 = vacuum)
    float* t_path;        // where the path ray's walk ended: the closest hit's distance, or the ray's own t_max
    float4* f1;
    float4* p1;           // VisibilityTester's far end (light.rs:137-160): every segment of the shadow ray is re-aimed at it
    float4* p1_err;
    float4* p1_n;
    float4* tr_s;         // transmittance gathered by the shadow ray so far (rgb)
    float4* tr_m;         // ... by the MIS ray (Scene::intersect_tr, scene.rs:48-71); w = 1: it ended on a surface with a material
    int2* ray_med;        // medium of the current segment: x shadow ray, y MIS ray
    uint4* hit_s;         // closest hit of the current segment (prim, b0 | t for a sphere, b1, b2)
    uint4* hit_m;
    float* t_s;
    float* t_m;
    uint8_t* st_s;        // after the walk: 1 = a hit; after k_vol_tr_post: 0 = another segment follows, 3 = done
    uint8_t* st_m;
};

struct PathParams {
    int max_depth;
    float rr_threshold;
    int integrator;               // PB2_INTEGRATOR_PATH (wavefront) / PB2_INTEGRATOR_VOLPATH (k_volpath)
};

struct Wavefront {
    uint64_t capacity = 0;
    PathBuffers b{};
    void* arena = nullptr;
    int sm_count = 0;
    unsigned select_launches = 0;          // k_select3 launches so far: picks the ticket counter and the epoch
    VolBuffers vol{};                      // wavefront VolPathIntegrator state (vol_arena; null until the first volpath render)
    void* vol_arena = nullptr;
    unsigned long long* h_counters = nullptr;   // pinned: queue counts read back between volpath iterations (interface scenes only)
    // A frame of several batches alternates them between this wavefront on the caller's stream and `peer` (a second arena of the
    // same size, created at the first such frame) on `aux_stream`, so that one batch's launches fill the ramps and tails of the
    // other's persistent kernels; the film accumulation of the batches stays in batch order (ev_acc).
    Wavefront* peer = nullptr;
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_acc[2] = {nullptr, nullptr};
    uint64_t totals[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

int wavefront_create(uint64_t capacity, Wavefront** out);
int wavefront_render(Wavefront* wf, const SceneView& sv, const ShadeView& sh, const CameraView& cam, const FilmView& film,
                     const PathParams& pp, const SamplerView& smp, int spp, int sample_begin, int sample_end, cudaStream_t st);
int wavefront_li(Wavefront* wf, const SceneView& sv, const ShadeView& sh, const CameraView& cam, const FilmView& film,
                 const PathParams& pp, const SamplerView& smp, int spp, const int32_t* d_xy, const uint32_t* d_s, uint64_t n, float* d_L,
                 float* d_pfilm, cudaStream_t st);
// Sampler::start_pixel of StratifiedSampler / ZeroTwoSequenceSampler for every pixel (one thread per pixel, stream
// RNG::new(seq0 + pixel)).
void pixel_tables_generate(int kind, uint32_t n_pix, int spp, int n_dims, int x_samples, int y_samples, int jitter, uint64_t seq0,
                           float* d_t1, float2* d_t2, cudaStream_t st);
// SpatialLightDistribution (light_distrib.cu)
void spatial_grid_extents(const float wb[6], int max_voxels, int nv[3]);
void spatial_distribution_build(const SpatialView& grid, const DLight* d_lights, int n_lights, float* d_func, float* d_cdf, float* d_func_int,
                                cudaStream_t st, const void* d_spheres = nullptr);
void film_finish(const FilmView& film, unsigned long long* counters, cudaStream_t st);
void film_add_samples(const FilmView& film, const float* d_pfilm, const float* d_L, const float* d_w, uint64_t n, cudaStream_t st);
void film_resolve(const FilmView& film, float scale, float splat_scale, float* d_rgb, cudaStream_t st);
void film_add_splats(const FilmView& film, const float* d_pfilm, const float* d_v, uint64_t n, cudaStream_t st);
void film_set_image(const FilmView& film, const float* d_rgb, cudaStream_t st);

}  // namespace pb2
