/* pbrt_b200.h — C ABI of the B200-native ray-tracing backend for the pbrt-rs hot path.
 *
 * Every entry point is `extern "C"`, takes plain pointers / sizes / POD structs and returns an int status
 * (0 = PB2_OK, <0 = error; message via pb2_last_error()).  Nothing aborts; there is NO CPU fallback — a missing
 * GPU or CUDA failure is an error.  Each function names the reference interface (file:line under
 * /root/reference) it replaces.  The Rust-side binding is shown in INTEGRATION.md and rust_shim/.
 *
 * Buffers: "host" pointers are caller-owned host memory (copied in/out synchronously).  "_device" variants take
 * device pointers plus a CUDA stream (cudaStream_t passed as void*) and are asynchronous on that stream.
 */
#ifndef PBRT_B200_H
#define PBRT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB2_OK 0
#define PB2_ERR_INVALID -1   /* bad argument */
#define PB2_ERR_CUDA -2      /* CUDA runtime failure / no device */
#define PB2_ERR_STATE -3     /* call order (e.g. intersect before build_bvh) */
#define PB2_ERR_NCCL -4
#define PB2_ERR_LIMIT -5     /* BVH deeper than the 64-entry traversal stack (bvh.rs:839) */

#define PB2_MISS 0xFFFFFFFFu

/* src/core/geometry.rs:756-763 `Ray {o, d, t_max, time}` — 32 bytes, two float4 loads on the device. */
typedef struct pb2_ray {
    float o[3];
    float t_max;
    float d[3];
    float time;
} pb2_ray;

/* What `Primitive::intersect` leaves behind (src/core/primitive.rs:65-78, src/shapes/triangle.rs:182-316):
 * prim_id = index into the caller's triangle list (PB2_MISS when nothing was hit), t = the shrunk ray.t_max,
 * (b1, b2) = barycentrics e1*inv_det, e2*inv_det.  16 bytes. */
typedef struct pb2_hit {
    uint32_t prim_id;
    float t;
    float b1;
    float b2;
} pb2_hit;

/* pbrt-v3 matte / plastic / glass / mirror / metal over the reference's BxDF blocks (the reference's src/materials/ files are
 * empty files): LambertianReflection or OrenNayar (reflection.rs:821-855, 917-971), MicrofacetReflection +
 * TrowbridgeReitz + FresnelDielectric / FresnelConductor (:977-1056, :571-604, :42-69), FresnelSpecular (:733-819),
 * SpecularReflection + FresnelNoOp (:606-659), MicrofacetTransmission (:1058-1192), FresnelBlend (:1194-1280; substrate: kd, ks, roughness). */
enum { PB2_MAT_MATTE = 0, PB2_MAT_PLASTIC = 1, PB2_MAT_GLASS = 2, PB2_MAT_MIRROR = 3, PB2_MAT_METAL = 4, PB2_MAT_SUBSTRATE = 5 };
typedef struct pb2_material {
    int32_t type;
    float kd[3];        /* matte, plastic, substrate: diffuse reflectance */
    float ks[3];        /* plastic, substrate: glossy reflectance */
    float roughness;    /* plastic, metal; glass: 0 = smooth (FresnelSpecular), > 0 = rough (MicrofacetReflection + MicrofacetTransmission) */
    int32_t remap_roughness;
    float kr[3];        /* glass, mirror */
    float kt[3];        /* glass */
    float eta;          /* glass index of refraction */
    float sigma;        /* matte: Oren-Nayar roughness in degrees, clamped to [0, 90]; 0 = Lambertian */
    float metal_eta[3]; /* metal: conductor index of refraction ... */
    float metal_k[3];   /* ... and absorption coefficient, per RGB channel */
} pb2_material;

/* src/lights/point.rs PointLight; src/lights/diffuse.rs DiffuseAreaLight (one per emissive triangle);
 * src/lights/spot.rs SpotLight; src/lights/distant.rs DistantLight. */
enum { PB2_LIGHT_POINT = 0, PB2_LIGHT_AREA = 1, PB2_LIGHT_SPOT = 2, PB2_LIGHT_DISTANT = 3 };
typedef struct pb2_light {
    int32_t type;
    float p[3];         /* point, spot: position (light_to_world * origin, point.rs:30, spot.rs:37) */
    float i[3];         /* point, spot: intensity I; area: emitted radiance L_emit; distant: radiance L */
    uint32_t prim_id;   /* area: the emissive triangle */
    int32_t two_sided;  /* area */
    float axis[3];      /* spot: third row of world_to_light's 3x3 block, the vector SpotLight::falloff projects onto
                         * (spot.rs:51-53) = normalize(to - from) for pbrt's from/to spot light;
                         * distant: the direction w TOWARDS the light (distant.rs:31; normalised by the library) */
    float total_width;  /* spot: cone half-angle in degrees (spot.rs:38) */
    float falloff_start;/* spot: half-angle where the falloff starts, degrees (spot.rs:39) */
} pb2_light;

/* src/shapes/sphere.rs:229-248 Sphere::new(object_to_world, world_to_object, reverse_orientation, radius, z_min, z_max, phi_max)
 * behind a GeometricPrimitive: an analytic sphere (EFloat quadratic, src/core/efloat.rs), optionally cut by z_min / z_max and
 * phi_max.  object_to_world is a row-major 4x4 AFFINE matrix (last row 0 0 0 1; rotation, non-uniform scale and translation are
 * fine, a projective row is PB2_ERR_INVALID); world_to_object is its inverse, computed by the library the way Transform::new
 * does (transform.rs:46-113).  phi_max in degrees.  88 bytes. */
typedef struct pb2_sphere {
    float object_to_world[16];
    float radius, z_min, z_max, phi_max;
    int32_t reverse_orientation;
    uint32_t material;          /* index into the scene's materials (ignored for pure ray casting) */
} pb2_sphere;

/* src/media/homogeneous.rs:20-28 HomogeneousMedium::new(sigma_a, sigma_s, g): absorption and scattering coefficients per RGB
 * channel (per unit length of the scene) and the Henyey-Greenstein asymmetry of its phase function (src/core/medium.rs:34-87). */
typedef struct pb2_medium {
    float sigma_a[3];
    float sigma_s[3];
    float g;
} pb2_medium;
/* tri_material[i] / pb2_sphere.material of a surface that has no material (GeometricPrimitive { material: None }): it only
 * separates two media, rays pass through it (volpath.rs:127-131).  VolPathIntegrator only. */
#define PB2_NO_MATERIAL 0xFFFFFFFFu

/* src/cameras/perspective.rs:34-82 PerspectiveCamera + Transform::look_at. */
typedef struct pb2_camera {
    float pos[3];
    float look[3];
    float up[3];
    float fov;          /* degrees, applies to the shorter image axis */
    int32_t res_x, res_y;   /* the film's full resolution */
    float lens_radius;  /* thin lens (perspective.rs:101-107); 0 = pinhole */
    float focal_distance;
} pb2_camera;

enum { PB2_FILTER_BOX = 0, PB2_FILTER_GAUSSIAN = 1, PB2_FILTER_TRIANGLE = 2, PB2_FILTER_MITCHELL = 3, PB2_FILTER_SINC = 4 };
/* src/core/film.rs:31-75 Film::new + src/filters/{boxf,gaussian,triangle,mitchell,sinc}.rs */
typedef struct pb2_film_desc {
    int32_t res_x, res_y;       /* full_resolution */
    int32_t filter;             /* PB2_FILTER_* */
    float radius_x, radius_y;
    float gaussian_alpha;       /* GaussianFilter::new(radius, alpha) */
    float mitchell_b, mitchell_c;   /* MitchellFilter::new(radius, b, c) */
    float sinc_tau;             /* LanczosSincFilter::new(radius, tau) */
    float crop_window[4];       /* {min.x, min.y, max.x, max.y} as fractions of the full resolution (film.rs:33,41-50);
                                 * all zero means the whole image {0, 0, 1, 1}.  The film stores, and pb2_film_read_xyzw /
                                 * pb2_film_resolve_rgb / pb2_film_write_image return, only cropped_pixel_bounds. */
    float max_sample_luminance; /* FilmTile::add_sample clamp (film.rs:259-261); <= 0 means infinity */
} pb2_film_desc;

/* create_light_sample_distribution (src/core/lightdistrib.rs:222-232): "uniform", "power", "spatial".  PB2_LIGHTS_SPATIAL is
 * SpatialLightDistribution (lightdistrib.rs:71-220): one Distribution1D per voxel of a grid over the world bound (64 voxels
 * along its longest axis), each from 128 Halton points x every light's sample_li.  The reference fills voxels lazily into a
 * hash table; this backend fills the whole grid on the device the first time a render asks for it (k_spatial_contrib,
 * k_spatial_distrib) — the same distributions, so the same paths.  A scene with one light uses "uniform" (:223). */
enum { PB2_LIGHTS_UNIFORM = 0, PB2_LIGHTS_POWER = 1, PB2_LIGHTS_SPATIAL = 2 };
/* PB2_SAMPLER_RANDOM: RandomSampler (src/samplers/random.rs), one PCG32 stream per (pixel, sample).
 * PB2_SAMPLER_HALTON: HaltonSampler (src/samplers/halton.rs, src/core/lowdiscrepancy.rs:293-390) — every dimension is a pure
 * function of (pixel, sample index, dimension), so the per-sample values equal the reference's tile-ordered render. */
/* PB2_SAMPLER_STRATIFIED / PB2_SAMPLER_ZEROTWO: StratifiedSampler (src/samplers/stratified.rs), ZeroTwoSequenceSampler
 * (src/samplers/zerotwosequence.rs) — PixelSamplers (src/core/sampler.rs:257-322): the first n_sampled_dimensions 1D and 2D
 * dimensions of every pixel come from per-pixel tables of spp values generated on the device by Sampler::start_pixel from
 * the stream RNG::new(n_pixels*spp + pixel); later dimensions fall back to the per-(pixel, sample) PCG32 stream. */
/* PB2_SAMPLER_SOBOL: SobolSampler (src/samplers/sobol.rs, src/core/lowdiscrepancy.rs:507-560, generator matrices of
 * src/core/sobolmatrices.rs) — a GlobalSampler like Halton: sample index = sobol_interval_to_index(log2 of the sample-bounds
 * extent rounded up to a power of two, sample number, pixel - sample_bounds.min); dimension d = index bits x the d-th
 * generator matrix; dimensions 0 / 1 are remapped into the pixel.  spp must be a power of two (sobol.rs:22-28 rounds up).
 * The reference leaves sample_dimension as todo!() (sobol.rs:56-58) and its sobol_interval_to_index cannot terminate
 * (lowdiscrepancy.rs:529-534); both follow pbrt-v3 here. */
enum { PB2_SAMPLER_RANDOM = 0, PB2_SAMPLER_HALTON = 1, PB2_SAMPLER_STRATIFIED = 2, PB2_SAMPLER_ZEROTWO = 3, PB2_SAMPLER_SOBOL = 4 };
/* PB2_INTEGRATOR_PATH: PathIntegrator (src/integrators/path.rs), the wavefront pipeline.  PB2_INTEGRATOR_VOLPATH:
 * VolPathIntegrator (src/integrators/volpath.rs:60-244) over HomogeneousMedium — medium sampling, Henyey-Greenstein phase
 * vertices, next-event estimation with transmittance (VisibilityTester::tr, src/core/light.rs:137-160) and MIS through
 * Scene::intersect_tr (src/core/scene.rs:48-71), as wavefront stages of their own (wavefront_volpath.cu); samplers: random,
 * stratified, (0,2) (the number of dimensions a volumetric path draws is unbounded, which the 1000 / 1024-dimension Halton /
 * Sobol' tables are not).  A scene with material-less interface surfaces is rendered with host read-backs of queue counts between
 * stages (crossing an interface is not a bounce, so the number of iterations is not known in advance): pb2_render_path then
 * returns when the frame is done instead of when it is enqueued. */
enum { PB2_INTEGRATOR_PATH = 0, PB2_INTEGRATOR_VOLPATH = 1 };
/* src/integrators/path.rs:31-46 PathIntegrator::new + src/samplers/random.rs:17-27 RandomSampler::new */
typedef struct pb2_path_desc {
    int32_t max_depth;
    float rr_threshold;
    int32_t light_strategy;     /* PB2_LIGHTS_* (src/core/lightdistrib.rs:222-232) */
    int32_t spp;                /* samples per pixel of the whole frame */
    int32_t sample_begin;       /* this call renders sample indices [sample_begin, sample_end) of every pixel */
    int32_t sample_end;
    int32_t sampler;            /* PB2_SAMPLER_* */
    int32_t n_sampled_dimensions; /* stratified, (0,2): PixelSampler::new (sampler.rs:268); pbrt's default is 4 */
    int32_t x_samples, y_samples; /* stratified: StratifiedSampler::new (stratified.rs:23-39); spp must equal x_samples * y_samples */
    int32_t jitter;             /* stratified: jitter_samples */
    int32_t integrator;         /* PB2_INTEGRATOR_* */
} pb2_path_desc;

typedef struct pb2_scene pb2_scene;
typedef struct pb2_film pb2_film;

/* ---- runtime ------------------------------------------------------------------------------------------ */
int pb2_init(int device);                 /* cudaSetDevice + capability check (sm_100 required) */
int pb2_shutdown(void);
const char* pb2_last_error(void);         /* thread-local message of the last failing call */
int pb2_device_count(int* out);
/* Pinned host memory for the host-buffer entry points (lets their H2D / D2H copies overlap the kernels), raw
 * device memory + blocking copies for callers of the *_device entry points that have no CUDA binding of their own. */
int pb2_host_alloc(uint64_t bytes, void** out);
int pb2_host_free(void* p);
int pb2_device_alloc(uint64_t bytes, void** out);
int pb2_device_free(void* p);
int pb2_memcpy_h2d(void* dst_device, const void* src_host, uint64_t bytes);
int pb2_memcpy_d2h(void* dst_host, const void* src_device, uint64_t bytes);
int pb2_device_synchronize(void);
/* Scheduling knobs of the traversal kernels (lanes refilled below / node-step quorum / leaf-step quorum / L2 prefetch of
 * deferred children); a negative value keeps the current setting.  Results never depend on them (tuning sweeps only;
 * no reference counterpart). */
int pb2_set_trace_tuning(int refill_below, int node_quorum, int leaf_quorum, int prefetch);

/* ---- scene + BVHAccel (src/accelerators/bvh.rs:216-271 BVHAccel::new) ---------------------------------- */
/* Copies the mesh.  tri_material (index into mats) / tri_light (index into lights or -1) may be NULL for pure
 * ray casting.  Replaces the Vec<PrimitiveDt> of GeometricPrimitive(Triangle) handed to BVHAccel::new. */
int pb2_scene_create(const float* verts, uint64_t n_verts, const uint32_t* indices, uint64_t n_tris,
                     const uint32_t* tri_material, const pb2_material* mats, uint32_t n_mats,
                     const pb2_light* lights, uint32_t n_lights, pb2_scene** out);
/* TriangleMesh's optional per-vertex shading normals `n`, tangents `s` and parametric coordinates `uv`
 * (src/shapes/triangle.rs:17-26; 3 / 3 / 2 floats per vertex, world space, NULL = absent), as Triangle::intersect (:251-311),
 * Triangle::get_uvs (:60-72) and Triangle::sample (:338-341) use them.  Call before pb2_scene_build_bvh. */
int pb2_scene_set_shading_geometry(pb2_scene* scene, const float* normals, const float* tangents, const float* uvs);
/* Appends n analytic spheres to the scene's primitive list, after the triangles: sphere k of the scene has primitive id
 * n_tris + k (the id pb2_hit.prim_id reports and pb2_light.prim_id names for a spherical DiffuseAreaLight).  Call before
 * pb2_scene_build_bvh; SplitMethod::HLBVH (the device build) takes triangles only and refuses a scene with spheres.  A closest
 * hit on a sphere reports b1 = u = phi / phi_max and b2 = v = (theta - theta_min) / (theta_max - theta_min) (sphere.rs:49-52);
 * pb2_intersect's optional b0 is 0 there. */
int pb2_scene_add_spheres(pb2_scene* scene, const pb2_sphere* spheres, uint32_t n);
/* Participating media.  media[n_media]: the scene's HomogeneousMedium list.  prim_inside / prim_outside: per primitive
 * (triangles, then spheres) the MediumInterface of its GeometricPrimitive (src/core/primitive.rs:33-38, src/core/medium.rs:
 * 90-115) as indices into media, -1 = no medium; either may be NULL (= all -1).  A primitive whose two sides are equal is not a
 * medium transition: a ray that hits it keeps its own medium on both sides (primitive.rs:72-76).  camera_medium: the medium
 * camera rays start in (src/cameras/perspective.rs:109), -1 = none.  Call before pb2_scene_build_bvh. */
int pb2_scene_set_media(pb2_scene* scene, const pb2_medium* media, uint32_t n_media, const int32_t* prim_inside,
                        const int32_t* prim_outside, int32_t camera_medium);
int pb2_scene_destroy(pb2_scene* scene);
/* Host SAH build (bvh.rs:273-473 recursive_build, :774-811 flatten_bvh_tree), repack to the 64-byte child-pair
 * node layout + 48-byte triangles, upload to the current device.  split_method (bvh.rs:199-204): 0 = SplitMethod::SAH,
 * built on the host; 1 = SplitMethod::HLBVH (bvh.rs:475-772), built on the GPU (Morton codes, radix sort, one LBVH treelet
 * per 12-bit Morton prefix, SAH over the treelet roots); 2 = SplitMethod::Middle, 3 = SplitMethod::EqualCounts
 * (bvh.rs:331-360), built on the host by the same recursive_build. */
int pb2_scene_build_bvh(pb2_scene* scene, int max_prims_in_node, int split_method);
/* Host half only (no device needed): the flattened array can then be inspected with pb2_bvh_info / pb2_bvh_export. */
int pb2_scene_build_bvh_host(pb2_scene* scene, int max_prims_in_node, int split_method);
/* bvh.rs:819-826 BVHAccel::world_bound -> {min.xyz, max.xyz} */
/* Stage times of the last HLBVH build in ms: upload + bounds + Morton codes, sort, treelets, upper SAH (host, <= 4096
 * treelet roots), flatten, repack into the traversal layout (device).  All zero after a SAH build. */
int pb2_bvh_build_stats(const pb2_scene* scene, double ms[6]);
int pb2_world_bound(const pb2_scene* scene, float out[6]);
/* Parity hooks: the flattened LinearBVHNode array (bvh.rs:129-135 as 32-byte nodes {bounds[6], offset u32,
 * n_prims u16, axis u8, pad}) and the ordered primitive list. */
int pb2_bvh_info(const pb2_scene* scene, uint64_t* n_nodes, uint64_t* n_prims, int* max_depth);
int pb2_bvh_export(const pb2_scene* scene, void* nodes32, uint32_t* ordered_prims);

/* ---- Primitive::intersect / intersect_p, batched (bvh.rs:828-879, :881-932) ---------------------------- */
/* Closest hit for n rays.  hits[i].t is the value Primitive::intersect would leave in ray.t_max.  b0 may be NULL. */
int pb2_intersect(pb2_scene* scene, const pb2_ray* rays, uint64_t n, pb2_hit* hits, float* b0);
/* Any hit: out[i] = 1 if Primitive::intersect_p would return true. */
int pb2_intersect_p(pb2_scene* scene, const pb2_ray* rays, uint64_t n, uint8_t* out);
/* Asynchronous forms ("async variants"): enqueue the batch on the scene's copy / compute ring and return.
 * The buffers must be pinned (pb2_host_alloc) and stay untouched until pb2_scene_wait(scene) — or any synchronous
 * pb2_intersect / pb2_intersect_p on the same scene — returns.  Batches enqueued back to back overlap: the drain of one
 * (its last kernels and device-to-host copies) runs under the host-to-device copies of the next, which is how a caller
 * with several independent batches (one per tile, per worker thread) keeps the link busy. */
int pb2_intersect_async(pb2_scene* scene, const pb2_ray* rays, uint64_t n, pb2_hit* hits, float* b0);
int pb2_intersect_p_async(pb2_scene* scene, const pb2_ray* rays, uint64_t n, uint8_t* out);
int pb2_scene_wait(pb2_scene* scene);
/* Returns when all but the `in_flight` most recently enqueued _async batches have their results in the caller's buffers
 * (in_flight = 0: pb2_scene_wait).  A caller that produces batches continuously keeps a few in flight — enqueue the next
 * ones, then wait for the older ones — so that the ring never drains between them. */
int pb2_scene_wait_until(pb2_scene* scene, uint32_t in_flight);
/* Device-resident variants (inputs/outputs already in HBM; asynchronous on `stream`). */
int pb2_intersect_device(pb2_scene* scene, const void* d_rays, uint64_t n, void* d_hits, void* d_b0, void* stream);
int pb2_intersect_p_device(pb2_scene* scene, const void* d_rays, uint64_t n, void* d_out, void* stream);

/* ---- Camera::generate_ray (src/cameras/perspective.rs:90-112), batched -------------------------------- */
/* One ray per CameraSample: p_film[i] = {x, y} in raster space and, for a thin lens, p_lens[i] = {u, v} in [0,1)^2
 * (NULL: the lens centre). */
int pb2_camera_generate_rays(const pb2_camera* cam, const float* p_film, const float* p_lens, uint64_t n, pb2_ray* rays);
/* One ray through every pixel centre (x+0.5, y+0.5), row-major; device output. */
int pb2_camera_primary_rays_device(const pb2_camera* cam, void* d_rays, void* stream);
/* Host-side matrices for parity checks: raster_to_camera, camera_to_world (row-major 4x4 each). */
int pb2_camera_matrices(const pb2_camera* cam, float r2c[16], float c2w[16]);

/* ---- secondary-ray builders used by the C3 workload (src/core/interaction.rs:132-153) ------------------ */
/* From each closest hit: Interaction::spawn_ray_to(point) shadow rays (t_max = 1 - SHADOW_EPSILON); misses get
 * a degenerate ray with t_max = -1 that hits nothing. */
int pb2_spawn_shadow_rays_device(pb2_scene* scene, const void* d_rays, const void* d_hits, uint64_t n,
                                 const float light_pos[3], void* d_out_rays, void* stream);
/* From each closest hit: Interaction::spawn_ray(wi), wi = cosine_sample_hemisphere about the geometric normal
 * (src/core/sampling.rs:289-294) drawn from PCG32 stream `i` (src/core/rng.rs). */
int pb2_spawn_bounce_rays_device(pb2_scene* scene, const void* d_rays, const void* d_hits, uint64_t n,
                                 void* d_out_rays, void* stream);
/* Both of the above from one pass over the hits (each hit's triangle is gathered and its interaction rebuilt once); the two
 * outputs equal those of the separate calls bit for bit. */
int pb2_spawn_shadow_bounce_rays_device(pb2_scene* scene, const void* d_rays, const void* d_hits, uint64_t n, const float light_pos[3],
                                        void* d_out_shadow_rays, void* d_out_bounce_rays, void* stream);

/* ---- RNG (src/core/rng.rs:14-48) — parity hook --------------------------------------------------------- */
/* out[s*n_per + k] = k-th uniform_float() of RNG::new(first_sequence + s). */
int pb2_rng_uniform_floats(uint64_t first_sequence, uint32_t n_sequences, uint32_t n_per, float* out);

/* ---- Film (src/core/film.rs) ---------------------------------------------------------------------------- */
int pb2_film_create(const pb2_film_desc* desc, pb2_film** out);     /* Film::new :31-75 */
int pb2_film_destroy(pb2_film* film);
int pb2_film_clear(pb2_film* film);
/* FilmTile::add_sample (:252-295) + Film::merge_film_tile (:111-123), batched: n samples, host arrays. */
int pb2_film_add_samples(pb2_film* film, const float* p_film, const float* L_rgb, const float* weight, uint64_t n);
/* Raw accumulators: float4 {X, Y, Z, filter_weight_sum} per pixel, row-major. */
int pb2_film_read_xyzw(pb2_film* film, float* out);
/* Film::write_image (:153-178): XYZ -> RGB, / weight, clamp >= 0, * scale. */
int pb2_film_resolve_rgb(pb2_film* film, float scale, float* rgb);
/* Film::add_splat (film.rs:137-151) for n splats: p_film[2n] film positions, v_rgb[3n] values; a splat lands on pixel
 * floor(p) when that lies inside cropped_pixel_bounds, after the max_sample_luminance clamp.  Film::set_image (film.rs:125-135):
 * rgb[3 * n_pixels] becomes the film (weight 1, splats cleared).  pb2_film_resolve_rgb_splat is Film::write_image's pixel loop
 * (film.rs:153-178) with its splat_scale argument: rgb = max(xyz_to_rgb(xyz) / weight, 0) + splat_scale * xyz_to_rgb(splat), times
 * scale; pb2_film_resolve_rgb and pb2_film_write_image use splat_scale = 1. */
int pb2_film_add_splats(pb2_film* film, const float* p_film, const float* v_rgb, uint64_t n);
int pb2_film_set_image(pb2_film* film, const float* rgb);
int pb2_film_resolve_rgb_splat(pb2_film* film, float scale, float splat_scale, float* rgb);
int pb2_film_device_ptr(pb2_film* film, void** d_xyzw, uint64_t* n_floats);
/* Film::cropped_pixel_bounds (film.rs:41-50) and Film::get_sample_bounds (:76-81): {x0, y0, x1, y1} each (x1, y1 exclusive). */
int pb2_film_bounds(const pb2_film* film, int32_t pixel_bounds[4], int32_t sample_bounds[4]);
/* Film::write_image (film.rs:153-180) through to a file — the reference stops at todo!() after building the RGB array
 * (imageio.rs:3-5).  ".pfm": float RGB, rows bottom-to-top, little-endian; ".ppm": 8-bit P6 with pbrt's sRGB gamma
 * (pbrt-v3 imageio.cpp GammaCorrect + 255 v + 0.5 clamp).  Any other extension is PB2_ERR_INVALID. */
int pb2_film_write_image(pb2_film* film, const char* filename, float scale);

/* ---- Integrator::render / PathIntegrator::li (src/core/integrator.rs:399-480, src/integrators/path.rs) -- */
/* Wavefront path tracer: renders sample indices [sample_begin, sample_end) of every pixel into `film`
 * (accumulating).  Sampler stream of (pixel x,y; sample s) = RNG::new((y*res_x + x)*spp + s). */
int pb2_render_path(pb2_scene* scene, const pb2_camera* cam, const pb2_path_desc* path, pb2_film* film, void* stream);
/* Per-sample radiance of PathIntegrator::li for explicit (pixel, sample) pairs — parity hook.  host arrays. */
int pb2_path_li(pb2_scene* scene, const pb2_camera* cam, const pb2_path_desc* path, const uint32_t* pixel_xy,
                const uint32_t* sample_index, uint64_t n, float* L_rgb, float* p_film);
/* SpatialLightDistribution probe (lightdistrib.rs:83-158): builds the voxel grid if needed and returns its extents; when
 * non-null, func receives every voxel's Distribution1D func ([voxel][n_lights], voxel = (z * ny + y) * nx + x), cdf its cdf
 * ([voxel][n_lights + 1]) and func_int its integral ([voxel]).  Host arrays. */
int pb2_spatial_light_distribution(pb2_scene* scene, int32_t n_voxels[3], float* func, float* cdf, float* func_int);
/* Counters of the last pb2_render_path call: {camera_samples, extend_rays, shadow_rays, mis_rays, kernel_launches}. */
int pb2_render_counters(pb2_scene* scene, uint64_t out[8]);

/* ---- multi-GPU film reduce (one process per GPU; NCCL over NVLink) -------------------------------------- */
int pb2_nccl_unique_id(char id[128]);
int pb2_nccl_init(const char id[128], int rank, int n_ranks);
int pb2_nccl_shutdown(void);
/* ncclReduce(sum, float32, 4*W*H, root) of the film accumulators. */
int pb2_film_reduce(pb2_film* film, int root, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PBRT_B200_H */
