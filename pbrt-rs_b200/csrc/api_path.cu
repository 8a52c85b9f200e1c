// api_path.cu — Film, wavefront PathIntegrator and NCCL entry points of include/pbrt_b200.h.
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>

#include "api_internal.hpp"

using pb2::UseChain;

#include <cstdio>
#include <string>
#include <vector>
#include "wavefront.cuh"

namespace pb2 {
size_t film_sort_scratch_bytes(uint32_t capacity);
}

struct pb2_film {
    pb2_film_desc desc;
    float table[256];
    float* d_table = nullptr;
    float4* d_xyzw = nullptr;
    float4* d_acc = nullptr;
    float4* d_splat = nullptr;                     // Pixel::splat_xyz, allocated by the first pb2_film_add_splats
    void* d_stray = nullptr;
    float4* d_stray_vals = nullptr;
    uint32_t stray_capacity = 0;
    unsigned long long* d_counters = nullptr;      // C_COUNT slots, used when the film is filled without a wavefront
    int px0, py0, px1, py1;                        // cropped_pixel_bounds (film.rs:41-50)
    int sb_x0, sb_y0, sb_x1, sb_y1;
    int device = 0;
    UseChain chain;                                // orders every use of the accumulators across streams (api_internal.hpp)
    size_t n_pixels() const { return (size_t)(px1 - px0) * (size_t)(py1 - py0); }
};

namespace pb2 {

static FilmView film_view(const pb2_film* f) {
    FilmView v;
    v.px0 = f->px0; v.py0 = f->py0; v.px1 = f->px1; v.py1 = f->py1;
    v.max_lum = f->desc.max_sample_luminance > 0.0f ? f->desc.max_sample_luminance : __builtin_huge_valf();
    v.sb_x0 = f->sb_x0; v.sb_y0 = f->sb_y0; v.sb_w = f->sb_x1 - f->sb_x0; v.sb_h = f->sb_y1 - f->sb_y0;
    v.sb_w_magic = fast_div_magic((uint32_t)v.sb_w);
    v.radius_x = f->desc.radius_x; v.radius_y = f->desc.radius_y;
    v.exact = (f->desc.filter == PB2_FILTER_BOX && f->desc.radius_x == 0.5f && f->desc.radius_y == 0.5f) ? 1 : 0;
    v.table = f->d_table;
    v.acc = f->d_acc;
    v.xyzw = f->d_xyzw;
    v.splat = f->d_splat;
    v.stray_keys = (unsigned long long*)f->d_stray;
    v.stray_vals = f->d_stray_vals;
    v.stray_capacity = f->stray_capacity;
    return v;
}

// TrowbridgeReitzDistribution::roughness_to_alpha (microfacet.rs:160-168), evaluated once per material on the host
static float roughness_to_alpha(float roughness) {
    roughness = std::fmax(roughness, 1e-3f);
    const float x = std::log(roughness);
    return (((1.62142f + 0.819955f * x) + 0.1734f * x * x) + 0.0171201f * x * x * x) + 0.000640711f * x * x * x * x;
}

// Distribution1D::new (sampling.rs:76-98, D29 FIX)
static void make_distribution(const std::vector<float>& func, std::vector<float>* cdf, float* func_int) {
    const size_t n = func.size();
    cdf->assign(n + 1, 0.0f);
    for (size_t i = 1; i < n + 1; ++i) (*cdf)[i] = (*cdf)[i - 1] + func[i - 1] / (float)n;
    *func_int = (*cdf)[n];
    if (*func_int == 0.0f) for (size_t i = 1; i < n + 1; ++i) (*cdf)[i] = (float)i / (float)n;
    else for (size_t i = 1; i < n + 1; ++i) (*cdf)[i] /= *func_int;
}

// Device tables for shading: per-primitive material / area-light ids, materials, lights and both light distributions
// (lightdistrib.rs:222-232: "uniform" and "power"; integrator.rs:268-277).
int upload_shading_tables(pb2_scene* scene) {
    const size_t n_mesh = scene->indices.size() / 3;
    const size_t n_tris = scene->n_primitives();                         // triangles, then analytic spheres
    if ((n_mesh && scene->tri_material.empty()) || scene->materials.empty() || n_tris == 0) return PB2_OK;       // ray-casting-only scene
    std::vector<uint32_t> prim_material(scene->tri_material);
    for (const pb2_sphere& sp : scene->spheres) {
        if (sp.material >= scene->materials.size() && sp.material != PB2_NO_MATERIAL) return set_error(PB2_ERR_INVALID, "a sphere references material %u >= %zu", sp.material, scene->materials.size());
        prim_material.push_back(sp.material);
    }
    std::vector<DMaterial> mats(scene->materials.size());
    unsigned class_mask = 0u;
    bool class1_all_plastic = true;
    for (size_t i = 0; i < mats.size(); ++i) {
        const pb2_material& m = scene->materials[i];
        if (m.type < PB2_MAT_MATTE || m.type > PB2_MAT_SUBSTRATE) return set_error(PB2_ERR_INVALID, "material %zu has unknown type %d", i, m.type);
        DMaterial& d = mats[i];
        d.type = m.type;
        for (int k = 0; k < 3; ++k) { d.kd[k] = m.kd[k]; d.ks[k] = m.ks[k]; d.kr[k] = m.kr[k]; d.kt[k] = m.kt[k]; d.metal_eta[k] = m.metal_eta[k]; d.metal_k[k] = m.metal_k[k]; }
        d.alpha = m.remap_roughness ? roughness_to_alpha(m.roughness) : m.roughness;
        d.eta = m.eta;
        {
            // OrenNayar::new (reflection.rs:925-937; sigma in radians: D61 FIX); MatteMaterial clamps sigma to [0, 90] degrees
            const float sd = m.sigma < 0.0f ? 0.0f : (m.sigma > 90.0f ? 90.0f : m.sigma);
            const float sig = PB2_PI / 180.0f * sd, sigma2 = sig * sig;
            d.on_a = 1.0f - (sigma2 / (2.0f * (sigma2 + 0.33f)));
            d.on_b = 0.45f * sigma2 / (sigma2 + 0.09f);
        }
        // shading class (shade.cuh: make_bsdf<CLS>)
        d.cls = (m.type == PB2_MAT_MATTE && m.sigma == 0.0f) ? 0 : (((m.type == PB2_MAT_GLASS && m.roughness == 0.0f) || m.type == PB2_MAT_MIRROR) ? 2 : 1);
        class_mask |= 1u << d.cls;
        if (d.cls == 1) {
            const bool two_lobe_plastic = m.type == PB2_MAT_PLASTIC && (m.kd[0] != 0.0f || m.kd[1] != 0.0f || m.kd[2] != 0.0f) &&
                                          (m.ks[0] != 0.0f || m.ks[1] != 0.0f || m.ks[2] != 0.0f);
            class1_all_plastic = class1_all_plastic && two_lobe_plastic;
        }
    }
    // bit 3: every class-1 material is a PlasticMaterial with both lobes -> k_shade<3> (the plastic-only kernel) shades queue 1
    if ((class_mask & 2u) && class1_all_plastic && !(getenv("PB2_GENERAL_CLASS1") && atoi(getenv("PB2_GENERAL_CLASS1")) != 0)) class_mask |= 8u;
    scene->shading_class_mask = class_mask;
    scene->has_material_less = false;
    for (uint32_t m : prim_material) scene->has_material_less |= m == PB2_NO_MATERIAL;
    if (!scene->media.empty()) {                                         // HomogeneousMedium::new: sigma_t = sigma_s + sigma_a (homogeneous.rs:26)
        if (scene->prim_inside.size() != n_tris) return set_error(PB2_ERR_STATE, "pb2_scene_set_media was called before the last primitives were added");
        std::vector<DMedium> dm(scene->media.size());
        for (size_t i = 0; i < dm.size(); ++i)
            for (int k = 0; k < 3; ++k) {
                dm[i].sigma_a[k] = scene->media[i].sigma_a[k];
                dm[i].sigma_s[k] = scene->media[i].sigma_s[k];
                dm[i].sigma_t[k] = scene->media[i].sigma_s[k] + scene->media[i].sigma_a[k];
                dm[i].g = scene->media[i].g;
            }
        PB2_CUDA(cudaMalloc(&scene->d_media, dm.size() * sizeof(DMedium)));
        PB2_CUDA(cudaMalloc(&scene->d_prim_inside, n_tris * 4));
        PB2_CUDA(cudaMalloc(&scene->d_prim_outside, n_tris * 4));
        PB2_CUDA(cudaMemcpy(scene->d_media, dm.data(), dm.size() * sizeof(DMedium), cudaMemcpyHostToDevice));
        PB2_CUDA(cudaMemcpy(scene->d_prim_inside, scene->prim_inside.data(), n_tris * 4, cudaMemcpyHostToDevice));
        PB2_CUDA(cudaMemcpy(scene->d_prim_outside, scene->prim_outside.data(), n_tris * 4, cudaMemcpyHostToDevice));
    }
    const size_t n_lights = scene->lights.size();
    std::vector<DLight> lights(std::max<size_t>(1, n_lights));
    std::vector<int32_t> tri_light(n_tris, -1);
    std::vector<float> power(n_lights), ones(n_lights, 1.0f);
    for (size_t i = 0; i < n_lights; ++i) {
        const pb2_light& l = scene->lights[i];
        DLight& d = lights[i];
        memset(&d, 0, sizeof d);
        d.type = l.type;
        for (int k = 0; k < 3; ++k) { d.p[k] = l.p[k]; d.l[k] = l.i[k]; }
        d.prim = l.prim_id;
        d.two_sided = l.two_sided;
        d.sphere = -1;
        rgb3 pw;
        if (l.type == PB2_LIGHT_AREA && l.prim_id >= n_mesh) {           // DiffuseAreaLight over a Sphere (sphere.rs:100-102)
            d.sphere = (int)(l.prim_id - n_mesh);
            d.area = sphere_area(scene->sphere_records[d.sphere]);
            tri_light[l.prim_id] = (int32_t)i;
            pw = mkc(l.i[0], l.i[1], l.i[2]) * ((l.two_sided ? 2.0f : 1.0f) * d.area * PB2_PI);      // diffuse.rs:83-85
        } else if (l.type == PB2_LIGHT_AREA) {
            const uint32_t* ix = &scene->indices[3ull * l.prim_id];
            const float* v = scene->verts.data();
            const vec3 p0 = mk(v[3 * ix[0]], v[3 * ix[0] + 1], v[3 * ix[0] + 2]);
            const vec3 p1 = mk(v[3 * ix[1]], v[3 * ix[1] + 1], v[3 * ix[1] + 2]);
            const vec3 p2 = mk(v[3 * ix[2]], v[3 * ix[2] + 1], v[3 * ix[2] + 2]);
            d.p0[0] = p0.x; d.p0[1] = p0.y; d.p0[2] = p0.z;
            d.p1[0] = p1.x; d.p1[1] = p1.y; d.p1[2] = p1.z;
            d.p2[0] = p2.x; d.p2[1] = p2.y; d.p2[2] = p2.z;
            d.area = len(cross3(p1 - p0, p2 - p0)) * 0.5f;               // triangle.rs:323-328
            d.has_n = scene->normals.empty() ? 0 : 1;
            d.has_uv = scene->uvs.empty() ? 0 : 1;
            for (int c = 0; c < 3; ++c) {
                if (d.has_n) { d.n0[c] = scene->normals[3 * ix[0] + c]; d.n1[c] = scene->normals[3 * ix[1] + c]; d.n2[c] = scene->normals[3 * ix[2] + c]; }
                if (d.has_uv && c < 2) { d.uv[c] = scene->uvs[2 * ix[0] + c]; d.uv[2 + c] = scene->uvs[2 * ix[1] + c]; d.uv[4 + c] = scene->uvs[2 * ix[2] + c]; }
            }
            {
                const vec3 ns = unit(cross3(p1 - p0, p2 - p0)), nh = unit(cross3(p0 - p2, p1 - p2));
                d.ns_sample[0] = ns.x; d.ns_sample[1] = ns.y; d.ns_sample[2] = ns.z;
                d.n_hit[0] = nh.x; d.n_hit[1] = nh.y; d.n_hit[2] = nh.z;
                d.inv_area = 1.0f / d.area;
                vec3 du, dv;
                d.frame_ok = (d.has_uv ? tri_frame_uv(p0, p1, p2, make_float2(d.uv[0], d.uv[1]), make_float2(d.uv[2], d.uv[3]), make_float2(d.uv[4], d.uv[5]), &du, &dv)
                                       : tri_frame(p0, p1, p2, &du, &dv)) ? 1 : 0;
            }
            tri_light[l.prim_id] = (int32_t)i;
            pw = mkc(l.i[0], l.i[1], l.i[2]) * ((l.two_sided ? 2.0f : 1.0f) * d.area * PB2_PI);      // diffuse.rs:83-85
        } else if (l.type == PB2_LIGHT_SPOT) {                           // spot.rs:30-49
            for (int k = 0; k < 3; ++k) d.axis[k] = l.axis[k];
            d.cos_total_width = std::cos(PB2_PI / 180.0f * l.total_width);       // radians(), pbrt.rs:133-135
            d.cos_falloff_start = std::cos(PB2_PI / 180.0f * l.falloff_start);
            pw = mkc(l.i[0], l.i[1], l.i[2]) * (2.0f * PB2_PI * (1.0f - 0.5f * (d.cos_falloff_start + d.cos_total_width)));      // spot.rs:87-89
        } else if (l.type == PB2_LIGHT_DISTANT) {                        // distant.rs:29-45
            const vec3 w = unit(mk(l.axis[0], l.axis[1], l.axis[2]));
            d.axis[0] = w.x; d.axis[1] = w.y; d.axis[2] = w.z;
            // Light::pre_process (distant.rs:73-77) + Bounds3::bounding_sphere (geometry.rs:473-480) of the scene's world bound
            const float* rb = scene->root_bounds;
            const vec3 lo = mk(rb[0], rb[1], rb[2]), hi = mk(rb[3], rb[4], rb[5]);
            const vec3 c = (lo + hi) / 2.0f;
            const bool inside = c.x >= lo.x && c.x <= hi.x && c.y >= lo.y && c.y <= hi.y && c.z >= lo.z && c.z <= hi.z;
            d.world_radius = inside ? len(c - hi) : 0.0f;
            pw = mkc(l.i[0], l.i[1], l.i[2]) * (PB2_PI * d.world_radius * d.world_radius);       // distant.rs:69-71
        } else {
            pw = mkc(l.i[0], l.i[1], l.i[2]) * (4.0f * PB2_PI);          // point.rs:68-70
        }
        power[i] = luminance(pw);
    }
    scene->light_func[0] = ones;
    scene->light_func[1] = (n_lights == 1) ? ones : power;               // lightdistrib.rs:223: one light -> uniform
    for (int k = 0; k < 2; ++k) make_distribution(scene->light_func[k], &scene->light_cdf[k], &scene->light_func_int[k]);
    PB2_CUDA(cudaMalloc(&scene->d_tri_material, n_tris * 4));
    PB2_CUDA(cudaMalloc(&scene->d_tri_light, n_tris * 4));
    PB2_CUDA(cudaMalloc(&scene->d_materials, mats.size() * sizeof(DMaterial)));
    PB2_CUDA(cudaMalloc(&scene->d_lights, lights.size() * sizeof(DLight)));
    PB2_CUDA(cudaMalloc(&scene->d_light_cdf, (4 * n_lights + 4) * sizeof(float)));
    PB2_CUDA(cudaMemcpy(scene->d_tri_material, prim_material.data(), n_tris * 4, cudaMemcpyHostToDevice));
    PB2_CUDA(cudaMemcpy(scene->d_tri_light, tri_light.data(), n_tris * 4, cudaMemcpyHostToDevice));
    PB2_CUDA(cudaMemcpy(scene->d_materials, mats.data(), mats.size() * sizeof(DMaterial), cudaMemcpyHostToDevice));
    PB2_CUDA(cudaMemcpy(scene->d_lights, lights.data(), lights.size() * sizeof(DLight), cudaMemcpyHostToDevice));
    // layout: [func uniform | cdf uniform | func power | cdf power]
    float* base = (float*)scene->d_light_cdf;
    size_t off = 0;
    for (int k = 0; k < 2; ++k) {
        if (n_lights) PB2_CUDA(cudaMemcpy(base + off, scene->light_func[k].data(), n_lights * 4, cudaMemcpyHostToDevice));
        off += n_lights;
        PB2_CUDA(cudaMemcpy(base + off, scene->light_cdf[k].data(), (n_lights + 1) * 4, cudaMemcpyHostToDevice));
        off += n_lights + 1;
    }
    return PB2_OK;
}

// SpatialLightDistribution::new + every voxel's compute_distribution (lightdistrib.rs:83-158); scene->mu is held.
static constexpr uint64_t kSpatialMaxBytes = 16ull << 30;
static SpatialView spatial_view(const pb2_scene* s) {
    SpatialView g;
    const size_t n = s->lights.size();
    const size_t n_vox = (size_t)s->spatial_nv[0] * s->spatial_nv[1] * s->spatial_nv[2];
    const float* base = (const float*)s->d_spatial;
    g.func = base;
    g.cdf = base + n_vox * n;
    g.func_int = base + n_vox * n + n_vox * (n + 1);
    for (int i = 0; i < 3; ++i) { g.nv[i] = s->spatial_nv[i]; g.lo[i] = s->root_bounds[i]; g.hi[i] = s->root_bounds[3 + i]; }
    return g;
}
static int ensure_spatial(pb2_scene* scene, cudaStream_t st) {
    if (scene->d_spatial) return PB2_OK;
    const size_t n = scene->lights.size();
    if (n == 0) return set_error(PB2_ERR_STATE, "the scene has no lights");
    int nv[3];
    spatial_grid_extents(scene->root_bounds, 64, nv);
    const uint64_t n_vox = (uint64_t)nv[0] * nv[1] * nv[2];
    const uint64_t bytes = (n_vox * n + n_vox * (n + 1) + n_vox) * sizeof(float);
    // The reference fills a voxel at its first lookup (lightdistrib.rs:165-220); here the whole grid is filled eagerly, every
    // emissive triangle being one light, so the tables are bounded by what the device can hold beside the wavefront arenas:
    // half of the free memory, at most kSpatialMaxBytes.
    size_t free_b = 0, total_b = 0;
    PB2_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const uint64_t limit = std::min<uint64_t>(kSpatialMaxBytes, (uint64_t)free_b / 2);
    if (bytes > limit)
        return set_error(PB2_ERR_LIMIT, "spatial light distribution: %llu voxels x %zu lights need %.1f GB of tables (limit %.1f GB: half of the free "
                         "device memory, at most %.0f GB); use \"power\"", (unsigned long long)n_vox, n, bytes / 1e9, limit / 1e9, kSpatialMaxBytes / 1e9);
    PB2_CUDA(cudaMalloc(&scene->d_spatial, bytes));
    for (int i = 0; i < 3; ++i) scene->spatial_nv[i] = nv[i];
    const SpatialView g = spatial_view(scene);
    spatial_distribution_build(g, (const DLight*)scene->d_lights, (int)n, (float*)g.func, (float*)g.cdf, (float*)g.func_int, st, scene->d_spheres);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        cudaFree(scene->d_spatial);
        scene->d_spatial = nullptr;
        return cuda_fail(e, "spatial_distribution_build", __FILE__, __LINE__);
    }
    return PB2_OK;
}

static ShadeView shade_view(const pb2_scene* s, int strategy) {
    ShadeView v;
    const size_t n = s->lights.size();
    v.tri_material = (const uint32_t*)s->d_tri_material;
    v.tri_light = (const int32_t*)s->d_tri_light;
    v.mats = (const DMaterial*)s->d_materials;
    v.lights = (const DLight*)s->d_lights;
    v.n_lights = (int)n;
    v.class_mask = s->shading_class_mask;
    const float* base = (const float*)s->d_light_cdf;
    const size_t off = strategy == PB2_LIGHTS_POWER ? (2 * n + 1) : 0;
    v.light_func = base + off;
    v.light_cdf = base + off + n;
    v.light_func_int = s->light_func_int[strategy == PB2_LIGHTS_POWER ? 1 : 0];
    memset(&v.spatial, 0, sizeof v.spatial);
    if (strategy == PB2_LIGHTS_SPATIAL && n > 1 && s->d_spatial) v.spatial = spatial_view(s);      // lightdistrib.rs:223: one light -> uniform
    v.indices = (const uint32_t*)s->d_indices;
    v.normals = (const float*)s->d_normals;
    v.tangents = (const float*)s->d_tangents;
    v.uvs = (const float2*)s->d_uvs;
    v.media = (const DMedium*)s->d_media;
    v.prim_inside = (const int32_t*)s->d_prim_inside;
    v.prim_outside = (const int32_t*)s->d_prim_outside;
    v.camera_medium = s->d_media ? s->camera_medium : -1;
    v.has_interfaces = s->has_material_less ? 1 : 0;
    return v;
}

static int check_path_args(pb2_scene* scene, const pb2_camera* cam, const pb2_path_desc* path, const pb2_film_desc* fd, CameraView* cv) {
    if (!scene || !cam || !path) return set_error(PB2_ERR_INVALID, "null argument");
    if (!scene->built) return set_error(PB2_ERR_STATE, "pb2_scene_build_bvh has not been called");
    if (!scene->d_tri_material) return set_error(PB2_ERR_STATE, "the scene was created without materials");
    if (path->max_depth < 0 || path->max_depth > 65535) return set_error(PB2_ERR_INVALID, "max_depth out of range");
    if (path->spp <= 0 || path->sample_begin < 0 || path->sample_end > path->spp || path->sample_begin > path->sample_end)
        return set_error(PB2_ERR_INVALID, "bad sample range [%d,%d) of %d", path->sample_begin, path->sample_end, path->spp);
    if (path->light_strategy < PB2_LIGHTS_UNIFORM || path->light_strategy > PB2_LIGHTS_SPATIAL)
        return set_error(PB2_ERR_INVALID, "unknown light strategy %d", path->light_strategy);
    if (path->integrator != PB2_INTEGRATOR_PATH && path->integrator != PB2_INTEGRATOR_VOLPATH)
        return set_error(PB2_ERR_INVALID, "unknown integrator %d", path->integrator);
    if (path->integrator == PB2_INTEGRATOR_PATH && scene->has_material_less)
        return set_error(PB2_ERR_INVALID, "the scene has surfaces without a material (medium interfaces): render it with PB2_INTEGRATOR_VOLPATH");
    if (path->integrator == PB2_INTEGRATOR_VOLPATH && (path->sampler == PB2_SAMPLER_HALTON || path->sampler == PB2_SAMPLER_SOBOL))
        return set_error(PB2_ERR_INVALID, "VolPathIntegrator draws an unbounded number of sampler dimensions: use the random, stratified or (0,2) sampler");
    if (path->sampler < PB2_SAMPLER_RANDOM || path->sampler > PB2_SAMPLER_SOBOL)
        return set_error(PB2_ERR_INVALID, "unknown sampler %d", path->sampler);
    if (path->sampler == PB2_SAMPLER_SOBOL) {
        if ((path->spp & (path->spp - 1)) != 0)
            return set_error(PB2_ERR_INVALID, "SobolSampler: spp %d is not a power of two (sobol.rs:22-28 rounds up; pass the rounded count)", path->spp);
        if (5 + 8 * (path->max_depth + 1) > 1024)                                              // sobolmatrices.rs:1 NUM_SOBOL_DIMENSIONS
            return set_error(PB2_ERR_LIMIT, "SobolSampler has 1024 dimensions; max_depth %d needs %d", path->max_depth, 5 + 8 * (path->max_depth + 1));
    }
    if (path->sampler == PB2_SAMPLER_STRATIFIED || path->sampler == PB2_SAMPLER_ZEROTWO) {
        if (path->n_sampled_dimensions < 0 || path->n_sampled_dimensions > 127)
            return set_error(PB2_ERR_INVALID, "n_sampled_dimensions %d outside [0, 127]", path->n_sampled_dimensions);
        if (path->sampler == PB2_SAMPLER_STRATIFIED && (path->x_samples <= 0 || path->y_samples <= 0 || (long long)path->x_samples * path->y_samples != path->spp))
            return set_error(PB2_ERR_INVALID, "StratifiedSampler: spp %d != x_samples %d * y_samples %d (stratified.rs:31-32)", path->spp, path->x_samples, path->y_samples);
        if (path->sampler == PB2_SAMPLER_ZEROTWO && (path->spp & (path->spp - 1)) != 0)
            return set_error(PB2_ERR_INVALID, "ZeroTwoSequenceSampler: spp %d is not a power of two (zerotwosequence.rs:21 rounds up; pass the rounded count)", path->spp);
    }
    if (path->sampler == PB2_SAMPLER_HALTON && 5 + 8 * (path->max_depth + 1) > 1000)      // lowdiscrepancy.rs:11 PRIME_TABLE_SIZE
        return set_error(PB2_ERR_LIMIT, "HaltonSampler has 1000 dimensions; max_depth %d needs %d", path->max_depth, 5 + 8 * (path->max_depth + 1));
    if (fd && (cam->res_x != fd->res_x || cam->res_y != fd->res_y)) return set_error(PB2_ERR_INVALID, "camera and film resolutions differ");
    return make_camera_view(cam, cv);
}

// ---- HaltonSampler set-up (samplers/halton.rs:64-103, lowdiscrepancy.rs:333-349; host, once per scene) ------------------------
static void halton_gcd(uint64_t a, uint64_t b, int64_t* x, int64_t* y) {                     // halton.rs:52-62
    if (b == 0) { *x = 1; *y = 0; return; }
    int64_t xp, yp;
    halton_gcd(b, a % b, &xp, &yp);
    *x = yp;
    *y = xp - (int64_t)(a / b) * yp;
}
static uint64_t halton_mult_inverse(int64_t a, int64_t n) {                                  // halton.rs:41-50
    int64_t x, y;
    halton_gcd((uint64_t)a, (uint64_t)n, &x, &y);
    const int64_t r = x - (x / n) * n;
    return (uint64_t)(r < 0 ? r + n : r);
}
// PixelSampler tables (stratified / (0,2)): generated on the device once per (sampler parameters, sample-bounds extent) and
// kept with the scene; a frame rendered in several pb2_render_path calls (sample ranges) reuses them.
static int pixel_sampler_tables(pb2_scene* scene, const pb2_path_desc* path, int sb_w, int sb_h, cudaStream_t st, SamplerView* out) {
    const uint64_t n_pix = (uint64_t)sb_w * (uint64_t)sb_h;
    const uint64_t entries = n_pix * (uint64_t)path->spp * (uint64_t)path->n_sampled_dimensions;
    if (entries >= (1ull << 32))
        return set_error(PB2_ERR_LIMIT, "PixelSampler tables: %llu pixels x %d spp x %d dimensions exceed 2^32 entries (12 bytes each); render in tiles of fewer pixels",
                         (unsigned long long)n_pix, path->spp, path->n_sampled_dimensions);
    const long long key[8] = {path->sampler, path->spp, path->n_sampled_dimensions, path->sampler == PB2_SAMPLER_STRATIFIED ? path->x_samples : 0,
                              path->sampler == PB2_SAMPLER_STRATIFIED ? path->y_samples : 0,
                              path->sampler == PB2_SAMPLER_STRATIFIED ? (path->jitter != 0) : 0, sb_w, sb_h};
    if (!scene->d_tab1 || memcmp(key, scene->tab_key, sizeof key) != 0) {
        if (scene->d_tab1) { cudaFree(scene->d_tab1); scene->d_tab1 = nullptr; }
        if (scene->d_tab2) { cudaFree(scene->d_tab2); scene->d_tab2 = nullptr; }
        cudaError_t e = cudaMalloc(&scene->d_tab1, std::max<uint64_t>(entries, 1) * 4);
        if (e == cudaSuccess) e = cudaMalloc(&scene->d_tab2, std::max<uint64_t>(entries, 1) * 8);
        if (e != cudaSuccess) {
            cudaFree(scene->d_tab1); scene->d_tab1 = nullptr; scene->d_tab2 = nullptr;
            (void)cudaGetLastError();
            return set_error(PB2_ERR_LIMIT, "PixelSampler tables need %.1f GB of device memory: %s", (double)entries * 12e-9, cudaGetErrorString(e));
        }
        // table stream of pixel p: RNG::new(n_pix * spp + p), disjoint from the per-(pixel, sample) streams [0, n_pix * spp)
        pixel_tables_generate(path->sampler, (uint32_t)n_pix, path->spp, path->n_sampled_dimensions, path->x_samples, path->y_samples, path->jitter != 0,
                              n_pix * (uint64_t)path->spp, (float*)scene->d_tab1, (float2*)scene->d_tab2, st);
        PB2_CUDA(cudaGetLastError());
        memcpy(scene->tab_key, key, sizeof key);
    }
    out->n_dims = path->n_sampled_dimensions;
    out->spp_tab = path->spp;
    out->tab_n_pix = (uint32_t)n_pix;
    out->t1 = (const float*)scene->d_tab1;
    out->t2 = (const float2*)scene->d_tab2;
    return PB2_OK;
}

// The Sobol' generator matrices (core/sobolmatrices.rs: constant data, converted by tools/make_sobol_tables.py into
// data/sobol_tables.bin and linked in by the Makefile: ld -r -b binary).
extern "C" const unsigned char _binary_data_sobol_tables_bin_start[], _binary_data_sobol_tables_bin_end[];
static int sobol_view(pb2_scene* scene, int sb_x0, int sb_y0, int sb_w, int sb_h, SamplerView* out) {
    const unsigned char* blob = _binary_data_sobol_tables_bin_start;
    const size_t bytes = (size_t)(_binary_data_sobol_tables_bin_end - _binary_data_sobol_tables_bin_start);
    uint32_t h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (bytes >= 32) memcpy(h, blob, 32);
    const size_t n32 = (size_t)h[1] * h[2], nv = (size_t)h[3] * h[2], ni = (size_t)h[4] * h[2];
    if (h[0] != 0x31424F53u || h[2] != 52u || h[1] < 1024u || bytes < 32 + n32 * 4 + (nv + ni) * 8)
        return set_error(PB2_ERR_STATE, "the embedded Sobol' table is malformed (rebuild with tools/make_sobol_tables.py)");
    if (!scene->d_sobol) {
        PB2_CUDA(cudaMalloc(&scene->d_sobol, bytes - 32));
        PB2_CUDA(cudaMemcpy(scene->d_sobol, blob + 32, bytes - 32, cudaMemcpyHostToDevice));
    }
    out->sobol_m32 = (const uint32_t*)scene->d_sobol;
    out->sobol_vdc = (const unsigned long long*)((const char*)scene->d_sobol + n32 * 4);
    out->sobol_vdc_inv = out->sobol_vdc + nv;
    out->sobol_min[0] = sb_x0;
    out->sobol_min[1] = sb_y0;
    int res = 1, lg = 0;                                                                     // round_up_pow2_i32 / log_2_int_i32, sobol.rs:29-30
    while (res < std::max(sb_w, sb_h)) { res <<= 1; ++lg; }
    if ((uint32_t)lg > h[3]) return set_error(PB2_ERR_LIMIT, "SobolSampler: sample bounds of %d pixels exceed the 2^%u the van der Corput matrices cover", std::max(sb_w, sb_h), h[3]);
    out->sobol_resolution = res;
    out->sobol_log2_resolution = lg;
    return PB2_OK;
}

static int sampler_view(pb2_scene* scene, const pb2_path_desc* path, int sb_x0, int sb_y0, int sb_w, int sb_h, cudaStream_t st, SamplerView* out) {
    memset(out, 0, sizeof *out);
    const int sampler = path->sampler;
    out->kind = sampler;
    if (sampler == PB2_SAMPLER_SOBOL) return sobol_view(scene, sb_x0, sb_y0, sb_w, sb_h, out);
    if (sampler == PB2_SAMPLER_STRATIFIED || sampler == PB2_SAMPLER_ZEROTWO) return pixel_sampler_tables(scene, path, sb_w, sb_h, st, out);
    if (sampler != PB2_SAMPLER_HALTON) return PB2_OK;
    if (!scene->d_halton_perms) {
        constexpr int kPrimes = 1000;                                                        // lowdiscrepancy.rs:11
        std::vector<uint32_t> primes, sums(kPrimes, 0);
        for (uint32_t c = 2; (int)primes.size() < kPrimes; ++c) {
            bool is_prime = true;
            for (uint32_t p : primes) { if (p * p > c) break; if (c % p == 0) { is_prime = false; break; } }
            if (is_prime) primes.push_back(c);
        }
        for (int i = 1; i < kPrimes; ++i) sums[i] = sums[i - 1] + primes[i - 1];
        std::vector<uint16_t> perms(sums.back() + primes.back());
        Pcg32 rng;                                                                           // RNG::default(), rng.rs:14-19
        rng.state = 0x853c49e6748fea9bULL;
        rng.inc = 0xda3e39cb94b95bdbULL;
        size_t off = 0;
        for (int i = 0; i < kPrimes; ++i) {
            const uint32_t n = primes[i];
            for (uint32_t j = 0; j < n; ++j) perms[off + j] = (uint16_t)j;
            for (uint32_t j = 0; j < n; ++j) {                                               // shuffle, sampling.rs:280-287
                const uint32_t b = n - j, threshold = (~b + 1u) % b;                         // uniform_u32_u32, rng.rs:36-44
                uint32_t r;
                do { r = rng.next_u32(); } while (r < threshold);
                std::swap(perms[off + j], perms[off + j + r % b]);
            }
            off += n;
        }
        PB2_CUDA(cudaMalloc(&scene->d_halton_perms, perms.size() * 2));
        PB2_CUDA(cudaMalloc(&scene->d_halton_primes, kPrimes * 4));
        PB2_CUDA(cudaMalloc(&scene->d_halton_sums, kPrimes * 4));
        PB2_CUDA(cudaMemcpy(scene->d_halton_perms, perms.data(), perms.size() * 2, cudaMemcpyHostToDevice));
        PB2_CUDA(cudaMemcpy(scene->d_halton_primes, primes.data(), kPrimes * 4, cudaMemcpyHostToDevice));
        PB2_CUDA(cudaMemcpy(scene->d_halton_sums, sums.data(), kPrimes * 4, cudaMemcpyHostToDevice));
    }
    out->perms = (const uint16_t*)scene->d_halton_perms;
    out->primes = (const uint32_t*)scene->d_halton_primes;
    out->prime_sums = (const uint32_t*)scene->d_halton_sums;
    const int res[2] = {sb_w, sb_h};
    for (int i = 0; i < 2; ++i) {
        const int base = i == 0 ? 2 : 3;
        int scale = 1, exp = 0;
        while (scale < std::min(128, res[i])) { scale *= base; ++exp; }                      // MAX_RESOLUTION, halton.rs:39,72-79
        out->base_scales[i] = scale;
        out->base_exponents[i] = exp;
    }
    out->sample_stride = (unsigned long long)out->base_scales[0] * (unsigned long long)out->base_scales[1];
    out->mult_inverse[0] = halton_mult_inverse(out->base_scales[1], out->base_scales[0]);
    out->mult_inverse[1] = halton_mult_inverse(out->base_scales[0], out->base_scales[1]);
    return PB2_OK;
}

// Path slots of the wavefront (233 B each).  A frame is rendered in batches of floor(slots / pixels) samples per pixel and every
// batch costs ~7 launches per bounce with their ramp-up and tail, so a batch should hold several samples of every pixel: C4
// (2.07 M pixels, 32 spp) takes 173.9 / 152.0 / 143.8 / 141.3 ms with 2^22 / 2^23 / 2^24 / 2^25 slots, C2 (262 K pixels, 64 spp)
// 32.2 / 31.8 / 33.4 / 33.3 ms (gpurun_out/tune_slots.log).  With two batches in flight (Wavefront::peer) C4 @ 32 spp takes 104.6 /
// 96.5 / 93.1 / 91.3 ms with 2^22..2^25 slots per arena and C2 @ 64 spp 17.7 / 17.5 / 18.0 / 18.1 ms (profiles/r02_slots.log).
// Default: 16 samples of every pixel, at least 2^23 and at most 2^26 slots (15.6 GB; a frame of several batches holds two such
// arenas); PB2_WAVEFRONT_LOG2_SLOTS overrides it for sweeps.
static uint64_t wavefront_slots(uint64_t n_pix) {
    static const int forced = [] { const char* e = getenv("PB2_WAVEFRONT_LOG2_SLOTS"); return e ? std::min(std::max(atoi(e), 16), 28) : 0; }();
    if (forced) return 1ull << forced;
    return std::min<uint64_t>(std::max<uint64_t>(16 * n_pix, 1ull << 23), 1ull << 26);
}
static int ensure_wavefront(pb2_scene* scene, uint64_t min_capacity) {
    const uint64_t want = std::max<uint64_t>(min_capacity, wavefront_slots(min_capacity));
    if (scene->wf && scene->wf->capacity >= want) return PB2_OK;
    if (scene->wf) { wavefront_destroy(scene->wf); scene->wf = nullptr; }
    int rc = wavefront_create(want, &scene->wf);
    if (rc != 0) return set_error(PB2_ERR_CUDA, "wavefront buffers for %llu paths: %s", (unsigned long long)want, cudaGetErrorString((cudaError_t)rc));
    return PB2_OK;
}

// NCCL is bound at first use with dlopen instead of at link time: a process that also imports PyTorch must end up with
// ONE libnccl.so.2 (PyTorch bundles a newer one than the system's, same SONAME), whichever side loads first.
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static ncclComm_t g_comm = nullptr;

static int nccl_bind() {
    if (g_nccl.lib) return PB2_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);       // already in the process (e.g. PyTorch's)?
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) return set_error(PB2_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
    NcclApi a;
    a.lib = h;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
    a.Reduce = (decltype(a.Reduce))dlsym(h, "ncclReduce");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.Reduce || !a.GetErrorString)
        return set_error(PB2_ERR_NCCL, "libnccl.so.2 lacks a required symbol");
    g_nccl = a;
    return PB2_OK;
}

}  // namespace pb2

using namespace pb2;

extern "C" {

// ---- Film ---------------------------------------------------------------------------------------------------------------------
int pb2_film_create(const pb2_film_desc* desc, pb2_film** out) {
    if (!desc || !out) return set_error(PB2_ERR_INVALID, "null argument");
    *out = nullptr;
    if (desc->res_x <= 0 || desc->res_y <= 0 || !(desc->radius_x > 0.0f) || !(desc->radius_y > 0.0f))
        return set_error(PB2_ERR_INVALID, "bad film description");
    if (desc->filter < PB2_FILTER_BOX || desc->filter > PB2_FILTER_SINC) return set_error(PB2_ERR_INVALID, "unknown filter %d", desc->filter);
    if (desc->filter == PB2_FILTER_SINC && !(desc->sinc_tau > 0.0f)) return set_error(PB2_ERR_INVALID, "LanczosSincFilter needs tau > 0");
    const float* cw = desc->crop_window;
    const bool full = cw[0] == 0.0f && cw[1] == 0.0f && cw[2] == 0.0f && cw[3] == 0.0f;
    const float c0x = full ? 0.0f : cw[0], c0y = full ? 0.0f : cw[1], c1x = full ? 1.0f : cw[2], c1y = full ? 1.0f : cw[3];
    if (!(c0x >= 0.0f && c0y >= 0.0f && c1x <= 1.0f && c1y <= 1.0f && c0x < c1x && c0y < c1y))
        return set_error(PB2_ERR_INVALID, "crop window {%g, %g, %g, %g} is not inside [0, 1]^2 with min < max", c0x, c0y, c1x, c1y);
    pb2_film* f = new pb2_film();
    f->desc = *desc;
    // cropped_pixel_bounds (film.rs:41-50)
    f->px0 = (int)std::ceil((float)desc->res_x * c0x);
    f->py0 = (int)std::ceil((float)desc->res_y * c0y);
    f->px1 = (int)std::ceil((float)desc->res_x * c1x);
    f->py1 = (int)std::ceil((float)desc->res_y * c1y);
    if (f->px1 <= f->px0 || f->py1 <= f->py0) { delete f; return set_error(PB2_ERR_INVALID, "crop window holds no pixel"); }
    // Film::new filter table (film.rs:53-63) with Filter::evaluate (boxf.rs:26-28, gaussian.rs:17-39, triangle.rs:20-22,
    // mitchell.rs:24-45, sinc.rs:22-44); exp / sin are evaluated once here by the host's libm
    const float a = desc->gaussian_alpha, rx = desc->radius_x, ry = desc->radius_y;
    const float ex = std::exp(-a * rx * rx), ey = std::exp(-a * ry * ry);
    const float B = desc->mitchell_b, Cc = desc->mitchell_c, tau = desc->sinc_tau;
    auto mitchell_1d = [&](float x) {
        x = std::fabs(2.0f * x);
        if (x > 1.0f) return ((-B - 6.0f * Cc) * x * x * x + (6.0f * B + 30.0f * Cc) * x * x + (-12.0f * B - 48.0f * Cc) * x + (8.0f * B + 24.0f * Cc)) * (1.0f / 6.0f);
        return ((12.0f - 9.0f * B - 6.0f * Cc) * x * x * x + (-18.0f + 12.0f * B + 6.0f * Cc) * x * x + (6.0f - 2.0f * B)) * (1.0f / 6.0f);
    };
    auto sinc = [](float x) {
        x = std::fabs(x);
        if (x < 1e-5f) return 1.0f;
        return std::sin(PB2_PI * x) / (PB2_PI * x);
    };
    auto windowed_sinc = [&](float x, float radius) {
        x = std::fabs(x);
        if (x > radius) return 0.0f;
        const float lanczos = sinc(x / tau);
        return sinc(x) * lanczos;
    };
    for (int y = 0; y < 16; ++y)
        for (int x = 0; x < 16; ++x) {
            const float px = ((float)x + 0.5f) * rx / 16.0f, py = ((float)y + 0.5f) * ry / 16.0f;
            float w = 1.0f;
            if (desc->filter == PB2_FILTER_GAUSSIAN) w = std::fmax(std::exp(-a * px * px) - ex, 0.0f) * std::fmax(std::exp(-a * py * py) - ey, 0.0f);
            else if (desc->filter == PB2_FILTER_TRIANGLE) w = std::fmax(rx - std::fabs(px), 0.0f) * std::fmax(ry - std::fabs(py), 0.0f);
            else if (desc->filter == PB2_FILTER_MITCHELL) w = mitchell_1d(px * (1.0f / rx)) * mitchell_1d(py * (1.0f / ry));
            else if (desc->filter == PB2_FILTER_SINC) w = windowed_sinc(px, rx) * windowed_sinc(py, ry);
            f->table[y * 16 + x] = w;
        }
    // Film::get_sample_bounds (film.rs:76-81, D42 FIX)
    f->sb_x0 = (int)std::floor((float)f->px0 + 0.5f - desc->radius_x);
    f->sb_y0 = (int)std::floor((float)f->py0 + 0.5f - desc->radius_y);
    f->sb_x1 = (int)std::ceil((float)f->px1 - 0.5f + desc->radius_x);
    f->sb_y1 = (int)std::ceil((float)f->py1 - 0.5f + desc->radius_y);
    const size_t npix = f->n_pixels();
    f->stray_capacity = 1u << 20;
    const size_t stray_bytes = (size_t)f->stray_capacity * (2 * 8 + 2 * 4) + film_sort_scratch_bytes(f->stray_capacity) + 1024;
    cudaError_t e = cudaGetDevice(&f->device);
    if (e == cudaSuccess) e = cudaMalloc(&f->d_table, sizeof f->table);
    if (e == cudaSuccess) e = cudaMalloc(&f->d_xyzw, npix * 16);
    if (e == cudaSuccess) e = cudaMalloc(&f->d_acc, npix * 16);
    if (e == cudaSuccess) e = cudaMalloc(&f->d_stray, stray_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&f->d_stray_vals, (size_t)f->stray_capacity * 16);
    if (e == cudaSuccess) e = cudaMalloc(&f->d_counters, C_COUNT * 8);
    if (e == cudaSuccess) e = cudaMemcpy(f->d_table, f->table, sizeof f->table, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(f->d_xyzw, 0, npix * 16);
    if (e == cudaSuccess) e = cudaMemset(f->d_acc, 0, npix * 16);
    if (e == cudaSuccess) e = cudaMemset(f->d_counters, 0, C_COUNT * 8);
    if (e != cudaSuccess) { pb2_film_destroy(f); return cuda_fail(e, "film allocation", __FILE__, __LINE__); }
    *out = f;
    return PB2_OK;
}

int pb2_film_destroy(pb2_film* f) {
    if (!f) return PB2_OK;
    cudaFree(f->d_splat);
    cudaFree(f->d_table); cudaFree(f->d_xyzw); cudaFree(f->d_acc); cudaFree(f->d_stray); cudaFree(f->d_stray_vals); cudaFree(f->d_counters);
    f->chain.destroy();
    delete f;
    return PB2_OK;
}

int pb2_film_clear(pb2_film* f) {
    if (!f) return set_error(PB2_ERR_INVALID, "null film");
    const size_t npix = f->n_pixels();
    PB2_CUDA(f->chain.enter(0));
    PB2_CUDA(cudaMemsetAsync(f->d_xyzw, 0, npix * 16, 0));
    PB2_CUDA(cudaMemsetAsync(f->d_acc, 0, npix * 16, 0));
    if (f->d_splat) PB2_CUDA(cudaMemsetAsync(f->d_splat, 0, npix * 16, 0));
    PB2_CUDA(f->chain.leave(0));
    return PB2_OK;
}

// Film::add_splat (film.rs:137-151) for n splats
int pb2_film_add_splats(pb2_film* f, const float* p_film, const float* v_rgb, uint64_t n) {
    if (!f) return set_error(PB2_ERR_INVALID, "null film");
    if (n != 0 && (!p_film || !v_rgb)) return set_error(PB2_ERR_INVALID, "null splat arrays");
    if (!f->d_splat) {
        // (n == 0 still creates the zeroed splat plane: pb2_film_reduce sums it only on films that have one, and a
        // collective needs every rank to take part — a rank with nothing to splat calls this with n = 0)
        PB2_CUDA(cudaMalloc(&f->d_splat, f->n_pixels() * 16));
        PB2_CUDA(cudaMemset(f->d_splat, 0, f->n_pixels() * 16));
    }
    if (n == 0) return PB2_OK;
    PB2_CUDA(f->chain.enter(0));
    float *d_p = nullptr, *d_v = nullptr;
    cudaError_t e = cudaMalloc(&d_p, n * 8);
    if (e == cudaSuccess) e = cudaMalloc(&d_v, n * 12);
    if (e == cudaSuccess) e = cudaMemcpy(d_p, p_film, n * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_v, v_rgb, n * 12, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) { film_add_splats(film_view(f), d_p, d_v, n, 0); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaFree(d_p); cudaFree(d_v);
    if (e != cudaSuccess) return cuda_fail(e, "film add_splats", __FILE__, __LINE__);
    return PB2_OK;
}

// Film::set_image (film.rs:125-135): every pixel = to_xyz(rgb), weight 1, splat 0
int pb2_film_set_image(pb2_film* f, const float* rgb) {
    if (!f || !rgb) return set_error(PB2_ERR_INVALID, "null argument");
    const size_t npix = f->n_pixels();
    float* d = nullptr;
    PB2_CUDA(f->chain.enter(0));
    PB2_CUDA(cudaMalloc(&d, npix * 12));
    cudaError_t e = cudaMemcpy(d, rgb, npix * 12, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) { film_set_image(film_view(f), d, 0); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaFree(d);
    if (e != cudaSuccess) return cuda_fail(e, "film set_image", __FILE__, __LINE__);
    return PB2_OK;
}

int pb2_film_add_samples(pb2_film* f, const float* p_film, const float* L_rgb, const float* weight, uint64_t n) {
    if (!f) return set_error(PB2_ERR_INVALID, "null film");
    if (n == 0) return PB2_OK;
    if (!p_film || !L_rgb || !weight) return set_error(PB2_ERR_INVALID, "null sample arrays");
    PB2_CUDA(f->chain.enter(0));
    float *d_p = nullptr, *d_L = nullptr, *d_w = nullptr;
    cudaError_t e = cudaMalloc(&d_p, n * 8);
    if (e == cudaSuccess) e = cudaMalloc(&d_L, n * 12);
    if (e == cudaSuccess) e = cudaMalloc(&d_w, n * 4);
    if (e == cudaSuccess) e = cudaMemcpy(d_p, p_film, n * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_L, L_rgb, n * 12, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_w, weight, n * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) { film_add_samples(film_view(f), d_p, d_L, d_w, n, 0); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaFree(d_p); cudaFree(d_L); cudaFree(d_w);
    if (e != cudaSuccess) return cuda_fail(e, "film add_samples", __FILE__, __LINE__);
    return PB2_OK;
}

int pb2_film_read_xyzw(pb2_film* f, float* out) {
    if (!f || !out) return set_error(PB2_ERR_INVALID, "null argument");
    PB2_CUDA(f->chain.enter(0));
    PB2_CUDA(cudaMemcpy(out, f->d_xyzw, f->n_pixels() * 16, cudaMemcpyDeviceToHost));
    return PB2_OK;
}

int pb2_film_resolve_rgb(pb2_film* f, float scale, float* rgb) { return pb2_film_resolve_rgb_splat(f, scale, 1.0f, rgb); }

int pb2_film_resolve_rgb_splat(pb2_film* f, float scale, float splat_scale, float* rgb) {
    if (!f || !rgb) return set_error(PB2_ERR_INVALID, "null argument");
    const size_t npix = f->n_pixels();
    float* d = nullptr;
    PB2_CUDA(f->chain.enter(0));
    PB2_CUDA(cudaMalloc(&d, npix * 12));
    film_resolve(film_view(f), scale, splat_scale, d, 0);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(rgb, d, npix * 12, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return cuda_fail(e, "film resolve", __FILE__, __LINE__);
    return PB2_OK;
}

int pb2_film_write_image(pb2_film* f, const char* filename, float scale) {
    if (!f || !filename) return set_error(PB2_ERR_INVALID, "null argument");
    const std::string name(filename);
    const size_t dot = name.rfind('.');
    const std::string ext = dot == std::string::npos ? "" : name.substr(dot);
    if (ext != ".pfm" && ext != ".ppm") return set_error(PB2_ERR_INVALID, "unsupported image extension '%s' (.pfm or .ppm)", ext.c_str());
    const int w = f->px1 - f->px0, h = f->py1 - f->py0;      // cropped_pixel_bounds, as Film::write_image hands to write_image
    std::vector<float> rgb((size_t)w * h * 3);
    int rc = pb2_film_resolve_rgb(f, scale, rgb.data());
    if (rc != PB2_OK) return rc;
    FILE* fp = fopen(filename, "wb");
    if (!fp) return set_error(PB2_ERR_INVALID, "cannot open '%s' for writing", filename);
    bool ok = true;
    if (ext == ".pfm") {
        ok = fprintf(fp, "PF\n%d %d\n-1.0\n", w, h) > 0;
        for (int y = h - 1; y >= 0 && ok; --y) ok = fwrite(rgb.data() + (size_t)y * w * 3, sizeof(float), (size_t)w * 3, fp) == (size_t)w * 3;
    } else {
        ok = fprintf(fp, "P6\n%d %d\n255\n", w, h) > 0;
        std::vector<unsigned char> row((size_t)w * 3);
        for (int y = 0; y < h && ok; ++y) {
            for (int i = 0; i < w * 3; ++i) {
                const float v = rgb[(size_t)y * w * 3 + i];
                const float g = v <= 0.0031308f ? 12.92f * v : 1.055f * powf(v, 1.0f / 2.4f) - 0.055f;
                float b = 255.0f * g + 0.5f;
                b = b < 0.0f ? 0.0f : (b > 255.0f ? 255.0f : b);
                row[i] = (unsigned char)b;
            }
            ok = fwrite(row.data(), 1, row.size(), fp) == row.size();
        }
    }
    ok = (fclose(fp) == 0) && ok;
    if (!ok) return set_error(PB2_ERR_INVALID, "short write to '%s'", filename);
    return PB2_OK;
}

int pb2_film_device_ptr(pb2_film* f, void** d_xyzw, uint64_t* n_floats) {
    if (!f) return set_error(PB2_ERR_INVALID, "null film");
    if (d_xyzw) *d_xyzw = f->d_xyzw;
    if (n_floats) *n_floats = (uint64_t)f->n_pixels() * 4;
    return PB2_OK;
}

int pb2_film_bounds(const pb2_film* f, int32_t pixel_bounds[4], int32_t sample_bounds[4]) {
    if (!f) return set_error(PB2_ERR_INVALID, "null film");
    if (pixel_bounds) { pixel_bounds[0] = f->px0; pixel_bounds[1] = f->py0; pixel_bounds[2] = f->px1; pixel_bounds[3] = f->py1; }
    if (sample_bounds) { sample_bounds[0] = f->sb_x0; sample_bounds[1] = f->sb_y0; sample_bounds[2] = f->sb_x1; sample_bounds[3] = f->sb_y1; }
    return PB2_OK;
}

// ---- Integrator::render ---------------------------------------------------------------------------------------------------------
int pb2_render_path(pb2_scene* scene, const pb2_camera* cam, const pb2_path_desc* path, pb2_film* film, void* stream) {
    if (!film) return set_error(PB2_ERR_INVALID, "null film");
    CameraView cv;
    int rc = check_path_args(scene, cam, path, &film->desc, &cv);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(scene->mu);
    const FilmView fv = film_view(film);
    rc = ensure_wavefront(scene, (uint64_t)fv.sb_w * fv.sb_h);
    if (rc) return rc;
    // the wavefront arena, its counters and the cached sampler tables are shared by every render of this scene, and the film
    // may have been cleared / rendered into on another stream: order this call behind the previous uses of both
    PB2_CUDA(scene->path_chain.enter((cudaStream_t)stream));
    PB2_CUDA(film->chain.enter((cudaStream_t)stream));
    const PathParams pp{path->max_depth, path->rr_threshold, path->integrator};
    SamplerView smp;
    rc = sampler_view(scene, path, fv.sb_x0, fv.sb_y0, fv.sb_w, fv.sb_h, (cudaStream_t)stream, &smp);
    if (rc) return rc;
    if (path->light_strategy == PB2_LIGHTS_SPATIAL && scene->lights.size() > 1) {      // PathIntegrator::pre_process (path.rs:58-63)
        rc = ensure_spatial(scene, (cudaStream_t)stream);
        if (rc) return rc;
    }
    wavefront_render(scene->wf, scene->view, shade_view(scene, path->light_strategy), cv, fv, pp, smp, path->spp, path->sample_begin,
                     path->sample_end, (cudaStream_t)stream);
    PB2_CUDA(cudaGetLastError());
    PB2_CUDA(scene->path_chain.leave((cudaStream_t)stream));
    PB2_CUDA(film->chain.leave((cudaStream_t)stream));
    return PB2_OK;
}

int pb2_path_li(pb2_scene* scene, const pb2_camera* cam, const pb2_path_desc* path, const uint32_t* pixel_xy,
                const uint32_t* sample_index, uint64_t n, float* L_rgb, float* p_film) {
    CameraView cv;
    int rc = check_path_args(scene, cam, path, nullptr, &cv);
    if (rc) return rc;
    if (n == 0) return PB2_OK;
    if (!pixel_xy || !sample_index || !L_rgb || !p_film) return set_error(PB2_ERR_INVALID, "null argument");
    // sampler streams, PixelSampler tables and the Halton offsets are indexed by (pixel, sample): anything outside the image or
    // beyond spp would read past them
    if (n > (1ull << 28)) return set_error(PB2_ERR_LIMIT, "pb2_path_li: %llu samples exceed the 2^28 path slots of one wavefront", (unsigned long long)n);
    for (uint64_t i = 0; i < n; ++i) {
        if (pixel_xy[2 * i] >= (uint32_t)cam->res_x || pixel_xy[2 * i + 1] >= (uint32_t)cam->res_y)
            return set_error(PB2_ERR_INVALID, "pb2_path_li: pixel %llu = (%u, %u) lies outside the %d x %d image", (unsigned long long)i,
                             pixel_xy[2 * i], pixel_xy[2 * i + 1], cam->res_x, cam->res_y);
        if (sample_index[i] >= (uint32_t)path->spp)
            return set_error(PB2_ERR_INVALID, "pb2_path_li: sample index %u of entry %llu is not below spp = %d", sample_index[i], (unsigned long long)i, path->spp);
    }
    std::lock_guard<std::mutex> lock(scene->mu);
    rc = ensure_wavefront(scene, n);
    if (rc) return rc;
    PB2_CUDA(scene->path_chain.enter(0));
    // box filter, r = 0.5 sample bounds: streams are indexed by image pixel
    FilmView fv;
    memset(&fv, 0, sizeof fv);
    fv.px1 = cam->res_x; fv.py1 = cam->res_y; fv.sb_w = cam->res_x; fv.sb_h = cam->res_y; fv.radius_x = fv.radius_y = 0.5f;
    fv.sb_w_magic = fast_div_magic((uint32_t)fv.sb_w);
    fv.max_lum = __builtin_huge_valf();
    int32_t* d_xy = nullptr;
    uint32_t* d_s = nullptr;
    float *d_L = nullptr, *d_pf = nullptr;
    SamplerView smp;
    rc = sampler_view(scene, path, fv.sb_x0, fv.sb_y0, fv.sb_w, fv.sb_h, 0, &smp);
    if (rc) return rc;
    if (path->light_strategy == PB2_LIGHTS_SPATIAL && scene->lights.size() > 1) {
        rc = ensure_spatial(scene, 0);
        if (rc) return rc;
    }
    cudaError_t e = cudaMalloc(&d_xy, n * 8);
    if (e == cudaSuccess) e = cudaMalloc(&d_s, n * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_L, n * 12);
    if (e == cudaSuccess) e = cudaMalloc(&d_pf, n * 8);
    if (e == cudaSuccess) e = cudaMemcpy(d_xy, pixel_xy, n * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_s, sample_index, n * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        const PathParams pp{path->max_depth, path->rr_threshold, path->integrator};
        wavefront_li(scene->wf, scene->view, shade_view(scene, path->light_strategy), cv, fv, pp, smp, path->spp, d_xy, d_s, n, d_L, d_pf, 0);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(L_rgb, d_L, n * 12, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(p_film, d_pf, n * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = scene->path_chain.leave(0);
    cudaFree(d_xy); cudaFree(d_s); cudaFree(d_L); cudaFree(d_pf);
    if (e != cudaSuccess) return cuda_fail(e, "pb2_path_li", __FILE__, __LINE__);
    return PB2_OK;
}

int pb2_spatial_light_distribution(pb2_scene* scene, int32_t n_voxels[3], float* func, float* cdf, float* func_int) {
    if (!scene || !n_voxels) return set_error(PB2_ERR_INVALID, "null argument");
    if (!scene->built) return set_error(PB2_ERR_STATE, "pb2_scene_build_bvh has not been called");
    if (!scene->d_tri_material) return set_error(PB2_ERR_STATE, "the scene was created without materials");
    std::lock_guard<std::mutex> lock(scene->mu);
    int rc = ensure_spatial(scene, 0);
    if (rc) return rc;
    const SpatialView g = spatial_view(scene);
    const size_t n = scene->lights.size(), n_vox = (size_t)g.nv[0] * g.nv[1] * g.nv[2];
    for (int i = 0; i < 3; ++i) n_voxels[i] = g.nv[i];
    if (func) PB2_CUDA(cudaMemcpy(func, g.func, n_vox * n * sizeof(float), cudaMemcpyDeviceToHost));
    if (cdf) PB2_CUDA(cudaMemcpy(cdf, g.cdf, n_vox * (n + 1) * sizeof(float), cudaMemcpyDeviceToHost));
    if (func_int) PB2_CUDA(cudaMemcpy(func_int, g.func_int, n_vox * sizeof(float), cudaMemcpyDeviceToHost));
    return PB2_OK;
}

int pb2_render_counters(pb2_scene* scene, uint64_t out[8]) {
    if (!scene || !out) return set_error(PB2_ERR_INVALID, "null argument");
    for (int i = 0; i < 8; ++i) out[i] = 0;
    if (!scene->wf) return PB2_OK;
    unsigned long long c[C_COUNT];
    PB2_CUDA(cudaDeviceSynchronize());
    PB2_CUDA(cudaMemcpy(c, scene->wf->b.counters, sizeof c, cudaMemcpyDeviceToHost));
    out[0] = c[T_CAMERA]; out[1] = c[T_EXTEND]; out[2] = c[T_SHADOW]; out[3] = c[T_MIS];
    out[4] = scene->wf->totals[4];
    out[5] = c[C_STRAY_OVERFLOW];
    if (scene->wf->peer) {                                 // the second wavefront of frames rendered two batches at a time
        PB2_CUDA(cudaMemcpy(c, scene->wf->peer->b.counters, sizeof c, cudaMemcpyDeviceToHost));
        out[0] += c[T_CAMERA]; out[1] += c[T_EXTEND]; out[2] += c[T_SHADOW]; out[3] += c[T_MIS];
        out[4] += scene->wf->peer->totals[4];
    }
    return PB2_OK;
}

// ---- NCCL film reduce -------------------------------------------------------------------------------------------------------------
int pb2_nccl_unique_id(char id[128]) {
    if (!id) return set_error(PB2_ERR_INVALID, "null id");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    int rc = nccl_bind();
    if (rc) return rc;
    ncclUniqueId u;
    ncclResult_t r = g_nccl.GetUniqueId(&u);
    if (r != ncclSuccess) return set_error(PB2_ERR_NCCL, "ncclGetUniqueId: %s", g_nccl.GetErrorString(r));
    memcpy(id, &u, 128);
    return PB2_OK;
}

int pb2_nccl_init(const char id[128], int rank, int n_ranks) {
    if (!id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return set_error(PB2_ERR_INVALID, "bad NCCL init arguments");
    if (g_comm) return set_error(PB2_ERR_STATE, "NCCL communicator already initialised");
    int rc = nccl_bind();
    if (rc) return rc;
    ncclUniqueId u;
    memcpy(&u, id, 128);
    ncclResult_t r = g_nccl.CommInitRank(&g_comm, n_ranks, u, rank);
    if (r != ncclSuccess) { g_comm = nullptr; return set_error(PB2_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString(r)); }
    return PB2_OK;
}

int pb2_nccl_shutdown(void) {
    if (g_comm) { g_nccl.CommDestroy(g_comm); g_comm = nullptr; }
    return PB2_OK;
}

int pb2_film_reduce(pb2_film* f, int root, void* stream) {
    if (!f) return set_error(PB2_ERR_INVALID, "null film");
    if (!g_comm) return set_error(PB2_ERR_STATE, "pb2_nccl_init has not been called");
    const size_t count = f->n_pixels() * 4;
    PB2_CUDA(f->chain.enter((cudaStream_t)stream));
    ncclResult_t r = g_nccl.Reduce(f->d_xyzw, f->d_xyzw, count, ncclFloat32, ncclSum, root, g_comm, (cudaStream_t)stream);
    // Pixel::splat_xyz (film.rs:9-15) is part of the film: summed too when this film has a splat plane (every rank's must —
    // see pb2_film_add_splats with n = 0)
    if (r == ncclSuccess && f->d_splat) r = g_nccl.Reduce(f->d_splat, f->d_splat, count, ncclFloat32, ncclSum, root, g_comm, (cudaStream_t)stream);
    if (r != ncclSuccess) return set_error(PB2_ERR_NCCL, "ncclReduce: %s", g_nccl.GetErrorString(r));
    PB2_CUDA(f->chain.leave((cudaStream_t)stream));
    return PB2_OK;
}

}  // extern "C"
