// api.cu — the C ABI of include/pbrt_b200.h: scene handles, BVH upload, batched intersect / intersect_p,
// camera rays, RNG parity hook.  (Film / path tracer / NCCL entry points live in api_path.cu.)
#include "api_internal.hpp"

#include <cmath>
#include <thread>
#include <cstdarg>
#include <cstdio>

namespace pb2 {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    return set_error(PB2_ERR_CUDA, "%s failed at %s:%d: %s", what, file, line, cudaGetErrorString(e));
}

void camera_setup(const float pos[3], const float look[3], const float up[3], float fov, int res_x, int res_y, CameraView* out);

int make_camera_view(const pb2_camera* cam, CameraView* out) {
    if (!cam || cam->res_x <= 0 || cam->res_y <= 0 || !(cam->fov > 0.0f && cam->fov < 180.0f))
        return set_error(PB2_ERR_INVALID, "invalid camera description");
    if (!(cam->lens_radius >= 0.0f) || (cam->lens_radius > 0.0f && !(cam->focal_distance > 0.0f)))
        return set_error(PB2_ERR_INVALID, "thin lens needs lens_radius >= 0 and focal_distance > 0");
    camera_setup(cam->pos, cam->look, cam->up, cam->fov, cam->res_x, cam->res_y, out);
    out->lens_radius = cam->lens_radius;
    out->focal_distance = cam->focal_distance;
    return PB2_OK;
}

}  // namespace pb2

using namespace pb2;

void pb2_scene::free_device() {
    if (d_pairs) cudaFree(d_pairs);
    if (d_quads) cudaFree(d_quads);
    if (d_lin_nodes) cudaFree(d_lin_nodes);
    if (d_lin_prims) cudaFree(d_lin_prims);
    d_lin_nodes = d_lin_prims = nullptr;
    if (d_tris) cudaFree(d_tris);
    if (d_slot_of_prim) cudaFree(d_slot_of_prim);
    if (d_tris_prim) cudaFree(d_tris_prim);
    d_tris_prim = nullptr;
    if (d_tri_material) cudaFree(d_tri_material);
    if (d_tri_light) cudaFree(d_tri_light);
    if (d_materials) cudaFree(d_materials);
    if (d_lights) cudaFree(d_lights);
    if (d_light_cdf) cudaFree(d_light_cdf);
    if (d_spatial) cudaFree(d_spatial);
    d_spatial = nullptr;
    if (d_spheres) cudaFree(d_spheres);
    d_spheres = nullptr;
    if (d_media) cudaFree(d_media);
    if (d_prim_inside) cudaFree(d_prim_inside);
    if (d_prim_outside) cudaFree(d_prim_outside);
    d_media = d_prim_inside = d_prim_outside = nullptr;
    path_chain.destroy();
    if (d_counters) cudaFree(d_counters);
    d_counters = nullptr;
    if (d_indices) cudaFree(d_indices);
    if (d_normals) cudaFree(d_normals);
    if (d_tangents) cudaFree(d_tangents);
    if (d_uvs) cudaFree(d_uvs);
    d_indices = d_normals = d_tangents = d_uvs = nullptr;
    d_quads = nullptr;
    d_pairs = d_tris = d_slot_of_prim = d_tri_material = d_tri_light = d_materials = d_lights = d_light_cdf = nullptr;
    if (pipe.pending && pipe.d2h) cudaStreamSynchronize(pipe.d2h);      // _async batches still in flight
    pipe.pending = false;
    pipe.chunks = 0;
    pipe.batches = 0;
    for (cudaEvent_t& e : pipe.batch_done) { if (e) cudaEventDestroy(e); e = nullptr; }
    for (int i = 0; i < kStages; ++i) {
        Stage& st = pipe.slot[i];
        if (st.d_in) cudaFree(st.d_in);
        if (st.d_out) cudaFree(st.d_out);
        if (st.d_aux) cudaFree(st.d_aux);
        if (st.in_ready) cudaEventDestroy(st.in_ready);
        if (st.done) cudaEventDestroy(st.done);
        if (st.drained) cudaEventDestroy(st.drained);
        st = Stage();
    }
    if (pipe.h2d) cudaStreamDestroy(pipe.h2d);
    for (int i = 0; i < 2; ++i) { if (pipe.compute[i]) cudaStreamDestroy(pipe.compute[i]); pipe.compute[i] = nullptr; }
    if (pipe.d2h) cudaStreamDestroy(pipe.d2h);
    pipe.h2d = pipe.d2h = nullptr;
    if (wf) { wavefront_destroy(wf); wf = nullptr; }
    if (d_halton_perms) cudaFree(d_halton_perms);
    if (d_halton_primes) cudaFree(d_halton_primes);
    if (d_halton_sums) cudaFree(d_halton_sums);
    d_halton_perms = d_halton_primes = d_halton_sums = nullptr;
    if (d_sobol) cudaFree(d_sobol);
    d_sobol = nullptr;
    if (d_tab1) cudaFree(d_tab1);
    if (d_tab2) cudaFree(d_tab2);
    d_tab1 = d_tab2 = nullptr;
}

// fn(begin, end) over [0, n) on the host's hardware threads (one range per thread; small inputs stay on the caller's thread).
template <class F>
static void parallel_chunks(uint64_t n, F&& fn) {
    const uint64_t t = std::max<uint64_t>(1, std::min<uint64_t>(std::thread::hardware_concurrency(), n / (1u << 20) + 1));
    if (t == 1) { fn((uint64_t)0, n); return; }
    std::vector<std::thread> pool;
    const uint64_t step = (n + t - 1) / t;
    for (uint64_t k = 0; k < t; ++k) {
        const uint64_t lo = k * step, hi = std::min(n, lo + step);
        if (lo < hi) pool.emplace_back([&fn, lo, hi] { fn(lo, hi); });
    }
    for (auto& th : pool) th.join();
}
// The smallest i in [0, n) with pred(i), or UINT64_MAX.
template <class P>
static uint64_t first_index_where(uint64_t n, P&& pred) {
    std::atomic<uint64_t> first{UINT64_MAX};
    parallel_chunks(n, [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; ++i)
            if (pred(i)) {
                uint64_t cur = first.load();
                while (i < cur && !first.compare_exchange_weak(cur, i)) {}
                return;
            }
    });
    return first.load();
}

extern "C" {

const char* pb2_last_error(void) { return g_err; }

int pb2_set_trace_tuning(int refill_below, int node_quorum, int leaf_quorum, int prefetch) {
    if (refill_below > 33 || node_quorum > 33 || leaf_quorum > 33) return set_error(PB2_ERR_INVALID, "quorums are lane counts (<= 33)");
    pb2::set_trace_tuning(refill_below, node_quorum, leaf_quorum, prefetch);
    return PB2_OK;
}

int pb2_device_count(int* out) {
    if (!out) return set_error(PB2_ERR_INVALID, "null out");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *out = 0; return cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__); }
    *out = n;
    return PB2_OK;
}

int pb2_init(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return set_error(PB2_ERR_CUDA, "no CUDA device available (%s); this backend has no CPU fallback",
                         e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return set_error(PB2_ERR_INVALID, "device %d out of range [0,%d)", device, n);
    PB2_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PB2_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return set_error(PB2_ERR_CUDA, "device %d is sm_%d%d; the kernels are built for sm_100a only", device, prop.major, prop.minor);
    PB2_CUDA(cudaFree(0));
    return PB2_OK;
}

int pb2_shutdown(void) {
    pb2_nccl_shutdown();
    return PB2_OK;
}

int pb2_host_alloc(uint64_t bytes, void** out) {
    if (!out) return set_error(PB2_ERR_INVALID, "null out");
    // PB2_HOST_ALLOC_WC=1 (tuning runs only): write-combined pinned memory — faster for the device to read over PCIe on some
    // hosts, very slow for the CPU to read back, so only for buffers the host writes and the device reads (ray batches)
    static const bool wc = getenv("PB2_HOST_ALLOC_WC") && atoi(getenv("PB2_HOST_ALLOC_WC")) != 0;
    PB2_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
    return PB2_OK;
}
int pb2_host_free(void* p) {
    if (p) PB2_CUDA(cudaFreeHost(p));
    return PB2_OK;
}
int pb2_device_alloc(uint64_t bytes, void** out) {
    if (!out) return set_error(PB2_ERR_INVALID, "null out");
    PB2_CUDA(cudaMalloc(out, bytes ? bytes : 1));
    return PB2_OK;
}
int pb2_device_free(void* p) {
    if (p) PB2_CUDA(cudaFree(p));
    return PB2_OK;
}
int pb2_memcpy_h2d(void* dst, const void* src, uint64_t bytes) {
    PB2_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return PB2_OK;
}
int pb2_memcpy_d2h(void* dst, const void* src, uint64_t bytes) {
    PB2_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return PB2_OK;
}
int pb2_device_synchronize(void) {
    PB2_CUDA(cudaDeviceSynchronize());
    return PB2_OK;
}

int pb2_scene_create(const float* verts, uint64_t n_verts, const uint32_t* indices, uint64_t n_tris,
                     const uint32_t* tri_material, const pb2_material* mats, uint32_t n_mats,
                     const pb2_light* lights, uint32_t n_lights, pb2_scene** out) {
    if (!out) return set_error(PB2_ERR_INVALID, "null out");
    *out = nullptr;
    if ((n_verts && !verts) || (n_tris && !indices)) return set_error(PB2_ERR_INVALID, "null mesh arrays");
    if (n_tris >= (1ull << 29)) return set_error(PB2_ERR_LIMIT, "at most 2^29-1 triangles (29-bit references in the device BVH records)");
    // validation and the two copies below run on the host's threads: a 10 M-triangle mesh is 180 MB, 0.1 s on one core
    {
        const uint64_t bad_v = first_index_where(3 * n_verts, [&](uint64_t i) { return !std::isfinite(verts[i]); });
        if (bad_v != UINT64_MAX) return set_error(PB2_ERR_INVALID, "vertex %llu has a non-finite coordinate", (unsigned long long)(bad_v / 3));
        const uint64_t bad_i = first_index_where(3 * n_tris, [&](uint64_t i) { return indices[i] >= n_verts; });
        if (bad_i != UINT64_MAX)
            return set_error(PB2_ERR_INVALID, "triangle %llu references vertex %u >= %llu", (unsigned long long)(bad_i / 3), indices[bad_i], (unsigned long long)n_verts);
    }
    if (tri_material) {
        if (!mats || n_mats == 0) return set_error(PB2_ERR_INVALID, "tri_material given without materials");
        for (uint64_t i = 0; i < n_tris; ++i)
            if (tri_material[i] >= n_mats && tri_material[i] != PB2_NO_MATERIAL) return set_error(PB2_ERR_INVALID, "triangle %llu references material %u >= %u", (unsigned long long)i, tri_material[i], n_mats);
    }
    for (uint32_t i = 0; i < n_lights; ++i) {
        // (an area light's prim_id is checked at build time: it may name a sphere added by pb2_scene_add_spheres)
        if (lights[i].type < PB2_LIGHT_POINT || lights[i].type > PB2_LIGHT_DISTANT) return set_error(PB2_ERR_INVALID, "light %u has unknown type %d", i, lights[i].type);
        if (lights[i].type == PB2_LIGHT_DISTANT || lights[i].type == PB2_LIGHT_SPOT) {
            const float* a = lights[i].axis;
            if (!(std::isfinite(a[0]) && std::isfinite(a[1]) && std::isfinite(a[2])) || (a[0] == 0.0f && a[1] == 0.0f && a[2] == 0.0f))
                return set_error(PB2_ERR_INVALID, "light %u needs a finite non-zero axis", i);
        }
    }
    pb2_scene* s = new pb2_scene();
    s->verts.resize(3 * n_verts);
    s->indices.resize(3 * n_tris);
    parallel_chunks(3 * n_verts, [&](uint64_t lo, uint64_t hi) { memcpy(s->verts.data() + lo, verts + lo, (hi - lo) * sizeof(float)); });
    parallel_chunks(3 * n_tris, [&](uint64_t lo, uint64_t hi) { memcpy(s->indices.data() + lo, indices + lo, (hi - lo) * sizeof(uint32_t)); });
    if (tri_material) s->tri_material.assign(tri_material, tri_material + n_tris);
    if (mats) s->materials.assign(mats, mats + n_mats);
    if (lights) s->lights.assign(lights, lights + n_lights);
    *out = s;
    return PB2_OK;
}

int pb2_scene_set_shading_geometry(pb2_scene* scene, const float* normals, const float* tangents, const float* uvs) {
    if (!scene) return set_error(PB2_ERR_INVALID, "null scene");
    std::lock_guard<std::mutex> lock(scene->mu);
    if (scene->built) return set_error(PB2_ERR_STATE, "set the mesh attributes before pb2_scene_build_bvh");
    const size_t nv = scene->verts.size() / 3;
    const struct { const float* p; size_t k; const char* name; } in[3] = {{normals, 3, "normal"}, {tangents, 3, "tangent"}, {uvs, 2, "uv"}};
    for (const auto& a : in)
        if (a.p)
            for (size_t i = 0; i < a.k * nv; ++i)
                if (!std::isfinite(a.p[i])) return set_error(PB2_ERR_INVALID, "%s of vertex %zu is not finite", a.name, i / a.k);
    scene->normals.assign(normals ? normals : nullptr, normals ? normals + 3 * nv : nullptr);
    scene->tangents.assign(tangents ? tangents : nullptr, tangents ? tangents + 3 * nv : nullptr);
    scene->uvs.assign(uvs ? uvs : nullptr, uvs ? uvs + 2 * nv : nullptr);
    return PB2_OK;
}

// Sphere::new (sphere.rs:229-248) + Transform::new's inverse (transform.rs:198-206) -> the 128-byte device record and
// Shape::world_bound (shape.rs:18-20, sphere.rs:31-36, transform.rs:569-606: the eight corners of the object bound).
static void make_sphere_record(const pb2_sphere& sp, uint32_t prim, DSphere* out, float bounds[6]) {
    mat4 m;
    memcpy(m.m, sp.object_to_world, sizeof m.m);
    const mat4 mi = invert_mat4(m);
    memcpy(out->m, m.m, 12 * sizeof(float));
    memcpy(out->mi, mi.m, 12 * sizeof(float));
    const float r = sp.radius;
    auto clampf = [](float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); };
    out->radius = r;
    out->z_min = clampf(fminf(sp.z_min, sp.z_max), -r, r);
    out->z_max = clampf(fmaxf(sp.z_max, sp.z_min), -r, r);
    out->theta_min = det_acos(clampf(fminf(sp.z_min, sp.z_max) / r, -1.0f, 1.0f));
    out->theta_max = det_acos(clampf(fmaxf(sp.z_max, sp.z_min) / r, -1.0f, 1.0f));
    out->phi_max = PB2_PI / 180.0f * clampf(sp.phi_max, 0.0f, 360.0f);
    const float (*a)[4] = m.m;                      // pbrt-v3 Transform::SwapsHandedness
    const float det = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
                      a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
    out->flags = (sp.reverse_orientation ? 1u : 0u) | (det < 0.0f ? 2u : 0u);
    out->prim = prim;
    const float lo[3] = {-r, -r, out->z_min}, hi[3] = {r, r, out->z_max};
    const int corner[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 1, 1}, {1, 1, 0}, {1, 0, 1}, {1, 1, 1}};
    for (int c = 0; c < 8; ++c) {
        const vec3 p = xf_point(out->m, mk(corner[c][0] ? hi[0] : lo[0], corner[c][1] ? hi[1] : lo[1], corner[c][2] ? hi[2] : lo[2]));
        const float q[3] = {p.x, p.y, p.z};
        for (int k = 0; k < 3; ++k) {
            bounds[k] = c == 0 ? q[k] : fminf(bounds[k], q[k]);
            bounds[3 + k] = c == 0 ? q[k] : fmaxf(bounds[3 + k], q[k]);
        }
    }
}

int pb2_scene_add_spheres(pb2_scene* scene, const pb2_sphere* spheres, uint32_t n) {
    if (!scene || (n && !spheres)) return set_error(PB2_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lock(scene->mu);
    if (scene->built || scene->built_host) return set_error(PB2_ERR_STATE, "add the spheres before pb2_scene_build_bvh");
    if (scene->n_primitives() + n >= (1ull << 29)) return set_error(PB2_ERR_LIMIT, "at most 2^29-1 primitives");
    for (uint32_t i = 0; i < n; ++i) {
        const pb2_sphere& sp = spheres[i];
        for (int k = 0; k < 16; ++k)
            if (!std::isfinite(sp.object_to_world[k])) return set_error(PB2_ERR_INVALID, "sphere %u: object_to_world is not finite", i);
        const float* m = sp.object_to_world;
        if (m[12] != 0.0f || m[13] != 0.0f || m[14] != 0.0f || m[15] != 1.0f)
            return set_error(PB2_ERR_INVALID, "sphere %u: object_to_world must be affine (last row 0 0 0 1)", i);
        const float det = m[0] * (m[5] * m[10] - m[6] * m[9]) - m[1] * (m[4] * m[10] - m[6] * m[8]) + m[2] * (m[4] * m[9] - m[5] * m[8]);
        if (det == 0.0f || !std::isfinite(det)) return set_error(PB2_ERR_INVALID, "sphere %u: object_to_world is singular (Matrix4x4::inverse panics, transform.rs:84)", i);
        if (!(sp.radius > 0.0f) || !std::isfinite(sp.radius)) return set_error(PB2_ERR_INVALID, "sphere %u: radius must be positive and finite", i);
        if (!std::isfinite(sp.z_min) || !std::isfinite(sp.z_max) || !std::isfinite(sp.phi_max)) return set_error(PB2_ERR_INVALID, "sphere %u: z_min / z_max / phi_max must be finite", i);
        if (!scene->materials.empty() && sp.material >= scene->materials.size() && sp.material != PB2_NO_MATERIAL)
            return set_error(PB2_ERR_INVALID, "sphere %u references material %u >= %zu", i, sp.material, scene->materials.size());
    }
    scene->spheres.insert(scene->spheres.end(), spheres, spheres + n);
    return PB2_OK;
}

int pb2_scene_set_media(pb2_scene* scene, const pb2_medium* media, uint32_t n_media, const int32_t* prim_inside, const int32_t* prim_outside,
                        int32_t camera_medium) {
    if (!scene || (n_media && !media)) return set_error(PB2_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lock(scene->mu);
    if (scene->built || scene->built_host) return set_error(PB2_ERR_STATE, "set the media before pb2_scene_build_bvh");
    for (uint32_t i = 0; i < n_media; ++i)
        for (int k = 0; k < 3; ++k) {
            if (!(media[i].sigma_a[k] >= 0.0f) || !(media[i].sigma_s[k] >= 0.0f) || !std::isfinite(media[i].sigma_a[k]) || !std::isfinite(media[i].sigma_s[k]))
                return set_error(PB2_ERR_INVALID, "medium %u: sigma_a / sigma_s must be finite and >= 0", i);
            if (!(media[i].sigma_a[k] + media[i].sigma_s[k] > 0.0f))
                return set_error(PB2_ERR_INVALID, "medium %u: sigma_t is zero in channel %d (HomogeneousMedium::sample divides by it, homogeneous.rs:44)", i, k);
        }
    for (uint32_t i = 0; i < n_media; ++i)
        if (!(std::fabs(media[i].g) < 1.0f)) return set_error(PB2_ERR_INVALID, "medium %u: |g| must be < 1", i);
    const uint64_t np = scene->n_primitives();
    for (const int32_t* side : {prim_inside, prim_outside})
        if (side)
            for (uint64_t i = 0; i < np; ++i)
                if (side[i] < -1 || side[i] >= (int32_t)n_media) return set_error(PB2_ERR_INVALID, "primitive %llu names medium %d of %u", (unsigned long long)i, side[i], n_media);
    if (camera_medium < -1 || camera_medium >= (int32_t)n_media) return set_error(PB2_ERR_INVALID, "camera medium %d of %u", camera_medium, n_media);
    scene->media.assign(media, media + n_media);
    scene->prim_inside.assign(np, -1);
    scene->prim_outside.assign(np, -1);
    if (prim_inside) scene->prim_inside.assign(prim_inside, prim_inside + np);
    if (prim_outside) scene->prim_outside.assign(prim_outside, prim_outside + np);
    scene->camera_medium = camera_medium;
    return PB2_OK;
}

int pb2_scene_destroy(pb2_scene* scene) {
    if (!scene) return PB2_OK;
    scene->free_device();
    delete scene;
    return PB2_OK;
}

static int check_build_args(int max_prims_in_node, int split_method) {
    if (split_method < 0 || split_method > 3)
        return set_error(PB2_ERR_INVALID, "split_method %d: SplitMethod::SAH (0), ::HLBVH (1, GPU build), ::Middle (2), ::EqualCounts (3) (bvh.rs:199-204)", split_method);
    if (max_prims_in_node < 1) return set_error(PB2_ERR_INVALID, "max_prims_in_node must be >= 1");
    return PB2_OK;
}

static int build_host_locked(pb2_scene* scene, int max_prims_in_node, int split_method) {
    int rc = check_build_args(max_prims_in_node, split_method);
    if (rc) return rc;
    if (split_method == 1) return set_error(PB2_ERR_INVALID, "SplitMethod::HLBVH is built on the device: use pb2_scene_build_bvh");
    scene->built = false;
    scene->built_host = false;
    const uint64_t n_tris = scene->indices.size() / 3;
    for (double& v : scene->build_ms) v = 0.0;
    for (size_t i = 0; i < scene->lights.size(); ++i)
        if (scene->lights[i].type == PB2_LIGHT_AREA && scene->lights[i].prim_id >= scene->n_primitives())
            return set_error(PB2_ERR_INVALID, "area light %zu references primitive %u >= %llu", i, scene->lights[i].prim_id, (unsigned long long)scene->n_primitives());
    const size_t n_sph = scene->spheres.size();
    scene->sphere_records.resize(n_sph);
    std::vector<float> sphere_bounds(6 * n_sph);
    for (size_t i = 0; i < n_sph; ++i) make_sphere_record(scene->spheres[i], (uint32_t)(n_tris + i), &scene->sphere_records[i], &sphere_bounds[6 * i]);
    build_sah_bvh(scene->verts.data(), scene->verts.size() / 3, scene->indices.data(), n_tris, max_prims_in_node, 0, &scene->bvh, split_method,
                  sphere_bounds.data(), n_sph);
    if (scene->bvh.max_depth > kStackDepth)
        return set_error(PB2_ERR_LIMIT, "BVH depth %d exceeds the 64-entry traversal stack of BVHAccel::intersect", scene->bvh.max_depth);
    if (scene->bvh.leaf_overflow)
        return set_error(PB2_ERR_LIMIT, "a leaf holds more than 65535 primitives with coincident centroids (16-bit n_primitives of the 32-byte node)");
    scene->n_nodes = scene->bvh.nodes.size();
    scene->n_prims = scene->bvh.ordered_prims.size();
    scene->tree_depth = scene->bvh.max_depth;
    if (scene->n_nodes)
        for (int k = 0; k < 3; ++k) { scene->root_bounds[k] = scene->bvh.nodes[0].bmin[k]; scene->root_bounds[3 + k] = scene->bvh.nodes[0].bmax[k]; }
    scene->built_host = true;
    return PB2_OK;
}

int pb2_scene_build_bvh_host(pb2_scene* scene, int max_prims_in_node, int split_method) {
    if (!scene) return set_error(PB2_ERR_INVALID, "null scene");
    std::lock_guard<std::mutex> lock(scene->mu);
    return build_host_locked(scene, max_prims_in_node, split_method);
}

// SplitMethod::HLBVH: everything up to and including the traversal layout is produced on the device (bvh_hlbvh.cu).
static int build_device_locked(pb2_scene* scene, int max_prims_in_node, SceneView* v) {
    scene->built = false;
    scene->built_host = false;
    for (double& x : scene->build_ms) x = 0.0;
    scene->bvh = HostBVH();
    const uint64_t n_tris = scene->indices.size() / 3;
    DeviceBVH dv;
    char msg[256] = "";
    const int brc = build_hlbvh_gpu(scene->verts.data(), scene->verts.size() / 3, scene->indices.data(), n_tris, max_prims_in_node, &dv, msg,
                                    (int)sizeof msg, scene->build_ms);
    scene->d_pairs = dv.d_pairs;
    scene->d_quads = dv.d_quads;
    scene->d_tris = dv.d_tris;
    scene->d_slot_of_prim = dv.d_slot_of_prim;
    scene->d_lin_nodes = dv.d_nodes;
    scene->d_lin_prims = dv.d_ordered_prims;                           // (free_device() releases whatever was allocated)
    if (brc != 0) return set_error(PB2_ERR_CUDA, "%s", msg);
    if (dv.max_depth > kStackDepth)
        return set_error(PB2_ERR_LIMIT, "BVH depth %d exceeds the 64-entry traversal stack of BVHAccel::intersect", dv.max_depth);
    scene->n_nodes = dv.n_nodes;
    scene->n_prims = dv.n_tris;
    scene->tree_depth = dv.max_depth;
    for (int k = 0; k < 6; ++k) scene->root_bounds[k] = dv.root_bounds[k];
    if (n_tris) {
        launch_mark_degenerate(scene->d_tris, n_tris, scene->d_indices, scene->d_uvs, 0);
        PB2_CUDA(cudaGetLastError());
        PB2_CUDA(cudaDeviceSynchronize());
        v->quads = (const float4*)scene->d_quads;
        v->quad_root_ref = dv.quad_root_ref;
        v->pairs = (const float4*)scene->d_pairs;
        v->tris = (const float4*)scene->d_tris;
        v->slot_of_prim = (const uint32_t*)scene->d_slot_of_prim;
        v->root_ref = dv.root_ref;
        for (int k = 0; k < 3; ++k) { v->root_lo[k] = dv.root_bounds[k]; v->root_hi[k] = dv.root_bounds[3 + k]; }
    }
    scene->built_host = true;
    return PB2_OK;
}

int pb2_scene_build_bvh(pb2_scene* scene, int max_prims_in_node, int split_method) {
    if (!scene) return set_error(PB2_ERR_INVALID, "null scene");
    std::lock_guard<std::mutex> lock(scene->mu);
    int rc = check_build_args(max_prims_in_node, split_method);
    if (rc) return rc;
    scene->free_device();
    if (split_method == 1 && !scene->spheres.empty())
        return set_error(PB2_ERR_INVALID, "SplitMethod::HLBVH (the device build) takes triangles only: build a scene with analytic spheres with SAH / Middle / EqualCounts");
    const uint64_t n_mesh_tris = scene->indices.size() / 3;
    const uint64_t n_tris = scene->n_primitives();               // every primitive takes one slot of the leaf-order array
    PB2_CUDA(cudaGetDevice(&scene->device));
    SceneView v;
    memset(&v, 0, sizeof v);
    v.n_tris = (uint32_t)n_tris;
    PB2_CUDA(cudaMalloc(&scene->d_counters, pb2_scene::kCounters * sizeof(unsigned long long)));
    PB2_CUDA(cudaMemset(scene->d_counters, 0, pb2_scene::kCounters * sizeof(unsigned long long)));
    // mesh attributes first: the per-triangle "degenerate frame" flag follows the mesh's UVs
    if (n_mesh_tris && (!scene->normals.empty() || !scene->tangents.empty() || !scene->uvs.empty())) {
        auto up = [](void** d, const void* h, size_t bytes) -> cudaError_t {
            cudaError_t e = cudaMalloc(d, bytes);
            return e == cudaSuccess ? cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice) : e;
        };
        PB2_CUDA(up(&scene->d_indices, scene->indices.data(), scene->indices.size() * 4));
        if (!scene->normals.empty()) PB2_CUDA(up(&scene->d_normals, scene->normals.data(), scene->normals.size() * 4));
        if (!scene->tangents.empty()) PB2_CUDA(up(&scene->d_tangents, scene->tangents.data(), scene->tangents.size() * 4));
        if (!scene->uvs.empty()) PB2_CUDA(up(&scene->d_uvs, scene->uvs.data(), scene->uvs.size() * 4));
    }
    if (split_method == 1) {
        rc = build_device_locked(scene, max_prims_in_node, &v);
        if (rc != PB2_OK) return rc;
    } else {
        rc = build_host_locked(scene, max_prims_in_node, split_method);
        if (rc != PB2_OK) return rc;
        HostBVH& b = scene->bvh;
        if (n_tris) {
            PB2_CUDA(cudaMalloc(&scene->d_pairs, std::max<size_t>(64, b.pairs.size() * sizeof(PairNode))));
            PB2_CUDA(cudaMalloc(&scene->d_tris, b.tris.size() * sizeof(PackedTri)));
            PB2_CUDA(cudaMalloc(&scene->d_slot_of_prim, n_tris * 4));
            if (!b.pairs.empty()) PB2_CUDA(cudaMemcpy(scene->d_pairs, b.pairs.data(), b.pairs.size() * sizeof(PairNode), cudaMemcpyHostToDevice));
            PB2_CUDA(cudaMalloc(&scene->d_quads, std::max<size_t>(128, b.quads.size() * sizeof(QuadNode))));
            if (!b.quads.empty()) PB2_CUDA(cudaMemcpy(scene->d_quads, b.quads.data(), b.quads.size() * sizeof(QuadNode), cudaMemcpyHostToDevice));
            v.quads = (const float4*)scene->d_quads;
            v.quad_root_ref = b.quad_root_ref;
            PB2_CUDA(cudaMemcpy(scene->d_tris, b.tris.data(), b.tris.size() * sizeof(PackedTri), cudaMemcpyHostToDevice));
            launch_mark_degenerate(scene->d_tris, b.tris.size(), scene->d_indices, scene->d_uvs, 0);
            PB2_CUDA(cudaGetLastError());
            PB2_CUDA(cudaDeviceSynchronize());
            std::vector<uint32_t> slot(n_tris);
            for (uint64_t i = 0; i < n_tris; ++i) slot[b.ordered_prims[i]] = (uint32_t)i;
            PB2_CUDA(cudaMemcpy(scene->d_slot_of_prim, slot.data(), n_tris * 4, cudaMemcpyHostToDevice));
            v.pairs = (const float4*)scene->d_pairs;
            v.tris = (const float4*)scene->d_tris;
            v.slot_of_prim = (const uint32_t*)scene->d_slot_of_prim;
            v.root_ref = b.root_ref;
            for (int k = 0; k < 3; ++k) { v.root_lo[k] = b.root_bounds[k]; v.root_hi[k] = b.root_bounds[3 + k]; }
        }
        if (!scene->sphere_records.empty()) {
            PB2_CUDA(cudaMalloc(&scene->d_spheres, scene->sphere_records.size() * sizeof(DSphere)));
            PB2_CUDA(cudaMemcpy(scene->d_spheres, scene->sphere_records.data(), scene->sphere_records.size() * sizeof(DSphere), cudaMemcpyHostToDevice));
            v.spheres = (const float4*)scene->d_spheres;
        }
        // the device copies are authoritative from here on; drop the host-side device-layout mirrors
        std::vector<PairNode>().swap(b.pairs);
        std::vector<QuadNode>().swap(b.quads);
        std::vector<PackedTri>().swap(b.tris);
    }
    // shading scenes: the triangle records once more, in primitive order, each with its shading frame (k_shade rebuilds a vertex
    // from hit.prim directly)
    if (n_tris && !scene->materials.empty() && (n_mesh_tris == 0 || !scene->tri_material.empty())) {
        PB2_CUDA(cudaMalloc(&scene->d_tris_prim, n_tris * 2 * sizeof(PackedTri)));      // record + shading frame (k_tris_by_prim)
        launch_tris_by_prim(scene->d_tris, n_tris, scene->d_tris_prim, 0);
        PB2_CUDA(cudaGetLastError());
        PB2_CUDA(cudaDeviceSynchronize());
        v.tris_prim = (const float4*)scene->d_tris_prim;
    }
    scene->view = v;
    rc = upload_shading_tables(scene);
    if (rc != PB2_OK) return rc;
    scene->built = true;
    return PB2_OK;
}

int pb2_bvh_build_stats(const pb2_scene* scene, double ms[6]) {
    if (!scene || !ms) return set_error(PB2_ERR_INVALID, "null argument");
    for (int k = 0; k < 6; ++k) ms[k] = scene->build_ms[k];
    return PB2_OK;
}

int pb2_world_bound(const pb2_scene* scene, float out[6]) {
    if (!scene || !out) return set_error(PB2_ERR_INVALID, "null argument");
    if (!scene->built_host) return set_error(PB2_ERR_STATE, "pb2_scene_build_bvh has not been called");
    if (scene->n_nodes == 0) {               // Bounds3f::default(): bvh.rs:825
        out[0] = out[1] = out[2] = 3.402823466e+38f;
        out[3] = out[4] = out[5] = -3.402823466e+38f;
        return PB2_OK;
    }
    for (int k = 0; k < 6; ++k) out[k] = scene->root_bounds[k];
    return PB2_OK;
}

int pb2_bvh_info(const pb2_scene* scene, uint64_t* n_nodes, uint64_t* n_prims, int* max_depth) {
    if (!scene) return set_error(PB2_ERR_INVALID, "null scene");
    if (!scene->built_host) return set_error(PB2_ERR_STATE, "pb2_scene_build_bvh has not been called");
    if (n_nodes) *n_nodes = scene->n_nodes;
    if (n_prims) *n_prims = scene->n_prims;
    if (max_depth) *max_depth = scene->tree_depth;
    return PB2_OK;
}

int pb2_bvh_export(const pb2_scene* scene, void* nodes32, uint32_t* ordered_prims) {
    if (!scene) return set_error(PB2_ERR_INVALID, "null scene");
    if (!scene->built_host) return set_error(PB2_ERR_STATE, "pb2_scene_build_bvh has not been called");
    if (scene->d_lin_nodes) {                // device-built tree: the reference-layout arrays live on the device
        if (nodes32) PB2_CUDA(cudaMemcpy(nodes32, scene->d_lin_nodes, scene->n_nodes * sizeof(LinearNode), cudaMemcpyDeviceToHost));
        if (ordered_prims) PB2_CUDA(cudaMemcpy(ordered_prims, scene->d_lin_prims, scene->n_prims * 4, cudaMemcpyDeviceToHost));
        return PB2_OK;
    }
    if (nodes32) memcpy(nodes32, scene->bvh.nodes.data(), scene->bvh.nodes.size() * sizeof(LinearNode));
    if (ordered_prims) memcpy(ordered_prims, scene->bvh.ordered_prims.data(), scene->bvh.ordered_prims.size() * 4);
    return PB2_OK;
}

// ---- batched intersect ------------------------------------------------------------------------------------
static int ensure_stage(pb2_scene* s, size_t chunk) {
    Pipe& p = s->pipe;
    if (!p.h2d) PB2_CUDA(cudaStreamCreateWithFlags(&p.h2d, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i)
        if (!p.compute[i]) PB2_CUDA(cudaStreamCreateWithFlags(&p.compute[i], cudaStreamNonBlocking));
    if (!p.d2h) PB2_CUDA(cudaStreamCreateWithFlags(&p.d2h, cudaStreamNonBlocking));
    for (int i = 0; i < kStages; ++i) {
        Stage& st = p.slot[i];
        if (!st.in_ready) PB2_CUDA(cudaEventCreateWithFlags(&st.in_ready, cudaEventDisableTiming));
        if (!st.done) PB2_CUDA(cudaEventCreateWithFlags(&st.done, cudaEventDisableTiming));
        if (!st.drained) PB2_CUDA(cudaEventCreateWithFlags(&st.drained, cudaEventDisableTiming));
        if (st.cap < chunk) {
            if (p.pending) { PB2_CUDA(cudaStreamSynchronize(p.d2h)); p.pending = false; }     // batches still in flight use the old buffers
            if (st.d_in) cudaFree(st.d_in);
            if (st.d_out) cudaFree(st.d_out);
            if (st.d_aux) cudaFree(st.d_aux);
            st.d_in = st.d_out = st.d_aux = nullptr;
            PB2_CUDA(cudaMalloc(&st.d_in, chunk * 32));
            PB2_CUDA(cudaMalloc(&st.d_out, chunk * 16));
            PB2_CUDA(cudaMalloc(&st.d_aux, chunk * 4));
            st.cap = chunk;
        }
    }
    return PB2_OK;
}

static int check_ready(const pb2_scene* scene) {
    if (!scene) return set_error(PB2_ERR_INVALID, "null scene");
    if (!scene->built) return set_error(PB2_ERR_STATE, "pb2_scene_build_bvh has not been called");
    return PB2_OK;
}

// Host-buffer entry points stream the batch through the ring (api_internal.hpp: Pipe): the H2D copy of chunk k+1, the
// traversal of chunk k and the D2H copy of chunk k-1 run on different engines at the same time (pinned caller memory:
// pb2_host_alloc; pageable memory works but its copies are staged by the driver).  128 K rays per chunk keeps the pipeline
// fill (one H2D) and drain (one kernel + one D2H) short against the 8+ chunks of a megaray batch.  The H2D engine is the
// bottleneck (32 B in per ray against 16 B out and ~0.4 ns of traversal), so a call ends one kernel + one D2H after the last
// H2D copy.  Chunks can halve towards the end of the batch (down to kTailChunk rays) to shorten that drain; measured on B200
// (profiles/r01_tuning.md, session 3) every extra chunk costs more than the drain it saves, so the default tail equals the
// chunk.  Reading the rays straight from pinned host memory inside the kernel (zero-copy, one launch per call) was built and
// measured too: its reads alone run at 42 GB/s, no faster than this ring's H2D copies, and per-ray result writes over PCIe
// (one small TLP each) made the call 1.5x slower — the link, not the pipeline, bounds these entry points.
static size_t env_size(const char* name, size_t dflt) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    const long long x = atoll(v);
    return x > 0 ? (size_t)x : dflt;
}
static const size_t kChunk = env_size("PB2_PIPE_CHUNK", 1u << 17);          // tuning overrides for sweeps (tools/e2e_sweep.py)
static const size_t kTailChunk = std::min(kChunk, env_size("PB2_PIPE_TAIL", 1u << 17));
// _async batches pay the ring's fill and drain once per wait, not once per call, so larger chunks win there: C3 sets, three
// batches per wait, 1058 / 1385 / 1436 / 1293 / 1212 Mrays/s with 64 K / 128 K / 256 K / 512 K / 1 M rays per chunk.
static const size_t kChunkAsync = getenv("PB2_PIPE_CHUNK") ? kChunk : (size_t)(1u << 18);
static const size_t kTailAsync = std::min(kChunkAsync, env_size("PB2_PIPE_TAIL_ASYNC", kChunkAsync));

extern "C++" {
template <class Launch>
// wait = false (the _async entry points): return once the batch is enqueued; the chunk counter runs on over calls, so the
// next batch's first H2D copies start while this batch's last kernels and D2H copies are still draining.
static int run_pipe(pb2_scene* scene, const pb2_ray* rays, uint64_t n, size_t out_bytes, void* out, float* b0, bool wait, Launch&& launch) {
    const uint64_t chunk = wait ? kChunk : kChunkAsync, tail = wait ? kTailChunk : kTailAsync;
    int rc = ensure_stage(scene, std::min<uint64_t>(chunk, std::max<uint64_t>(n, 1)));
    if (rc) return rc;
    Pipe& p = scene->pipe;
    uint64_t m = 0;
    for (uint64_t off = 0; off < n; off += m) {
        const uint64_t left = n - off;
        m = std::min<uint64_t>(left, std::min<uint64_t>(chunk, std::max<uint64_t>(tail, (left + 1) / 2)));
        const uint64_t c = p.chunks++;
        Stage& st = p.slot[c % kStages];
        if (c >= (uint64_t)kStages) PB2_CUDA(cudaStreamWaitEvent(p.h2d, st.drained, 0));       // slot's previous chunk fully out
        PB2_CUDA(cudaMemcpyAsync(st.d_in, rays + off, m * 32, cudaMemcpyHostToDevice, p.h2d));
        PB2_CUDA(cudaEventRecord(st.in_ready, p.h2d));
        cudaStream_t cs = p.compute[c & 1];
        PB2_CUDA(cudaStreamWaitEvent(cs, st.in_ready, 0));
        launch(st, m, cs);
        PB2_CUDA(cudaGetLastError());
        PB2_CUDA(cudaEventRecord(st.done, cs));
        PB2_CUDA(cudaStreamWaitEvent(p.d2h, st.done, 0));
        PB2_CUDA(cudaMemcpyAsync((char*)out + off * out_bytes, st.d_out, m * out_bytes, cudaMemcpyDeviceToHost, p.d2h));
        if (b0) PB2_CUDA(cudaMemcpyAsync(b0 + off, st.d_aux, m * 4, cudaMemcpyDeviceToHost, p.d2h));
        PB2_CUDA(cudaEventRecord(st.drained, p.d2h));
    }
    if (!wait) {
        cudaEvent_t& ev = p.batch_done[p.batches % Pipe::kBatchRing];
        if (!ev) PB2_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        else if (p.batches >= (uint64_t)Pipe::kBatchRing) PB2_CUDA(cudaEventSynchronize(ev));     // (the batch that used it 64 batches ago)
        PB2_CUDA(cudaEventRecord(ev, p.d2h));
        ++p.batches;
        p.pending = true;
        return PB2_OK;
    }
    PB2_CUDA(cudaStreamSynchronize(p.d2h));               // in-order stream: earlier _async batches are out as well
    p.pending = false;
    return PB2_OK;
}
}  // extern "C++"

static int intersect_host(pb2_scene* scene, const pb2_ray* rays, uint64_t n, pb2_hit* hits, float* b0, bool wait) {
    int rc = check_ready(scene);
    if (rc) return rc;
    if (n && (!rays || !hits)) return set_error(PB2_ERR_INVALID, "null ray/hit buffer");
    std::lock_guard<std::mutex> lock(scene->mu);
    PB2_CUDA(cudaSetDevice(scene->device));
    return run_pipe(scene, rays, n, 16, hits, b0, wait, [&](Stage& st, uint64_t m, cudaStream_t s) {
        launch_closest_hit(scene->view, st.d_in, m, st.d_out, b0 ? st.d_aux : nullptr, scene->next_counter(), s);
    });
}
static int intersect_p_host(pb2_scene* scene, const pb2_ray* rays, uint64_t n, uint8_t* out, bool wait) {
    int rc = check_ready(scene);
    if (rc) return rc;
    if (n && (!rays || !out)) return set_error(PB2_ERR_INVALID, "null ray/output buffer");
    std::lock_guard<std::mutex> lock(scene->mu);
    PB2_CUDA(cudaSetDevice(scene->device));
    return run_pipe(scene, rays, n, 1, out, nullptr, wait, [&](Stage& st, uint64_t m, cudaStream_t s) {
        launch_any_hit(scene->view, st.d_in, m, st.d_out, scene->next_counter(), s);
    });
}

int pb2_intersect(pb2_scene* scene, const pb2_ray* rays, uint64_t n, pb2_hit* hits, float* b0) { return intersect_host(scene, rays, n, hits, b0, true); }
int pb2_intersect_p(pb2_scene* scene, const pb2_ray* rays, uint64_t n, uint8_t* out) { return intersect_p_host(scene, rays, n, out, true); }
int pb2_intersect_async(pb2_scene* scene, const pb2_ray* rays, uint64_t n, pb2_hit* hits, float* b0) { return intersect_host(scene, rays, n, hits, b0, false); }
int pb2_intersect_p_async(pb2_scene* scene, const pb2_ray* rays, uint64_t n, uint8_t* out) { return intersect_p_host(scene, rays, n, out, false); }

int pb2_scene_wait(pb2_scene* scene) {
    if (!scene) return set_error(PB2_ERR_INVALID, "null scene");
    std::lock_guard<std::mutex> lock(scene->mu);
    if (!scene->pipe.pending) return PB2_OK;
    PB2_CUDA(cudaSetDevice(scene->device));
    PB2_CUDA(cudaStreamSynchronize(scene->pipe.d2h));
    scene->pipe.pending = false;
    return PB2_OK;
}

int pb2_scene_wait_until(pb2_scene* scene, uint32_t in_flight) {
    if (!scene) return set_error(PB2_ERR_INVALID, "null scene");
    std::lock_guard<std::mutex> lock(scene->mu);
    Pipe& p = scene->pipe;
    if (!p.pending || p.batches <= (uint64_t)in_flight) return PB2_OK;
    PB2_CUDA(cudaSetDevice(scene->device));
    if (in_flight == 0) {
        PB2_CUDA(cudaStreamSynchronize(p.d2h));
        p.pending = false;
        return PB2_OK;
    }
    if (in_flight >= (uint32_t)Pipe::kBatchRing) return PB2_OK;      // (older batches were waited for when their events were reused)
    const uint64_t last = p.batches - 1 - in_flight;                   // the newest batch that must be complete; the d2h stream is in order
    PB2_CUDA(cudaEventSynchronize(p.batch_done[last % Pipe::kBatchRing]));
    return PB2_OK;
}

int pb2_intersect_device(pb2_scene* scene, const void* d_rays, uint64_t n, void* d_hits, void* d_b0, void* stream) {
    int rc = check_ready(scene);
    if (rc) return rc;
    if (n >> 32) return set_error(PB2_ERR_INVALID, "device batches are limited to 2^32 - 1 rays per call");
    launch_closest_hit(scene->view, d_rays, n, d_hits, d_b0, scene->next_counter(), (cudaStream_t)stream);
    PB2_CUDA(cudaGetLastError());
    return PB2_OK;
}

int pb2_intersect_p_device(pb2_scene* scene, const void* d_rays, uint64_t n, void* d_out, void* stream) {
    int rc = check_ready(scene);
    if (rc) return rc;
    if (n >> 32) return set_error(PB2_ERR_INVALID, "device batches are limited to 2^32 - 1 rays per call");
    launch_any_hit(scene->view, d_rays, n, d_out, scene->next_counter(), (cudaStream_t)stream);
    PB2_CUDA(cudaGetLastError());
    return PB2_OK;
}

// ---- camera -------------------------------------------------------------------------------------------------
int pb2_camera_matrices(const pb2_camera* cam, float r2c[16], float c2w[16]) {
    CameraView v;
    int rc = make_camera_view(cam, &v);
    if (rc) return rc;
    if (r2c) memcpy(r2c, v.raster_to_camera.m, 64);
    if (c2w) memcpy(c2w, v.camera_to_world.m, 64);
    return PB2_OK;
}

int pb2_camera_generate_rays(const pb2_camera* cam, const float* p_film, const float* p_lens, uint64_t n, pb2_ray* rays) {
    CameraView v;
    int rc = make_camera_view(cam, &v);
    if (rc) return rc;
    if (n == 0) return PB2_OK;
    if (!p_film || !rays) return set_error(PB2_ERR_INVALID, "null buffer");
    void *d_p = nullptr, *d_l = nullptr, *d_r = nullptr;
    PB2_CUDA(cudaMalloc(&d_p, n * 8));
    cudaError_t e = cudaMalloc(&d_r, n * 32);
    if (e == cudaSuccess && p_lens) e = cudaMalloc(&d_l, n * 8);
    if (e != cudaSuccess) { cudaFree(d_p); cudaFree(d_r); return cuda_fail(e, "cudaMalloc", __FILE__, __LINE__); }
    e = cudaMemcpy(d_p, p_film, n * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && p_lens) e = cudaMemcpy(d_l, p_lens, n * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) { launch_camera_rays(v, d_p, d_l, n, d_r, 0); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaMemcpy(rays, d_r, n * 32, cudaMemcpyDeviceToHost);
    cudaFree(d_p);
    cudaFree(d_l);
    cudaFree(d_r);
    if (e != cudaSuccess) return cuda_fail(e, "camera ray generation", __FILE__, __LINE__);
    return PB2_OK;
}

int pb2_camera_primary_rays_device(const pb2_camera* cam, void* d_rays, void* stream) {
    CameraView v;
    int rc = make_camera_view(cam, &v);
    if (rc) return rc;
    if (!d_rays) return set_error(PB2_ERR_INVALID, "null buffer");
    launch_camera_rays(v, nullptr, nullptr, (uint64_t)v.res_x * (uint64_t)v.res_y, d_rays, (cudaStream_t)stream);
    PB2_CUDA(cudaGetLastError());
    return PB2_OK;
}

// ---- secondary-ray builders -----------------------------------------------------------------------------------
int pb2_spawn_shadow_rays_device(pb2_scene* scene, const void* d_rays, const void* d_hits, uint64_t n,
                                 const float light_pos[3], void* d_out_rays, void* stream) {
    int rc = check_ready(scene);
    if (rc) return rc;
    if (!light_pos) return set_error(PB2_ERR_INVALID, "null light position");
    launch_spawn_shadow(scene->view, d_rays, d_hits, n, light_pos, d_out_rays, (cudaStream_t)stream);
    PB2_CUDA(cudaGetLastError());
    return PB2_OK;
}

int pb2_spawn_bounce_rays_device(pb2_scene* scene, const void* d_rays, const void* d_hits, uint64_t n,
                                 void* d_out_rays, void* stream) {
    int rc = check_ready(scene);
    if (rc) return rc;
    launch_spawn_bounce(scene->view, d_rays, d_hits, n, d_out_rays, (cudaStream_t)stream);
    PB2_CUDA(cudaGetLastError());
    return PB2_OK;
}

int pb2_spawn_shadow_bounce_rays_device(pb2_scene* scene, const void* d_rays, const void* d_hits, uint64_t n, const float light_pos[3],
                                        void* d_out_shadow_rays, void* d_out_bounce_rays, void* stream) {
    int rc = check_ready(scene);
    if (rc) return rc;
    if (!light_pos) return set_error(PB2_ERR_INVALID, "null light position");
    launch_spawn_shadow_bounce(scene->view, d_rays, d_hits, n, light_pos, d_out_shadow_rays, d_out_bounce_rays, (cudaStream_t)stream);
    PB2_CUDA(cudaGetLastError());
    return PB2_OK;
}

// ---- RNG parity hook ----------------------------------------------------------------------------------------
int pb2_rng_uniform_floats(uint64_t first_sequence, uint32_t n_sequences, uint32_t n_per, float* out) {
    const uint64_t total = (uint64_t)n_sequences * n_per;
    if (total == 0) return PB2_OK;
    if (!out) return set_error(PB2_ERR_INVALID, "null out");
    float* d = nullptr;
    PB2_CUDA(cudaMalloc(&d, total * 4));
    launch_rng_floats(first_sequence, n_sequences, n_per, d, 0);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(out, d, total * 4, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return cuda_fail(e, "rng kernel", __FILE__, __LINE__);
    return PB2_OK;
}

}  // extern "C"
