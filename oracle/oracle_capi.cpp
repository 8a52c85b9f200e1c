// ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.  C entry points for ctypes (tests/, smoke(), bench cpu_baseline).
// PARITY UNPINNED — see oracle_core.hpp header.
#include "oracle_core.hpp"
#include "oracle_path.hpp"

#include <atomic>
#include <chrono>
#include <thread>

using namespace orc;

// The Sobol' generator matrices (constant data; tools/make_sobol_tables.py) embedded from the package's data directory.
__asm__(".section .rodata\n"
        ".balign 16\n"
        ".global orc_sobol_blob\n"
        "orc_sobol_blob:\n"
        ".incbin \"../pbrt-rs_b200/data/sobol_tables.bin\"\n"
        ".global orc_sobol_blob_end\n"
        "orc_sobol_blob_end:\n"
        ".previous\n");
extern "C" const unsigned char orc_sobol_blob[], orc_sobol_blob_end[];
namespace orc {
const SobolTables* sobol_tables_base() {
    static SobolTables t;
    static const bool ok = t.load(orc_sobol_blob, (size_t)(orc_sobol_blob_end - orc_sobol_blob));
    return ok ? &t : nullptr;
}
}  // namespace orc

extern "C" {

// ---- scalar helpers ------------------------------------------------------------------------
float orc_gamma(float n) { return gamma(n); }
float orc_next_float_up(float v) { return next_float_up(v); }
float orc_next_float_down(float v) { return next_float_down(v); }
float orc_slab_widen() { return 1.0f + 2.0f * gamma(3.0f); }

// bounds = {min.xyz, max.xyz}; ray = 8 floats {o, t_max, d, time}
int orc_slab_test(const float* bounds, const float* ray8, float* t_entry) {
    Bounds3 b{{bounds[0], bounds[1], bounds[2]}, {bounds[3], bounds[4], bounds[5]}};
    Ray r;
    std::memcpy(&r, ray8, 32);
    V3 inv{1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z};
    int neg[3] = {inv.x < 0.0f, inv.y < 0.0f, inv.z < 0.0f};
    return slab_test(b, r, inv, neg, t_entry) ? 1 : 0;
}

// tri = 9 floats; out = {b0,b1,b2,t}
int orc_triangle_test(const float* tri, const float* ray8, float* out4) {
    Ray r;
    std::memcpy(&r, ray8, 32);
    TriHit h = triangle_intersect_test({tri[0], tri[1], tri[2]}, {tri[3], tri[4], tri[5]}, {tri[6], tri[7], tri[8]}, r);
    out4[0] = h.b0; out4[1] = h.b1; out4[2] = h.b2; out4[3] = h.t;
    return h.hit ? 1 : 0;
}

// ---- PCG32 ----------------------------------------------------------------------------------
void orc_pcg32_u32(uint64_t sequence, uint64_t init_state, int n, uint32_t* out) {
    RNG r;
    r.set_sequence(sequence, init_state);
    for (int i = 0; i < n; ++i) out[i] = r.uniform_u32();
}
void orc_pcg32_float(uint64_t sequence, int n, float* out) {
    RNG r;
    r.set_sequence(sequence);
    for (int i = 0; i < n; ++i) out[i] = r.uniform_float();
}

// ---- BVH -------------------------------------------------------------------------------------
void* orc_bvh_build(const float* verts, uint64_t nv, const uint32_t* idx, uint64_t nt, int max_prims_in_node) {
    BVHAccel* b = new BVHAccel();
    b->build(verts, nv, idx, nt, max_prims_in_node);
    return b;
}
// split_method: 0 = SplitMethod::SAH, 1 = SplitMethod::HLBVH (bvh.rs:199-204)
void* orc_bvh_build2(const float* verts, uint64_t nv, const uint32_t* idx, uint64_t nt, int max_prims_in_node, int split_method) {
    BVHAccel* b = new BVHAccel();
    b->build(verts, nv, idx, nt, max_prims_in_node, split_method);
    return b;
}
void orc_bvh_free(void* h) { delete (BVHAccel*)h; }
uint64_t orc_bvh_num_nodes(void* h) { return ((BVHAccel*)h)->nodes.size(); }
uint64_t orc_bvh_num_prims(void* h) { return ((BVHAccel*)h)->ordered_prims.size(); }
int orc_bvh_max_depth(void* h) { return ((BVHAccel*)h)->max_depth_seen; }
// 32-byte nodes: {min.xyz, max.xyz, offset u32, n_prims u16, axis u8, pad u8}
void orc_bvh_get_nodes(void* h, void* out) {
    BVHAccel* b = (BVHAccel*)h;
    std::memcpy(out, b->nodes.data(), b->nodes.size() * sizeof(LinearBVHNode));
}
void orc_bvh_get_ordered_prims(void* h, uint32_t* out) {
    BVHAccel* b = (BVHAccel*)h;
    std::memcpy(out, b->ordered_prims.data(), b->ordered_prims.size() * 4);
}
void orc_bvh_world_bound(void* h, float* out6) {
    Bounds3 b = ((BVHAccel*)h)->world_bound();
    out6[0] = b.mn.x; out6[1] = b.mn.y; out6[2] = b.mn.z; out6[3] = b.mx.x; out6[4] = b.mx.y; out6[5] = b.mx.z;
}

static int clamp_threads(int t) {
    if (t <= 0) t = (int)std::thread::hardware_concurrency();
    return t < 1 ? 1 : t;
}

// Worker threads pull 4096-ray chunks from an atomic counter (mirrors one rayon task per work unit,
// parallel.rs:4-17).  counters2 = {nodes_tested, tris_tested} totals (may be null).  Returns seconds.
double orc_intersect(void* h, const float* rays8, uint64_t n, void* hits16, float* b0, uint64_t* counters2, int threads) {
    const BVHAccel* bvh = (const BVHAccel*)h;
    threads = clamp_threads(threads);
    std::atomic<uint64_t> next{0};
    std::atomic<uint64_t> cn{0}, ct{0};
    const uint64_t chunk = 4096;
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&]() {
        TraversalCounters c;
        for (;;) {
            uint64_t s = next.fetch_add(chunk);
            if (s >= n) break;
            uint64_t e = std::min(n, s + chunk);
            for (uint64_t i = s; i < e; ++i) {
                Ray r;
                std::memcpy(&r, rays8 + 8 * i, 32);
                bvh->intersect(r, (Hit*)hits16 + i, b0 ? b0 + i : nullptr, counters2 ? &c : nullptr);
            }
        }
        cn += c.nodes_tested;
        ct += c.tris_tested;
    };
    std::vector<std::thread> pool;
    for (int i = 1; i < threads; ++i) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (counters2) { counters2[0] = cn; counters2[1] = ct; }
    return dt;
}

double orc_intersect_p(void* h, const float* rays8, uint64_t n, uint8_t* out, uint64_t* counters2, int threads) {
    const BVHAccel* bvh = (const BVHAccel*)h;
    threads = clamp_threads(threads);
    std::atomic<uint64_t> next{0};
    std::atomic<uint64_t> cn{0}, ct{0};
    const uint64_t chunk = 4096;
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&]() {
        TraversalCounters c;
        for (;;) {
            uint64_t s = next.fetch_add(chunk);
            if (s >= n) break;
            uint64_t e = std::min(n, s + chunk);
            for (uint64_t i = s; i < e; ++i) {
                Ray r;
                std::memcpy(&r, rays8 + 8 * i, 32);
                out[i] = bvh->intersect_p(r, counters2 ? &c : nullptr) ? 1 : 0;
            }
        }
        cn += c.nodes_tested;
        ct += c.tris_tested;
    };
    std::vector<std::thread> pool;
    for (int i = 1; i < threads; ++i) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (counters2) { counters2[0] = cn; counters2[1] = ct; }
    return dt;
}

void orc_brute_force(void* h, const float* rays8, uint64_t n, void* hits16, int threads) {
    const BVHAccel* bvh = (const BVHAccel*)h;
    threads = clamp_threads(threads);
    std::atomic<uint64_t> next{0};
    auto work = [&]() {
        for (;;) {
            uint64_t i = next.fetch_add(1);
            if (i >= n) break;
            Ray r;
            std::memcpy(&r, rays8 + 8 * i, 32);
            bvh->brute_force(r, (Hit*)hits16 + i);
        }
    };
    std::vector<std::thread> pool;
    for (int i = 1; i < threads; ++i) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
}

// ---- secondary rays of the C3 workload -------------------------------------------------------------
static bool hit_interaction(const BVHAccel* bvh, const Hit& h, Float b0, Interaction* it) {
    if (h.prim_id == 0xFFFFFFFFu) return false;
    V3 p0, p1, p2;
    bvh->tri(h.prim_id, &p0, &p1, &p2);
    *it = triangle_interaction(p0, p1, p2, b0, h.b1, h.b2);
    return true;
}
static void dead_ray(float* out8) {
    const float r[8] = {0, 0, 0, -1.0f, 0, 0, 1.0f, 0};
    std::memcpy(out8, r, 32);
}
// interaction.rs:138-144 spawn_ray_to(light position) from every closest hit; misses -> t_max = -1
void orc_spawn_shadow_rays(void* h, const void* hits16, const float* b0, uint64_t n, const float* light3, float* out8) {
    const BVHAccel* bvh = (const BVHAccel*)h;
    for (uint64_t i = 0; i < n; ++i) {
        Interaction it;
        if (!hit_interaction(bvh, ((const Hit*)hits16)[i], b0[i], &it)) { dead_ray(out8 + 8 * i); continue; }
        Ray r = spawn_ray_to(it, V3{light3[0], light3[1], light3[2]});
        std::memcpy(out8 + 8 * i, &r, 32);
    }
}
// interaction.rs:132-135 spawn_ray(wi), wi cosine-distributed about the geometric normal facing the incoming side,
// (u0,u1) = first two floats of RNG::new(i)
void orc_spawn_bounce_rays(void* h, const float* rays8, const void* hits16, const float* b0, uint64_t n, float* out8) {
    const BVHAccel* bvh = (const BVHAccel*)h;
    for (uint64_t i = 0; i < n; ++i) {
        Interaction it;
        if (!hit_interaction(bvh, ((const Hit*)hits16)[i], b0[i], &it)) { dead_ray(out8 + 8 * i); continue; }
        V3 d_in{rays8[8 * i + 4], rays8[8 * i + 5], rays8[8 * i + 6]};
        V3 n_face = it.n;
        if (dot(n_face, d_in) > 0.0f) n_face = -n_face;
        RNG rng;
        rng.set_sequence(i);
        Float u0 = rng.uniform_float();
        Float u1 = rng.uniform_float();
        V3 l = cosine_sample_hemisphere(u0, u1);
        V3 s, t;
        coordinate_system(n_face, &s, &t);
        V3 wi = (s * l.x + t * l.y) + n_face * l.z;
        Ray r = spawn_ray(it, wi);
        std::memcpy(out8 + 8 * i, &r, 32);
    }
}

// ---- camera ----------------------------------------------------------------------------------
// cam9 = {pos, look, up}; out: raster_to_camera (16), camera_to_world (16)
void orc_camera_matrices(const float* cam9, float fov, int rx, int ry, float* r2c16, float* c2w16) {
    Camera c;
    c.init({cam9[0], cam9[1], cam9[2]}, {cam9[3], cam9[4], cam9[5]}, {cam9[6], cam9[7], cam9[8]}, fov, rx, ry);
    std::memcpy(r2c16, c.raster_to_camera.m, 64);
    std::memcpy(c2w16, c.camera_to_world.m, 64);
}
// One ray per pixel centre p_film = (x + 0.5, y + 0.5), row-major.
void orc_camera_primary_rays(const float* cam9, float fov, int rx, int ry, float* rays8) {
    Camera c;
    c.init({cam9[0], cam9[1], cam9[2]}, {cam9[3], cam9[4], cam9[5]}, {cam9[6], cam9[7], cam9[8]}, fov, rx, ry);
    for (int y = 0; y < ry; ++y)
        for (int x = 0; x < rx; ++x) {
            Ray r = c.generate_ray((float)x + 0.5f, (float)y + 0.5f);
            std::memcpy(rays8 + 8 * ((size_t)y * rx + x), &r, 32);
        }
}
// Arbitrary film points.
void orc_camera_rays(const float* cam9, float fov, int rx, int ry, const float* pfilm2, uint64_t n, float* rays8) {
    Camera c;
    c.init({cam9[0], cam9[1], cam9[2]}, {cam9[3], cam9[4], cam9[5]}, {cam9[6], cam9[7], cam9[8]}, fov, rx, ry);
    for (uint64_t i = 0; i < n; ++i) {
        Ray r = c.generate_ray(pfilm2[2 * i], pfilm2[2 * i + 1]);
        std::memcpy(rays8 + 8 * i, &r, 32);
    }
}
// Thin lens (perspective.rs:101-107): film points + lens samples.
void orc_camera_rays_lens(const float* cam9, float fov, int rx, int ry, float lens_radius, float focal_distance, const float* pfilm2,
                          const float* plens2, uint64_t n, float* rays8) {
    Camera c;
    c.init({cam9[0], cam9[1], cam9[2]}, {cam9[3], cam9[4], cam9[5]}, {cam9[6], cam9[7], cam9[8]}, fov, rx, ry);
    c.lens_radius = lens_radius;
    c.focal_distance = focal_distance;
    for (uint64_t i = 0; i < n; ++i) {
        Ray r = c.generate_ray(pfilm2[2 * i], pfilm2[2 * i + 1], plens2[2 * i], plens2[2 * i + 1]);
        std::memcpy(rays8 + 8 * i, &r, 32);
    }
}

}  // extern "C"

#include "oracle_path_capi.inc"
