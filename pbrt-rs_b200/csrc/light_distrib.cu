// light_distrib.cu — SpatialLightDistribution (src/core/lightdistrib.rs:71-220) for the whole voxel grid, on the device.
//
// The reference builds a voxel's Distribution1D the first time a path vertex falls into it and parks it in a lock-free hash
// table (:160-219).  A voxel's distribution is a pure function of the voxel and the lights (128 Halton points, no RNG), so
// filling every voxel up front gives the distributions the lazy table would have held: k_spatial_contrib integrates one
// (voxel, light) pair per thread in the reference's sample order (:127-158), k_spatial_distrib applies the 0.1 % floor and
// Distribution1D::new (sampling.rs:76-98) per voxel.  k_shade then reads its vertex's voxel (spatial_voxel, wavefront.cuh).
#include "wavefront.cuh"

namespace pb2 {
namespace {

constexpr int kSpatialSamples = 128;                                     // lightdistrib.rs:126

// radical_inverse(b, i) for b = 0..4 and i < 128 (lightdistrib.rs:128-142), evaluated once on the host.
struct SpatialSamples {
    float v[5][kSpatialSamples];
};

// Light::sample_li at a bare point (no normal, no error bounds: lightdistrib.rs:133-140) without its VisibilityTester:
// point.rs:47-66, spot.rs:71-85, distant.rs:50-67, diffuse.rs:60-81 + shape.rs:38-53 + triangle.rs:330-348.  Same operations,
// in the same order, as the head of direct_lighting() in wavefront.cu.
__device__ __forceinline__ rgb3 light_sample_li(const DLight& light, const DSphere* spheres, vec3 p, float ul0, float ul1, float* pdf_out) {
    const rgb3 l_emit = mkc(light.l[0], light.l[1], light.l[2]);
    if (light.type == 1 && light.sphere >= 0) {                          // DiffuseAreaLight over a Sphere: sphere.rs:127-193 from a bare point
        vec3 ps, pe, ns;
        float pdf;
        sphere_sample2(spheres[light.sphere], p, mk(0.f, 0.f, 0.f), mk(0.f, 0.f, 0.f), ul0, ul1, &ps, &pe, &ns, &pdf);
        *pdf_out = pdf;
        if (pdf == 0.0f || len2(ps - p) == 0.0f) { *pdf_out = 0.0f; return gray(0.0f); }
        const vec3 wi = unit(ps - p);
        return (light.two_sided || dot3(ns, -wi) > 0.0f) ? l_emit : gray(0.0f);
    }
    if (light.type != 1) {                                               // delta lights: point, spot, distant
        *pdf_out = 1.0f;
        if (light.type == 3) return l_emit;
        const vec3 pl = mk(light.p[0], light.p[1], light.p[2]);
        const vec3 wi = unit(pl - p);
        if (light.type == 2) {                                           // SpotLight::falloff(-wi), spot.rs:51-63
            const vec3 w = -wi;
            const float cos_theta = (light.axis[0] * w.x + light.axis[1] * w.y) + light.axis[2] * w.z;
            float fall = 1.0f;
            if (cos_theta < light.cos_total_width) fall = 0.0f;
            else if (!(cos_theta >= light.cos_falloff_start)) {
                const float dl = (cos_theta - light.cos_total_width) / (light.cos_falloff_start - light.cos_total_width);
                fall = (dl * dl) * (dl * dl);
            }
            return l_emit * fall / len2(pl - p);
        }
        return l_emit / len2(pl - p);
    }
    const vec3 lp0 = mk(light.p0[0], light.p0[1], light.p0[2]), lp1 = mk(light.p1[0], light.p1[1], light.p1[2]),
               lp2 = mk(light.p2[0], light.p2[1], light.p2[2]);
    const float su0 = sqrtf(ul0);
    const float b0 = 1.0f - su0, b1 = ul1 * su0;                         // sampling.rs:275-278
    const float b2 = (1.0f - b0) - b1;
    const vec3 ps = (lp0 * b0 + lp1 * b1) + lp2 * b2;
    vec3 ns = unit(cross3(lp1 - lp0, lp2 - lp0));
    if (light.has_n)
        ns = face_toward(ns, (mk(light.n0[0], light.n0[1], light.n0[2]) * b0 + mk(light.n1[0], light.n1[1], light.n1[2]) * b1) +
                                 mk(light.n2[0], light.n2[1], light.n2[2]) * b2);
    float pdf = 1.0f / light.area;
    vec3 w = ps - p;
    if (len2(w) == 0.0f) pdf = 0.0f;
    else {
        w = unit(w);
        pdf = pdf * (len2(p - ps) / fabsf(dot3(ns, -w)));
        if (isinf(pdf)) pdf = 0.0f;
    }
    *pdf_out = pdf;
    if (pdf == 0.0f || len2(ps - p) == 0.0f) { *pdf_out = 0.0f; return gray(0.0f); }
    const vec3 wi = unit(ps - p);
    return (light.two_sided || dot3(ns, -wi) > 0.0f) ? l_emit : gray(0.0f);      // D55 FIX
}

// pbrt.rs:224-226 lerp per component (geometry.rs:454-458)
__device__ __forceinline__ vec3 bounds_lerp(vec3 lo, vec3 hi, vec3 t) {
    return mk((1.0f - t.x) * lo.x + t.x * hi.x, (1.0f - t.y) * lo.y + t.y * hi.y, (1.0f - t.z) * lo.z + t.z * hi.z);
}

// compute_distribution (lightdistrib.rs:107-158), the per-light sums: thread = (light, voxel).
__global__ void __launch_bounds__(128) k_spatial_contrib(SpatialView g, const DLight* __restrict__ lights, int n_lights, SpatialSamples smp,
                                                         float* __restrict__ func, const DSphere* __restrict__ spheres) {
    const size_t n_vox = (size_t)g.nv[0] * g.nv[1] * g.nv[2];
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_vox * (size_t)n_lights) return;
    // light-major thread order: a warp integrates one light over 32 neighbouring voxels, so the light-type branches of
    // sample_li stay uniform (voxel-major ran 12 of 32 lanes with the four light kinds side by side)
    const int j = (int)(idx / n_vox);
    const size_t vox = idx - (size_t)j * n_vox;
    const int px = (int)(vox % (size_t)g.nv[0]), py = (int)((vox / (size_t)g.nv[0]) % (size_t)g.nv[1]),
              pz = (int)(vox / ((size_t)g.nv[0] * g.nv[1]));
    const vec3 lo = mk(g.lo[0], g.lo[1], g.lo[2]), hi = mk(g.hi[0], g.hi[1], g.hi[2]);
    const vec3 p0 = mk((float)px / (float)g.nv[0], (float)py / (float)g.nv[1], (float)pz / (float)g.nv[2]);
    const vec3 p1 = mk((float)(px + 1) / (float)g.nv[0], (float)(py + 1) / (float)g.nv[1], (float)(pz + 1) / (float)g.nv[2]);   // D63 FIX
    const vec3 c0 = bounds_lerp(lo, hi, p0), c1 = bounds_lerp(lo, hi, p1);
    const vec3 vlo = mk(fminf(c0.x, c1.x), fminf(c0.y, c1.y), fminf(c0.z, c1.z));       // Bounds3::from((p, p)), geometry.rs:549-559
    const vec3 vhi = mk(fmaxf(c0.x, c1.x), fmaxf(c0.y, c1.y), fmaxf(c0.z, c1.z));
    const DLight light = lights[j];
    float contrib = 0.0f;
    for (int i = 0; i < kSpatialSamples; ++i) {
        const vec3 po = bounds_lerp(vlo, vhi, mk(smp.v[0][i], smp.v[1][i], smp.v[2][i]));
        float pdf = 0.0f;
        const rgb3 li = light_sample_li(light, spheres, po, smp.v[3][i], smp.v[4][i], &pdf);
        if (pdf > 0.0f) contrib = contrib + luminance(li) / pdf;
    }
    func[vox * (size_t)n_lights + (size_t)j] = contrib;
}

// The rest of compute_distribution (:159-172) and Distribution1D::new (sampling.rs:76-98, D29 FIX): thread = voxel.
__global__ void __launch_bounds__(128) k_spatial_distrib(size_t n_vox, int n_lights, float* __restrict__ func, float* __restrict__ cdf,
                                                         float* __restrict__ func_int) {
    const size_t vox = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vox >= n_vox) return;
    float* f = func + vox * (size_t)n_lights;
    float* c = cdf + vox * (size_t)(n_lights + 1);
    float sum_contrib = 0.0f;
    for (int j = 0; j < n_lights; ++j) sum_contrib = sum_contrib + f[j];
    const float avg_contrib = sum_contrib / (float)((size_t)kSpatialSamples * (size_t)n_lights);
    const float min_contrib = avg_contrib > 0.0f ? 0.001f * avg_contrib : 1.0f;
    for (int j = 0; j < n_lights; ++j) f[j] = fmaxf(min_contrib, f[j]);
    c[0] = 0.0f;
    for (int j = 1; j < n_lights + 1; ++j) c[j] = c[j - 1] + f[j - 1] / (float)n_lights;
    const float fi = c[n_lights];
    func_int[vox] = fi;
    if (fi == 0.0f) for (int j = 1; j < n_lights + 1; ++j) c[j] = (float)j / (float)n_lights;
    else for (int j = 1; j < n_lights + 1; ++j) c[j] = c[j] / fi;
}

// lowdiscrepancy.rs:322-331 radical_inverse, bases 2, 3, 5, 7, 11 (pbrt-v3 digit loop: DESIGN.md §8, S1)
float radical_inverse_small(int base_index, uint64_t a) {
    static const uint64_t kBases[5] = {2, 3, 5, 7, 11};
    if (base_index == 0) {
        uint64_t r = 0;
        for (int i = 0; i < 64; ++i) r |= ((a >> i) & 1ull) << (63 - i);
        return (float)r * 5.4210108624275222e-20f;
    }
    const uint64_t base = kBases[base_index];
    const float inv_base = 1.0f / (float)base;
    uint64_t reversed = 0;
    float inv_base_n = 1.0f;
    while (a != 0) {
        const uint64_t next = a / base, digit = a - next * base;
        reversed = reversed * base + digit;
        inv_base_n *= inv_base;
        a = next;
    }
    return fminf((float)reversed * inv_base_n, PB2_ONE_MINUS_EPS);
}

}  // namespace

// SpatialLightDistribution::new (lightdistrib.rs:83-105): voxels per axis from the world bound, 64 along its longest axis.
void spatial_grid_extents(const float wb[6], int max_voxels, int nv[3]) {
    const vec3 diag = mk(wb[3] - wb[0], wb[4] - wb[1], wb[5] - wb[2]);
    const float b_max = comp(diag, max_dim(diag));
    for (int i = 0; i < 3; ++i) {
        const float r = roundf(comp(diag, i) / b_max * (float)max_voxels);       // f32::round: halves away from zero
        const int v = (std::isnan(r) || r <= 0.0f) ? 0 : (r >= 1.0e9f ? 1000000000 : (int)r);   // `as usize` saturates, NaN -> 0
        nv[i] = v < 1 ? 1 : v;
    }
}

// Fills func [n_vox][n_lights], cdf [n_vox][n_lights + 1] and func_int [n_vox] (device arrays) on stream st.
void spatial_distribution_build(const SpatialView& grid, const DLight* d_lights, int n_lights, float* d_func, float* d_cdf, float* d_func_int,
                                cudaStream_t st, const void* d_spheres) {
    SpatialSamples smp;
    for (int b = 0; b < 5; ++b)
        for (int i = 0; i < kSpatialSamples; ++i) smp.v[b][i] = radical_inverse_small(b, (uint64_t)i);
    const size_t n_vox = (size_t)grid.nv[0] * grid.nv[1] * grid.nv[2];
    const size_t n_pairs = n_vox * (size_t)n_lights;
    k_spatial_contrib<<<(unsigned)((n_pairs + 127) / 128), 128, 0, st>>>(grid, d_lights, n_lights, smp, d_func, (const DSphere*)d_spheres);
    k_spatial_distrib<<<(unsigned)((n_vox + 127) / 128), 128, 0, st>>>(n_vox, n_lights, d_func, d_cdf, d_func_int);
}

}  // namespace pb2
