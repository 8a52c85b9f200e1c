//! Rust side of the drop-in for lazytiger/pbrt-rs: the complete `extern "C"` binding of include/pbrt_b200.h and the types
//! that stand where the crate's own stand —
//!
//! * [`B200Accel`]          implements `Primitive` (src/core/primitive.rs:17-30) in place of `BVHAccel`
//!                          (src/accelerators/bvh.rs:216-271, 819-953);
//! * [`B200PathIntegrator`] implements `Integrator` (src/core/integrator.rs:29-42) in place of `PathIntegrator`
//!                          (src/integrators/path.rs:31-63) over `SamplerIntegrator::render` (integrator.rs:399-480);
//! * [`B200Film`]           is the device film (`Film::new` film.rs:31-75) whose accumulators are handed to the crate's
//!                          `Film` through `Film::set_image` (film.rs:125-135).
//!
//! NOT COMPILED in the build image (no rustc / cargo there, and the reference itself needs a 2021 nightly plus un-vendored
//! crates): this is complete source for a maintainer, checked here only mechanically — tests/test_host_side.py holds the
//! extern block to the header symbol for symbol and the `abi_size!` lines to the sizes of the C structs.  The interface that
//! is exercised on hardware is the identical C ABI through the Python ctypes binding (pbrt-rs_b200/__init__.py).
#![allow(non_camel_case_types)]
#![allow(clippy::too_many_arguments)]

use std::any::Any;
use std::ffi::{CStr, CString};
use std::marker::PhantomData;
use std::os::raw::{c_char, c_int, c_void};
use std::sync::Arc;

use pbrt::core::{
    camera::CameraDt,
    film::Film,
    geometry::{Bounds2i, Bounds3f, Normal3f, Point2f, Point3f, Ray, Vector3f},
    integrator::Integrator,
    interaction::SurfaceInteraction,
    light::LightDt,
    material::{MaterialDt, TransportMode},
    medium::MediumInterface,
    pbrt::{gamma, Float},
    primitive::{Primitive, PrimitiveDt},
    scene::Scene,
    spectrum::{Spectrum, SpectrumType},
};

// ------------------------------------------------------------------------------------------------------------------
// POD structs of include/pbrt_b200.h.  abi_size!(T, n) is a compile-time check against the C sizeof, and the same numbers are
// compared with the header by tests/test_host_side.py::test_rust_shim_matches_header.
// ------------------------------------------------------------------------------------------------------------------
macro_rules! abi_size {
    ($t:ty, $n:expr) => {
        const _: [(); $n] = [(); std::mem::size_of::<$t>()];
    };
}

/// `Ray {o, d, t_max, time}` (src/core/geometry.rs:756-763).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct pb2_ray {
    pub o: [f32; 3],
    pub t_max: f32,
    pub d: [f32; 3],
    pub time: f32,
}
abi_size!(pb2_ray, 32);

/// What `Primitive::intersect` leaves behind: primitive index, the shrunk `ray.t_max`, barycentrics b1 / b2.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct pb2_hit {
    pub prim_id: u32,
    pub t: f32,
    pub b1: f32,
    pub b2: f32,
}
abi_size!(pb2_hit, 16);

pub const PB2_MAT_MATTE: i32 = 0;
pub const PB2_MAT_PLASTIC: i32 = 1;
pub const PB2_MAT_GLASS: i32 = 2;
pub const PB2_MAT_MIRROR: i32 = 3;
pub const PB2_MAT_METAL: i32 = 4;
pub const PB2_MAT_SUBSTRATE: i32 = 5;

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct pb2_material {
    pub ty: i32,
    pub kd: [f32; 3],
    pub ks: [f32; 3],
    pub roughness: f32,
    pub remap_roughness: i32,
    pub kr: [f32; 3],
    pub kt: [f32; 3],
    pub eta: f32,
    /// matte: Oren-Nayar sigma in degrees, 0 = Lambertian
    pub sigma: f32,
    pub metal_eta: [f32; 3],
    pub metal_k: [f32; 3],
}
abi_size!(pb2_material, 92);

pub const PB2_LIGHT_POINT: i32 = 0;
pub const PB2_LIGHT_AREA: i32 = 1;
pub const PB2_LIGHT_SPOT: i32 = 2;
pub const PB2_LIGHT_DISTANT: i32 = 3;

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct pb2_light {
    pub ty: i32,
    pub p: [f32; 3],
    pub i: [f32; 3],
    pub prim_id: u32,
    pub two_sided: i32,
    /// spot: row 2 of world_to_light's 3x3 block (spot.rs:51-53); distant: the direction towards the light
    pub axis: [f32; 3],
    pub total_width: f32,
    pub falloff_start: f32,
}
abi_size!(pb2_light, 56);

/// `Sphere::new` (src/shapes/sphere.rs:229-248) + the material of its `GeometricPrimitive`; `object_to_world` row-major, affine.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct pb2_sphere {
    pub object_to_world: [f32; 16],
    pub radius: f32,
    pub z_min: f32,
    pub z_max: f32,
    /// degrees
    pub phi_max: f32,
    pub reverse_orientation: i32,
    pub material: u32,
}
abi_size!(pb2_sphere, 88);

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct pb2_camera {
    pub pos: [f32; 3],
    pub look: [f32; 3],
    pub up: [f32; 3],
    pub fov: f32,
    pub res_x: i32,
    pub res_y: i32,
    /// 0 = pinhole
    pub lens_radius: f32,
    pub focal_distance: f32,
}
abi_size!(pb2_camera, 56);

pub const PB2_FILTER_BOX: i32 = 0;
pub const PB2_FILTER_GAUSSIAN: i32 = 1;
pub const PB2_FILTER_TRIANGLE: i32 = 2;
pub const PB2_FILTER_MITCHELL: i32 = 3;
pub const PB2_FILTER_SINC: i32 = 4;

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct pb2_film_desc {
    pub res_x: i32,
    pub res_y: i32,
    pub filter: i32,
    pub radius_x: f32,
    pub radius_y: f32,
    pub gaussian_alpha: f32,
    pub mitchell_b: f32,
    pub mitchell_c: f32,
    pub sinc_tau: f32,
    /// {min.x, min.y, max.x, max.y}; all zero = the whole image
    pub crop_window: [f32; 4],
    /// <= 0 = infinity
    pub max_sample_luminance: f32,
}
abi_size!(pb2_film_desc, 56);

pub const PB2_LIGHTS_UNIFORM: i32 = 0;
pub const PB2_LIGHTS_POWER: i32 = 1;
pub const PB2_LIGHTS_SPATIAL: i32 = 2;
pub const PB2_SAMPLER_RANDOM: i32 = 0;
pub const PB2_SAMPLER_HALTON: i32 = 1;
pub const PB2_SAMPLER_STRATIFIED: i32 = 2;
pub const PB2_SAMPLER_ZEROTWO: i32 = 3;
pub const PB2_SAMPLER_SOBOL: i32 = 4;

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct pb2_path_desc {
    pub max_depth: i32,
    pub rr_threshold: f32,
    pub light_strategy: i32,
    pub spp: i32,
    pub sample_begin: i32,
    pub sample_end: i32,
    pub sampler: i32,
    pub n_sampled_dimensions: i32,
    pub x_samples: i32,
    pub y_samples: i32,
    pub jitter: i32,
    /// PB2_INTEGRATOR_*
    pub integrator: i32,
}
abi_size!(pb2_path_desc, 48);
pub const PB2_INTEGRATOR_PATH: i32 = 0;
pub const PB2_INTEGRATOR_VOLPATH: i32 = 1;

/// `HomogeneousMedium::new(sigma_a, sigma_s, g)` (src/media/homogeneous.rs:20-28).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct pb2_medium {
    pub sigma_a: [f32; 3],
    pub sigma_s: [f32; 3],
    pub g: f32,
}
abi_size!(pb2_medium, 28);
/// `tri_material` / `pb2_sphere::material` of a surface without a material (a medium interface only).
pub const PB2_NO_MATERIAL: u32 = 0xFFFF_FFFF;

pub enum pb2_scene {}
pub enum pb2_film {}

pub const PB2_OK: c_int = 0;
pub const PB2_ERR_INVALID: c_int = -1;
pub const PB2_ERR_CUDA: c_int = -2;
pub const PB2_ERR_STATE: c_int = -3;
pub const PB2_ERR_NCCL: c_int = -4;
pub const PB2_ERR_LIMIT: c_int = -5;
pub const PB2_MISS: u32 = 0xFFFF_FFFF;

// ------------------------------------------------------------------------------------------------------------------
// The whole header, in the header's order.
// ------------------------------------------------------------------------------------------------------------------
extern "C" {
    // runtime
    pub fn pb2_init(device: c_int) -> c_int;
    pub fn pb2_shutdown() -> c_int;
    pub fn pb2_last_error() -> *const c_char;
    pub fn pb2_device_count(out: *mut c_int) -> c_int;
    pub fn pb2_host_alloc(bytes: u64, out: *mut *mut c_void) -> c_int;
    pub fn pb2_host_free(p: *mut c_void) -> c_int;
    pub fn pb2_device_alloc(bytes: u64, out: *mut *mut c_void) -> c_int;
    pub fn pb2_device_free(p: *mut c_void) -> c_int;
    pub fn pb2_memcpy_h2d(dst_device: *mut c_void, src_host: *const c_void, bytes: u64) -> c_int;
    pub fn pb2_memcpy_d2h(dst_host: *mut c_void, src_device: *const c_void, bytes: u64) -> c_int;
    pub fn pb2_device_synchronize() -> c_int;
    pub fn pb2_set_trace_tuning(refill_below: c_int, node_quorum: c_int, leaf_quorum: c_int, prefetch: c_int) -> c_int;
    // scene + BVHAccel
    pub fn pb2_scene_create(verts: *const f32, n_verts: u64, indices: *const u32, n_tris: u64, tri_material: *const u32,
                            mats: *const pb2_material, n_mats: u32, lights: *const pb2_light, n_lights: u32,
                            out: *mut *mut pb2_scene) -> c_int;
    pub fn pb2_scene_set_shading_geometry(scene: *mut pb2_scene, normals: *const f32, tangents: *const f32, uvs: *const f32) -> c_int;
    pub fn pb2_scene_add_spheres(scene: *mut pb2_scene, spheres: *const pb2_sphere, n: u32) -> c_int;
    pub fn pb2_scene_set_media(scene: *mut pb2_scene, media: *const pb2_medium, n_media: u32, prim_inside: *const i32, prim_outside: *const i32,
                               camera_medium: i32) -> c_int;
    pub fn pb2_scene_destroy(scene: *mut pb2_scene) -> c_int;
    pub fn pb2_scene_build_bvh(scene: *mut pb2_scene, max_prims_in_node: c_int, split_method: c_int) -> c_int;
    pub fn pb2_scene_build_bvh_host(scene: *mut pb2_scene, max_prims_in_node: c_int, split_method: c_int) -> c_int;
    pub fn pb2_bvh_build_stats(scene: *const pb2_scene, ms: *mut f64) -> c_int;
    pub fn pb2_world_bound(scene: *const pb2_scene, out: *mut f32) -> c_int;
    pub fn pb2_bvh_info(scene: *const pb2_scene, n_nodes: *mut u64, n_prims: *mut u64, max_depth: *mut c_int) -> c_int;
    pub fn pb2_bvh_export(scene: *const pb2_scene, nodes32: *mut c_void, ordered_prims: *mut u32) -> c_int;
    // Primitive::intersect / intersect_p, batched
    pub fn pb2_intersect(scene: *mut pb2_scene, rays: *const pb2_ray, n: u64, hits: *mut pb2_hit, b0: *mut f32) -> c_int;
    pub fn pb2_intersect_p(scene: *mut pb2_scene, rays: *const pb2_ray, n: u64, out: *mut u8) -> c_int;
    pub fn pb2_intersect_async(scene: *mut pb2_scene, rays: *const pb2_ray, n: u64, hits: *mut pb2_hit, b0: *mut f32) -> c_int;
    pub fn pb2_intersect_p_async(scene: *mut pb2_scene, rays: *const pb2_ray, n: u64, out: *mut u8) -> c_int;
    pub fn pb2_scene_wait(scene: *mut pb2_scene) -> c_int;
    pub fn pb2_scene_wait_until(scene: *mut pb2_scene, in_flight: u32) -> c_int;
    pub fn pb2_intersect_device(scene: *mut pb2_scene, d_rays: *const c_void, n: u64, d_hits: *mut c_void, d_b0: *mut c_void,
                                stream: *mut c_void) -> c_int;
    pub fn pb2_intersect_p_device(scene: *mut pb2_scene, d_rays: *const c_void, n: u64, d_out: *mut c_void, stream: *mut c_void) -> c_int;
    // Camera::generate_ray
    pub fn pb2_camera_generate_rays(cam: *const pb2_camera, p_film: *const f32, p_lens: *const f32, n: u64, rays: *mut pb2_ray) -> c_int;
    pub fn pb2_camera_primary_rays_device(cam: *const pb2_camera, d_rays: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn pb2_camera_matrices(cam: *const pb2_camera, r2c: *mut f32, c2w: *mut f32) -> c_int;
    // secondary-ray builders of the C3 workload
    pub fn pb2_spawn_shadow_rays_device(scene: *mut pb2_scene, d_rays: *const c_void, d_hits: *const c_void, n: u64,
                                        light_pos: *const f32, d_out_rays: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn pb2_spawn_bounce_rays_device(scene: *mut pb2_scene, d_rays: *const c_void, d_hits: *const c_void, n: u64,
                                        d_out_rays: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn pb2_spawn_shadow_bounce_rays_device(scene: *mut pb2_scene, d_rays: *const c_void, d_hits: *const c_void, n: u64,
                                               light_pos: *const f32, d_out_shadow_rays: *mut c_void, d_out_bounce_rays: *mut c_void,
                                               stream: *mut c_void) -> c_int;
    // RNG parity hook
    pub fn pb2_rng_uniform_floats(first_sequence: u64, n_sequences: u32, n_per: u32, out: *mut f32) -> c_int;
    // Film
    pub fn pb2_film_create(desc: *const pb2_film_desc, out: *mut *mut pb2_film) -> c_int;
    pub fn pb2_film_destroy(film: *mut pb2_film) -> c_int;
    pub fn pb2_film_clear(film: *mut pb2_film) -> c_int;
    pub fn pb2_film_add_samples(film: *mut pb2_film, p_film: *const f32, l_rgb: *const f32, weight: *const f32, n: u64) -> c_int;
    pub fn pb2_film_read_xyzw(film: *mut pb2_film, out: *mut f32) -> c_int;
    pub fn pb2_film_resolve_rgb(film: *mut pb2_film, scale: f32, rgb: *mut f32) -> c_int;
    pub fn pb2_film_add_splats(film: *mut pb2_film, p_film: *const f32, v_rgb: *const f32, n: u64) -> c_int;
    pub fn pb2_film_set_image(film: *mut pb2_film, rgb: *const f32) -> c_int;
    pub fn pb2_film_resolve_rgb_splat(film: *mut pb2_film, scale: f32, splat_scale: f32, rgb: *mut f32) -> c_int;
    pub fn pb2_film_device_ptr(film: *mut pb2_film, d_xyzw: *mut *mut c_void, n_floats: *mut u64) -> c_int;
    pub fn pb2_film_bounds(film: *const pb2_film, pixel_bounds: *mut i32, sample_bounds: *mut i32) -> c_int;
    pub fn pb2_film_write_image(film: *mut pb2_film, filename: *const c_char, scale: f32) -> c_int;
    // Integrator::render / PathIntegrator::li
    pub fn pb2_render_path(scene: *mut pb2_scene, cam: *const pb2_camera, path: *const pb2_path_desc, film: *mut pb2_film,
                           stream: *mut c_void) -> c_int;
    pub fn pb2_path_li(scene: *mut pb2_scene, cam: *const pb2_camera, path: *const pb2_path_desc, pixel_xy: *const u32,
                       sample_index: *const u32, n: u64, l_rgb: *mut f32, p_film: *mut f32) -> c_int;
    pub fn pb2_spatial_light_distribution(scene: *mut pb2_scene, n_voxels: *mut i32, func: *mut f32, cdf: *mut f32, func_int: *mut f32) -> c_int;
    pub fn pb2_render_counters(scene: *mut pb2_scene, out: *mut u64) -> c_int;
    // multi-GPU film reduce
    pub fn pb2_nccl_unique_id(id: *mut c_char) -> c_int;
    pub fn pb2_nccl_init(id: *const c_char, rank: c_int, n_ranks: c_int) -> c_int;
    pub fn pb2_nccl_shutdown() -> c_int;
    pub fn pb2_film_reduce(film: *mut pb2_film, root: c_int, stream: *mut c_void) -> c_int;
}

/// The reference's own error convention is `panic!` / `unwrap` / `unimplemented!` (no `Result` on this path), so a failing
/// status becomes a panic carrying `pb2_last_error()`.
fn check(rc: c_int) {
    if rc != PB2_OK {
        let msg = unsafe { CStr::from_ptr(pb2_last_error()) }.to_string_lossy().into_owned();
        panic!("pbrt_b200 error {}: {}", rc, msg);
    }
}

fn v3(a: &[f32]) -> Vector3f {
    Vector3f::new(a[0], a[1], a[2])
}
fn p3(a: &[f32]) -> Point3f {
    Point3f::new(a[0], a[1], a[2])
}

impl From<&Ray> for pb2_ray {
    fn from(r: &Ray) -> Self {
        pb2_ray { o: [r.o.x, r.o.y, r.o.z], t_max: r.t_max, d: [r.d.x, r.d.y, r.d.z], time: r.time }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Pinned host memory for the asynchronous entry points.
// ------------------------------------------------------------------------------------------------------------------
/// `n` elements of `T` in page-locked memory (`pb2_host_alloc`), zero-initialised by the library.
pub struct PinnedBuf<T: Copy> {
    ptr: *mut T,
    len: usize,
    _t: PhantomData<T>,
}
unsafe impl<T: Copy + Send> Send for PinnedBuf<T> {}

impl<T: Copy> PinnedBuf<T> {
    pub fn new(len: usize) -> Self {
        let mut p: *mut c_void = std::ptr::null_mut();
        unsafe { check(pb2_host_alloc((len.max(1) * std::mem::size_of::<T>()) as u64, &mut p)) };
        PinnedBuf { ptr: p as *mut T, len, _t: PhantomData }
    }
    pub fn as_slice(&self) -> &[T] {
        unsafe { std::slice::from_raw_parts(self.ptr, self.len) }
    }
    pub fn as_mut_slice(&mut self) -> &mut [T] {
        unsafe { std::slice::from_raw_parts_mut(self.ptr, self.len) }
    }
    pub fn len(&self) -> usize {
        self.len
    }
    pub fn is_empty(&self) -> bool {
        self.len == 0
    }
}
impl<T: Copy> Drop for PinnedBuf<T> {
    fn drop(&mut self) {
        unsafe { pb2_host_free(self.ptr as *mut c_void) };
    }
}

/// A closest-hit or any-hit batch in flight on the scene's copy / compute ring.  The borrow keeps the ray and result buffers
/// untouched until `B200Accel::wait` has returned, which is the contract `pb2_intersect_async` states.
pub struct PendingBatch<'a> {
    _rays: &'a PinnedBuf<pb2_ray>,
    _out: PhantomData<&'a mut ()>,
}

// ------------------------------------------------------------------------------------------------------------------
// B200Accel — stands where BVHAccel stands: Scene::new(Arc::new(Box::new(B200Accel::new(..))), lights).
// ------------------------------------------------------------------------------------------------------------------
/// `SplitMethod` of src/accelerators/bvh.rs:199-204 in the numbering of `pb2_scene_build_bvh`.
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
#[repr(i32)]
pub enum SplitMethod {
    SAH = 0,
    HLBVH = 1,
    Middle = 2,
    EqualCounts = 3,
}

/// The mesh arrays a `TriangleMesh` holds (src/shapes/triangle.rs:17-26; its fields are private, so the caller hands over the
/// same arrays it built the mesh from).  World space, as `TriangleMesh::new` leaves them.
#[derive(Clone, Debug, Default)]
pub struct MeshData {
    pub p: Vec<f32>,               // 3 per vertex
    pub vertex_indices: Vec<u32>,  // 3 per triangle
    pub n: Option<Vec<f32>>,       // 3 per vertex
    pub s: Option<Vec<f32>>,       // 3 per vertex
    pub uv: Option<Vec<f32>>,      // 2 per vertex
    pub face_indices: Vec<i32>,    // empty, or 1 per triangle
    pub reverse_orientation: bool,
    pub transform_swaps_handedness: bool,
}

#[derive(Debug)]
pub struct B200Accel {
    scene: *mut pb2_scene,
    mesh: MeshData,
    /// The caller's `GeometricPrimitive`s in triangle order, when given: `intersect` points `si.primitive` at the one that was
    /// hit, so `SurfaceInteraction::compute_scattering_functions` / `le` reach its material and area light (interaction.rs:323-330).
    primitives: Vec<PrimitiveDt>,
    n_spheres: usize,
}
unsafe impl Send for B200Accel {}
unsafe impl Sync for B200Accel {}

impl B200Accel {
    /// `BVHAccel::new(prims, max_prims_in_node, SplitMethod::SAH)` (bvh.rs:216-271) for pure ray casting.
    pub fn new(verts: &[f32], indices: &[u32], max_prims_in_node: usize) -> Self {
        Self::with_split(verts, indices, max_prims_in_node, SplitMethod::SAH)
    }

    pub fn with_split(verts: &[f32], indices: &[u32], max_prims_in_node: usize, split_method: SplitMethod) -> Self {
        let mesh = MeshData { p: verts.to_vec(), vertex_indices: indices.to_vec(), ..Default::default() };
        Self::from_mesh(mesh, Vec::new(), None, &[], &[], max_prims_in_node, split_method)
    }

    /// The full constructor: mesh with optional shading geometry, the matching `GeometricPrimitive`s (may be empty), and the
    /// material / light tables the device path tracer shades from (`tri_material[i]` indexes `mats`; an area light names its
    /// emissive triangle in `pb2_light::prim_id`).
    pub fn from_mesh(mesh: MeshData, primitives: Vec<PrimitiveDt>, tri_material: Option<&[u32]>, mats: &[pb2_material],
                     lights: &[pb2_light], max_prims_in_node: usize, split_method: SplitMethod) -> Self {
        Self::from_mesh_and_spheres(mesh, &[], primitives, tri_material, mats, lights, max_prims_in_node, split_method)
    }

    /// The same with analytic `Sphere`s (src/shapes/sphere.rs) appended to the primitive list: sphere `k` has primitive id
    /// `n_triangles + k`; `primitives`, when given, lists the triangles' `GeometricPrimitive`s and then the spheres'.
    pub fn from_mesh_and_spheres(mesh: MeshData, spheres: &[pb2_sphere], primitives: Vec<PrimitiveDt>, tri_material: Option<&[u32]>,
                                 mats: &[pb2_material], lights: &[pb2_light], max_prims_in_node: usize, split_method: SplitMethod) -> Self {
        assert!(mesh.p.len() % 3 == 0 && mesh.vertex_indices.len() % 3 == 0);
        let n_tris = mesh.vertex_indices.len() / 3;
        assert!(primitives.is_empty() || primitives.len() == n_tris + spheres.len());
        if let Some(tm) = tri_material {
            assert_eq!(tm.len(), n_tris);
        }
        let mut scene = std::ptr::null_mut();
        unsafe {
            check(pb2_init(current_device()));
            check(pb2_scene_create(mesh.p.as_ptr(), (mesh.p.len() / 3) as u64, mesh.vertex_indices.as_ptr(), n_tris as u64,
                                   tri_material.map_or(std::ptr::null(), |t| t.as_ptr()),
                                   if mats.is_empty() { std::ptr::null() } else { mats.as_ptr() }, mats.len() as u32,
                                   if lights.is_empty() { std::ptr::null() } else { lights.as_ptr() }, lights.len() as u32,
                                   &mut scene));
            if mesh.n.is_some() || mesh.s.is_some() || mesh.uv.is_some() {
                check(pb2_scene_set_shading_geometry(scene,
                                                     mesh.n.as_ref().map_or(std::ptr::null(), |v| v.as_ptr()),
                                                     mesh.s.as_ref().map_or(std::ptr::null(), |v| v.as_ptr()),
                                                     mesh.uv.as_ref().map_or(std::ptr::null(), |v| v.as_ptr())));
            }
            if !spheres.is_empty() {
                check(pb2_scene_add_spheres(scene, spheres.as_ptr(), spheres.len() as u32));
            }
            check(pb2_scene_build_bvh(scene, max_prims_in_node as c_int, split_method as c_int));
        }
        B200Accel { scene, mesh, primitives, n_spheres: spheres.len() }
    }

    pub fn raw(&self) -> *mut pb2_scene {
        self.scene
    }

    /// Batched `Primitive::intersect`: shrinks `t_max` of every ray that hits and returns the hit records.
    pub fn intersect_many(&self, rays: &mut [pb2_ray]) -> Vec<pb2_hit> {
        let mut hits = vec![pb2_hit::default(); rays.len()];
        unsafe { check(pb2_intersect(self.scene, rays.as_ptr(), rays.len() as u64, hits.as_mut_ptr(), std::ptr::null_mut())) };
        for (r, h) in rays.iter_mut().zip(&hits) {
            if h.prim_id != PB2_MISS {
                r.t_max = h.t;
            }
        }
        hits
    }

    /// Batched `Primitive::intersect_p`.
    pub fn intersect_p_many(&self, rays: &[pb2_ray]) -> Vec<bool> {
        let mut out = vec![0u8; rays.len()];
        unsafe { check(pb2_intersect_p(self.scene, rays.as_ptr(), rays.len() as u64, out.as_mut_ptr())) };
        out.into_iter().map(|b| b != 0).collect()
    }

    /// Enqueue a closest-hit batch from pinned buffers and return at once; several batches enqueued back to back overlap on
    /// the scene's ring (`pb2_intersect_async`).  `hits` holds the results after `wait()`; `t_max` of the rays is not touched.
    pub fn intersect_many_async<'a>(&self, rays: &'a PinnedBuf<pb2_ray>, hits: &'a mut PinnedBuf<pb2_hit>) -> PendingBatch<'a> {
        assert!(hits.len() >= rays.len());
        unsafe { check(pb2_intersect_async(self.scene, rays.as_slice().as_ptr(), rays.len() as u64, hits.as_mut_slice().as_mut_ptr(),
                                           std::ptr::null_mut())) };
        PendingBatch { _rays: rays, _out: PhantomData }
    }

    pub fn intersect_p_many_async<'a>(&self, rays: &'a PinnedBuf<pb2_ray>, out: &'a mut PinnedBuf<u8>) -> PendingBatch<'a> {
        assert!(out.len() >= rays.len());
        unsafe { check(pb2_intersect_p_async(self.scene, rays.as_slice().as_ptr(), rays.len() as u64, out.as_mut_slice().as_mut_ptr())) };
        PendingBatch { _rays: rays, _out: PhantomData }
    }

    /// Blocks until every batch enqueued on this scene has its results in the host buffers (`pb2_scene_wait`).
    pub fn wait(&self, pending: Vec<PendingBatch<'_>>) {
        unsafe { check(pb2_scene_wait(self.scene)) };
        drop(pending);
    }

    /// Blocks until all but the `keep` newest batches are complete (`pb2_scene_wait_until`) and hands those back: a producer of
    /// batches (one per tile) enqueues ahead and retires the older ones, so the copy ring never drains between batches.
    pub fn wait_until<'a>(&self, mut pending: Vec<PendingBatch<'a>>, keep: usize) -> Vec<PendingBatch<'a>> {
        unsafe { check(pb2_scene_wait_until(self.scene, keep as u32)) };
        let done = pending.len().saturating_sub(keep);
        pending.drain(..done);
        pending
    }

    /// `(n_nodes, n_prims, max_depth)` of the flattened tree (`pb2_bvh_info`).
    pub fn info(&self) -> (u64, u64, i32) {
        let (mut n, mut p, mut d) = (0u64, 0u64, 0 as c_int);
        unsafe { check(pb2_bvh_info(self.scene, &mut n, &mut p, &mut d)) };
        (n, p, d)
    }

    fn tri(&self, prim: usize) -> [usize; 3] {
        let i = &self.mesh.vertex_indices[3 * prim..3 * prim + 3];
        [i[0] as usize, i[1] as usize, i[2] as usize]
    }

    /// `Triangle::get_uvs` (triangle.rs:60-72).
    fn uvs(&self, v: &[usize; 3]) -> [Point2f; 3] {
        match &self.mesh.uv {
            Some(uv) => [Point2f::new(uv[2 * v[0]], uv[2 * v[0] + 1]), Point2f::new(uv[2 * v[1]], uv[2 * v[1] + 1]),
                         Point2f::new(uv[2 * v[2]], uv[2 * v[2] + 1])],
            None => [Point2f::new(0.0, 0.0), Point2f::new(1.0, 0.0), Point2f::new(1.0, 1.0)],
        }
    }

    /// The part of `Triangle::intersect` after `intersect_test` (triangle.rs:193-316): builds the `SurfaceInteraction` of a hit
    /// from the primitive index and the barycentrics the device returned (b0 = 1 - b1 - b2 is what the device computed too:
    /// `pb2_intersect`'s optional b0 output carries the exact value and `fill_interaction_b0` takes it).
    pub fn fill_interaction(&self, ray: &Ray, hit: &pb2_hit, b0: Float, si: &mut SurfaceInteraction) -> bool {
        let v = self.tri(hit.prim_id as usize);
        let (b1, b2) = (hit.b1, hit.b2);
        let p = &self.mesh.p;
        let (p0, p1, p2) = (p3(&p[3 * v[0]..]), p3(&p[3 * v[1]..]), p3(&p[3 * v[2]..]));
        let uv = self.uvs(&v);
        let duv02 = uv[0] - uv[2];
        let duv12 = uv[1] - uv[2];
        let dp02 = p0 - p2;
        let dp12 = p1 - p2;
        let determinant = duv02[0] * duv12[1] - duv02[1] * duv12[0];
        let degenerate_uv = determinant.abs() < 1e-8; // pbrt-v3; the reference omits abs(): FIX on both sides
        let mut dpdu = Vector3f::default();
        let mut dpdv = Vector3f::default();
        if !degenerate_uv {
            let inv_det = 1.0 / determinant;
            dpdu = (dp02 * duv12[1] - dp12 * duv02[1]) * inv_det;
            dpdv = (dp02 * -duv12[0] + dp12 * duv02[0]) * inv_det;
        }
        if degenerate_uv || dpdu.cross(&dpdv).length_squared() == 0.0 {
            let ng = (p2 - p0).cross(&(p1 - p0));
            if ng.length_squared() == 0.0 {
                return false; // the device marks these triangles (k_mark_degenerate) and never reports them; kept for symmetry
            }
            let (u, w) = ng.normalize().coordinate_system();
            dpdu = u;
            dpdv = w;
        }
        let x_abs_sum = (b0 * p0.x).abs() + (b1 * p1.x).abs() + (b2 * p2.x).abs();
        let y_abs_sum = (b0 * p0.y).abs() + (b1 * p1.y).abs() + (b2 * p2.y).abs();
        let z_abs_sum = (b0 * p0.z).abs() + (b1 * p1.z).abs() + (b2 * p2.z).abs();
        let p_error = Vector3f::new(x_abs_sum, y_abs_sum, z_abs_sum) * gamma(7.0);
        let p_hit = p0 * b0 + p1 * b1 + p2 * b2;
        let uv_hit = uv[0] * b0 + uv[1] * b1 + uv[2] * b2;
        let face_index = if self.mesh.face_indices.is_empty() { 0 } else { self.mesh.face_indices[hit.prim_id as usize] };
        *si = SurfaceInteraction::new(p_hit, p_error, uv_hit, -ray.d, dpdu, dpdv, Normal3f::default(), Normal3f::default(),
                                      ray.time, None, face_index);
        si.shading.n = dp02.cross(&dp12).normalize().into();
        si.n = si.shading.n;
        if self.mesh.reverse_orientation ^ self.mesh.transform_swaps_handedness {
            si.n = -si.n;
            si.shading.n = si.n;
        }
        if self.mesh.n.is_some() || self.mesh.s.is_some() {
            let at = |a: &Vec<f32>, i: usize| v3(&a[3 * i..]);
            let ns: Vector3f = match &self.mesh.n {
                Some(n) => {
                    let ns = at(n, v[0]) * b0 + at(n, v[1]) * b1 + at(n, v[2]) * b2;
                    if ns.length_squared() > 0.0 { ns.normalize() } else { si.n.into() }
                }
                None => si.n.into(),
            };
            let mut ss: Vector3f = match &self.mesh.s {
                Some(s) => {
                    let ss = at(s, v[0]) * b0 + at(s, v[1]) * b1 + at(s, v[2]) * b2;
                    if ss.length_squared() > 0.0 { ss.normalize() } else { dpdu.normalize() }
                }
                None => dpdu.normalize(),
            };
            let mut ts = ss.cross(&ns);
            if ts.length_squared() > 0.0 {
                ts = ts.normalize();
                ss = ts.cross(&ns);
            } else {
                let (a, b) = ns.coordinate_system();
                ss = a;
                ts = b;
            }
            let (dndu, dndv) = match &self.mesh.n {
                Some(n) => {
                    let dn1 = at(n, v[0]) - at(n, v[2]);
                    let dn2 = at(n, v[1]) - at(n, v[2]);
                    if degenerate_uv {
                        let dn = (at(n, v[2]) - at(n, v[0])).cross(&(at(n, v[1]) - at(n, v[0])));
                        if dn.length_squared() == 0.0 {
                            (Vector3f::default(), Vector3f::default())
                        } else {
                            let (dnu, dnv) = dn.coordinate_system();
                            (dnu.normalize(), dnv.normalize())
                        }
                    } else {
                        let inv_det = 1.0 / determinant;
                        ((dn1 * duv12[1] - dn2 * duv02[1]) * inv_det, (dn1 * -duv12[0] + dn2 * duv02[0]) * inv_det)
                    }
                }
                None => (Vector3f::default(), Vector3f::default()),
            };
            if self.mesh.reverse_orientation {
                ts = -ts;
            }
            si.set_shading_geometry(ss, ts, dndu, dndv, true);
        }
        true
    }
}

impl Drop for B200Accel {
    fn drop(&mut self) {
        unsafe { pb2_scene_destroy(self.scene) };
    }
}

/// The device this process renders on: `PB2_DEVICE` (one process per GPU sets it to its local rank), else 0.
fn current_device() -> c_int {
    std::env::var("PB2_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0)
}

impl Primitive for B200Accel {
    fn as_any(&self) -> &dyn Any {
        self
    }

    /// bvh.rs:819-826.
    fn world_bound(&self) -> Bounds3f {
        let mut b = [0f32; 6];
        unsafe { check(pb2_world_bound(self.scene, b.as_mut_ptr())) };
        Bounds3f::from((p3(&b[0..3]), p3(&b[3..6])))
    }

    /// bvh.rs:828-879 + primitive.rs:65-78 + triangle.rs:182-316 for ONE ray: a batch of one through `pb2_intersect` — correct,
    /// and as slow as a kernel launch per ray; tile workers should batch with `intersect_many` / `intersect_many_async`.
    fn intersect(&self, r: &mut Ray, si: &mut SurfaceInteraction) -> bool {
        let ray = pb2_ray::from(&*r);
        let mut hit = pb2_hit::default();
        let mut b0: f32 = 0.0;
        unsafe { check(pb2_intersect(self.scene, &ray, 1, &mut hit, &mut b0)) };
        if hit.prim_id == PB2_MISS {
            return false;
        }
        if hit.prim_id as usize >= self.mesh.vertex_indices.len() / 3 {
            // an analytic sphere: the device found the closest primitive; its `GeometricPrimitive(Sphere)` rebuilds the
            // interaction (sphere.rs:38-93) — one CPU sphere test, the same arithmetic the device ran
            debug_assert!((hit.prim_id as usize) < self.mesh.vertex_indices.len() / 3 + self.n_spheres);
            let p = self.primitives.get(hit.prim_id as usize).expect("pass the spheres' GeometricPrimitives to from_mesh_and_spheres");
            let found = p.intersect(r, si);
            if found {
                si.primitive = Some(p.clone());
            }
            return found;
        }
        if !self.fill_interaction(r, &hit, b0, si) {
            return false;
        }
        r.t_max = hit.t; // primitive.rs:70
        // primitive.rs:72-76: no medium transitions on this path (media are out of scope), so both sides are the ray's medium
        si.medium_interface = MediumInterface::new(r.medium.clone(), r.medium.clone());
        if let Some(p) = self.primitives.get(hit.prim_id as usize) {
            si.primitive = Some(p.clone()); // what `//si.primitive = Some(self); todo in upper calling` (primitive.rs:71) leaves undone
        }
        true
    }

    /// bvh.rs:881-932.
    fn intersect_p(&self, r: &Ray) -> bool {
        let ray = pb2_ray::from(r);
        let mut out = 0u8;
        unsafe { check(pb2_intersect_p(self.scene, &ray, 1, &mut out)) };
        out != 0
    }

    // bvh.rs:934-953: an aggregate answers none of these.
    fn get_area_light(&self) -> Option<LightDt> {
        unimplemented!("Aggregate does not support get_area_light method, use GeometricPrimitive instead")
    }
    fn get_material(&self) -> Option<MaterialDt> {
        unimplemented!("Aggregate does not support get_material method, use GeometricPrimitive instead")
    }
    fn compute_scattering_functions(&self, _si: &mut SurfaceInteraction, _mode: TransportMode, _allow_multiple_lobes: bool) {
        unimplemented!("Aggregate does not support compute_scattering_function method, use GeometricPrimitive instead")
    }
}

// ------------------------------------------------------------------------------------------------------------------
// B200Film — the device film.
// ------------------------------------------------------------------------------------------------------------------
pub struct B200Film {
    film: *mut pb2_film,
    desc: pb2_film_desc,
}
unsafe impl Send for B200Film {}

impl B200Film {
    /// `Film::new` (film.rs:31-75): resolution, crop window, filter and `max_sample_luminance`; `scale` is an argument of
    /// `resolve_rgb` / `write_image` here.
    pub fn new(desc: pb2_film_desc) -> Self {
        let mut film = std::ptr::null_mut();
        unsafe { check(pb2_film_create(&desc, &mut film)) };
        B200Film { film, desc }
    }

    /// A device film shaped like the crate's `Film`: its resolution, cropped pixel bounds and filter radius, with the filter kind
    /// and parameters named by the caller (the `Filter` trait exposes `radius()` only, filter.rs:10-15).
    pub fn for_film(host: &Film, filter: i32, gaussian_alpha: f32, mitchell_bc: (f32, f32), sinc_tau: f32, max_sample_luminance: f32) -> Self {
        let (rx, ry) = (host.full_resolution.x as f32, host.full_resolution.y as f32);
        let b = &host.cropped_pixel_bounds;
        let r = host.filter.radius();
        Self::new(pb2_film_desc {
            res_x: host.full_resolution.x, res_y: host.full_resolution.y, filter, radius_x: r.x, radius_y: r.y,
            gaussian_alpha, mitchell_b: mitchell_bc.0, mitchell_c: mitchell_bc.1, sinc_tau,
            crop_window: [b.min.x as f32 / rx, b.min.y as f32 / ry, b.max.x as f32 / rx, b.max.y as f32 / ry],
            max_sample_luminance,
        })
    }

    pub fn raw(&self) -> *mut pb2_film {
        self.film
    }
    pub fn desc(&self) -> &pb2_film_desc {
        &self.desc
    }
    pub fn clear(&mut self) {
        unsafe { check(pb2_film_clear(self.film)) };
    }

    /// `(cropped_pixel_bounds, sample_bounds)` as `{x0, y0, x1, y1}` (film.rs:41-50, 76-81).
    pub fn bounds(&self) -> ([i32; 4], [i32; 4]) {
        let (mut p, mut s) = ([0i32; 4], [0i32; 4]);
        unsafe { check(pb2_film_bounds(self.film, p.as_mut_ptr(), s.as_mut_ptr())) };
        (p, s)
    }
    pub fn n_pixels(&self) -> usize {
        let (p, _) = self.bounds();
        ((p[2] - p[0]) * (p[3] - p[1])) as usize
    }

    /// `FilmTile::add_sample` + `Film::merge_film_tile` for a batch of host samples (film.rs:252-295, 111-123).
    pub fn add_samples(&mut self, p_film: &[f32], l_rgb: &[f32], weight: &[f32]) {
        let n = weight.len();
        assert!(p_film.len() == 2 * n && l_rgb.len() == 3 * n);
        unsafe { check(pb2_film_add_samples(self.film, p_film.as_ptr(), l_rgb.as_ptr(), weight.as_ptr(), n as u64)) };
    }
    /// `Film::add_splat` (film.rs:137-151) for a batch.
    pub fn add_splats(&mut self, p_film: &[f32], v_rgb: &[f32]) {
        let n = p_film.len() / 2;
        assert!(v_rgb.len() == 3 * n);
        unsafe { check(pb2_film_add_splats(self.film, p_film.as_ptr(), v_rgb.as_ptr(), n as u64)) };
    }
    /// Raw accumulators `{X, Y, Z, filter_weight_sum}` per pixel of the cropped bounds, row-major.
    pub fn read_xyzw(&self) -> Vec<f32> {
        let mut out = vec![0f32; 4 * self.n_pixels()];
        unsafe { check(pb2_film_read_xyzw(self.film, out.as_mut_ptr())) };
        out
    }
    /// The pixel loop of `Film::write_image` (film.rs:153-178).
    pub fn resolve_rgb(&self, scale: f32, splat_scale: f32) -> Vec<f32> {
        let mut out = vec![0f32; 3 * self.n_pixels()];
        unsafe { check(pb2_film_resolve_rgb_splat(self.film, scale, splat_scale, out.as_mut_ptr())) };
        out
    }
    /// `.pfm` or `.ppm`.
    pub fn write_image(&self, filename: &str, scale: f32) {
        let c = CString::new(filename).expect("file name with an interior NUL");
        unsafe { check(pb2_film_write_image(self.film, c.as_ptr(), scale)) };
    }
    /// Sum of every rank's film into `root`'s (`ncclReduce` over NVLink); `pb2_nccl_init` first.
    pub fn reduce(&mut self, root: i32, stream: *mut c_void) {
        unsafe { check(pb2_film_reduce(self.film, root, stream)) };
    }

    /// The hand-off to the crate's `Film`: resolved pixel values (XYZ -> RGB, / weight, clamp) become its image through
    /// `Film::set_image` (film.rs:125-135), after which `Film::write_image` proceeds as it would after a CPU render.
    pub fn hand_off(&self, host: &mut Film) {
        let rgb = self.resolve_rgb(1.0, 1.0);
        let img: Vec<Spectrum> = rgb.chunks_exact(3).map(|c| Spectrum::from_rgb(c, SpectrumType::Reflectance)).collect();
        assert_eq!(img.len(), host.cropped_pixel_bounds.area() as usize);
        host.set_image(&img);
    }
}
impl Drop for B200Film {
    fn drop(&mut self) {
        unsafe { pb2_film_destroy(self.film) };
    }
}

// ------------------------------------------------------------------------------------------------------------------
// B200PathIntegrator — stands where PathIntegrator stands.
// ------------------------------------------------------------------------------------------------------------------
/// Which `Sampler` the reference would have been given (src/samplers/*.rs), in `pb2_path_desc` terms.
#[derive(Clone, Copy, Debug)]
pub enum SamplerKind {
    /// `RandomSampler::new(spp)` (random.rs:17-27)
    Random,
    /// `HaltonSampler::new(spp, sample_bounds, false)` (halton.rs:64-103)
    Halton,
    /// `StratifiedSampler::new(x_samples, y_samples, jitter, n_sampled_dimensions)` (stratified.rs:23-39)
    Stratified { x_samples: i32, y_samples: i32, jitter: bool, n_sampled_dimensions: i32 },
    /// `ZeroTwoSequenceSampler::new(spp, n_sampled_dimensions)` (zerotwosequence.rs:17-25)
    ZeroTwo { n_sampled_dimensions: i32 },
    /// `SobolSampler::new(spp, sample_bounds)` (sobol.rs:20-35)
    Sobol,
}

pub struct B200PathIntegrator {
    accel: Arc<B200Accel>,
    camera: CameraDt,
    camera_desc: pb2_camera,
    path: pb2_path_desc,
    film: B200Film,
    _pixel_bounds: Bounds2i,
}

impl B200PathIntegrator {
    /// `PathIntegrator::new(max_depth, camera, sampler, pixel_bounds, rr_threshold, light_sample_strategy)` (path.rs:31-46)
    /// with a `RandomSampler` of `spp` samples.  `camera` is the crate's camera (its `film()` receives the image);
    /// `camera_desc` restates its look-at, field of view and lens for the device (`PerspectiveCamera::new`, perspective.rs:34-82);
    /// `film` is the device film (`B200Film::for_film(&camera.film().read().unwrap(), ..)`).
    pub fn new(accel: Arc<B200Accel>, max_depth: usize, camera: CameraDt, camera_desc: pb2_camera, film: B200Film, spp: usize,
               pixel_bounds: Bounds2i, rr_threshold: Float, light_sample_strategy: &str) -> Self {
        Self::with_sampler(accel, max_depth, camera, camera_desc, film, spp, SamplerKind::Random, pixel_bounds, rr_threshold,
                           light_sample_strategy)
    }

    pub fn with_sampler(accel: Arc<B200Accel>, max_depth: usize, camera: CameraDt, camera_desc: pb2_camera, film: B200Film, spp: usize,
                        sampler: SamplerKind, pixel_bounds: Bounds2i, rr_threshold: Float, light_sample_strategy: &str) -> Self {
        // create_light_sample_distribution (lightdistrib.rs:222-232); any other name panics there too
        let light_strategy = match light_sample_strategy {
            "uniform" => PB2_LIGHTS_UNIFORM,
            "power" => PB2_LIGHTS_POWER,
            "spatial" => PB2_LIGHTS_SPATIAL,
            other => panic!("Light sample distribution type '{}' unknown", other),
        };
        let mut path = pb2_path_desc { max_depth: max_depth as i32, rr_threshold, light_strategy, spp: spp as i32, sample_begin: 0,
                                       sample_end: spp as i32, ..Default::default() };
        match sampler {
            SamplerKind::Random => path.sampler = PB2_SAMPLER_RANDOM,
            SamplerKind::Halton => path.sampler = PB2_SAMPLER_HALTON,
            SamplerKind::Stratified { x_samples, y_samples, jitter, n_sampled_dimensions } => {
                path.sampler = PB2_SAMPLER_STRATIFIED;
                path.x_samples = x_samples;
                path.y_samples = y_samples;
                path.jitter = jitter as i32;
                path.n_sampled_dimensions = n_sampled_dimensions;
            }
            SamplerKind::ZeroTwo { n_sampled_dimensions } => {
                path.sampler = PB2_SAMPLER_ZEROTWO;
                path.n_sampled_dimensions = n_sampled_dimensions;
            }
            SamplerKind::Sobol => path.sampler = PB2_SAMPLER_SOBOL,
        }
        B200PathIntegrator { accel, camera, camera_desc, path, film, _pixel_bounds: pixel_bounds }
    }

    /// `VolPathIntegrator::new` (src/integrators/volpath.rs:32-50) in place of `PathIntegrator::new`: same arguments; the scene's
    /// media come from `pb2_scene_set_media` on the accelerator's scene.
    pub fn use_volpath(&mut self) {
        self.path.integrator = PB2_INTEGRATOR_VOLPATH;
    }

    /// One process per GPU: render only sample indices `[begin, end)` of every pixel (the caller reduces the films).
    pub fn set_sample_range(&mut self, begin: usize, end: usize) {
        self.path.sample_begin = begin as i32;
        self.path.sample_end = end as i32;
    }

    pub fn film(&self) -> &B200Film {
        &self.film
    }
    pub fn film_mut(&mut self) -> &mut B200Film {
        &mut self.film
    }

    /// The wavefront render of `[sample_begin, sample_end)` into the device film, synchronous (`pb2_render_path` on the
    /// default stream, then a device synchronize).
    pub fn render_device(&mut self) {
        unsafe {
            check(pb2_render_path(self.accel.raw(), &self.camera_desc, &self.path, self.film.raw(), std::ptr::null_mut()));
            check(pb2_device_synchronize());
        }
    }

    /// `PathIntegrator::li` (path.rs:65-213) for explicit (pixel, sample index) pairs: `(L_rgb[3n], p_film[2n])`.
    pub fn li_many(&self, pixel_xy: &[u32], sample_index: &[u32]) -> (Vec<f32>, Vec<f32>) {
        let n = sample_index.len();
        assert_eq!(pixel_xy.len(), 2 * n);
        let (mut l, mut p) = (vec![0f32; 3 * n], vec![0f32; 2 * n]);
        unsafe { check(pb2_path_li(self.accel.raw(), &self.camera_desc, &self.path, pixel_xy.as_ptr(), sample_index.as_ptr(), n as u64,
                                   l.as_mut_ptr(), p.as_mut_ptr())) };
        (l, p)
    }

    /// `{camera_samples, extend_rays, shadow_rays, mis_rays, kernel_launches, ..}` of the last render.
    pub fn counters(&self) -> [u64; 8] {
        let mut c = [0u64; 8];
        unsafe { check(pb2_render_counters(self.accel.raw(), c.as_mut_ptr())) };
        c
    }
}

impl Integrator for B200PathIntegrator {
    fn as_any(&self) -> &dyn Any {
        self
    }

    /// `SamplerIntegrator::render` (integrator.rs:399-480): the whole tile loop is one `pb2_render_path`; then the device film
    /// becomes the camera film's image and `Film::write_image(1.0)` runs as at integrator.rs:479.  `_scene` is the crate's
    /// `Scene`, whose aggregate is the same `B200Accel` this integrator holds (`Scene.aggregate` is private, scene.rs:14).
    fn render(&mut self, _scene: &Scene) {
        self.film.clear();
        self.render_device();
        let host_film = self.camera.film();
        let mut host_film = host_film.write().unwrap();
        self.film.hand_off(&mut host_film);
        host_film.write_image(1.0);
    }
    // pre_process: the light distribution is built inside pb2_render_path (path.rs:58-63 does it in pre_process);
    // li: the trait default (`unimplemented!`, integrator.rs:33-41) stands — per-sample radiance is `li_many`.
}

// ------------------------------------------------------------------------------------------------------------------
// One process per GPU (north_star: samples split across GPUs, one NCCL reduce of the film).
// ------------------------------------------------------------------------------------------------------------------
/// Rank 0 creates the id and ships it to the other ranks by whatever channel the launcher has (file, env, MPI).
pub fn nccl_unique_id() -> [u8; 128] {
    let mut id = [0u8; 128];
    unsafe { check(pb2_nccl_unique_id(id.as_mut_ptr() as *mut c_char)) };
    id
}
pub fn nccl_init(id: &[u8; 128], rank: usize, n_ranks: usize) {
    unsafe { check(pb2_nccl_init(id.as_ptr() as *const c_char, rank as c_int, n_ranks as c_int)) };
}
pub fn nccl_shutdown() {
    unsafe { check(pb2_nccl_shutdown()) };
}
/// Sample indices `[begin, end)` of `spp` that rank `rank` of `n_ranks` renders: contiguous, sizes differing by at most one.
pub fn partition_samples(spp: usize, rank: usize, n_ranks: usize) -> (usize, usize) {
    let (q, r) = (spp / n_ranks, spp % n_ranks);
    let begin = rank * q + rank.min(r);
    (begin, begin + q + (rank < r) as usize)
}
