//! Rust side of the drop-in: `extern "C"` declarations for include/pbrt_b200.h and wrappers that implement the crate's
//! `Primitive` (src/core/primitive.rs:17-30) and `Integrator` (src/core/integrator.rs:29-42) traits on top of them.
//! Source only — the build image has no Rust toolchain; the Python ctypes binding (pbrt-rs_b200/__init__.py) exercises
//! the identical C ABI in the tests.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct pb2_ray { pub o: [f32; 3], pub t_max: f32, pub d: [f32; 3], pub time: f32 }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct pb2_hit { pub prim_id: u32, pub t: f32, pub b1: f32, pub b2: f32 }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct pb2_material { pub ty: i32 /* 0 matte, 1 plastic, 2 glass, 3 mirror, 4 metal, 5 substrate */, pub kd: [f32; 3], pub ks: [f32; 3], pub roughness: f32, pub remap_roughness: i32,
                          pub kr: [f32; 3], pub kt: [f32; 3], pub eta: f32,
                          pub sigma: f32 /* matte: Oren-Nayar, degrees */, pub metal_eta: [f32; 3], pub metal_k: [f32; 3] }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct pb2_light { pub ty: i32 /* 0 point, 1 area, 2 spot, 3 distant */, pub p: [f32; 3], pub i: [f32; 3], pub prim_id: u32, pub two_sided: i32,
                       pub axis: [f32; 3] /* spot: row 2 of world_to_light; distant: w */, pub total_width: f32, pub falloff_start: f32 }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct pb2_camera { pub pos: [f32; 3], pub look: [f32; 3], pub up: [f32; 3], pub fov: f32, pub res_x: i32, pub res_y: i32,
                        pub lens_radius: f32 /* 0 = pinhole */, pub focal_distance: f32 }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct pb2_film_desc { pub res_x: i32, pub res_y: i32, pub filter: i32 /* 0 box, 1 gaussian, 2 triangle, 3 mitchell, 4 sinc */,
                           pub radius_x: f32, pub radius_y: f32, pub gaussian_alpha: f32, pub mitchell_b: f32, pub mitchell_c: f32, pub sinc_tau: f32,
                           pub crop_window: [f32; 4] /* all zero = whole image */, pub max_sample_luminance: f32 /* <= 0 = infinity */ }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct pb2_path_desc { pub max_depth: i32, pub rr_threshold: f32, pub light_strategy: i32, pub spp: i32,
                           pub sample_begin: i32, pub sample_end: i32,
                           pub sampler: i32 /* 0 RandomSampler, 1 HaltonSampler, 2 StratifiedSampler, 3 ZeroTwoSequenceSampler */,
                           pub n_sampled_dimensions: i32, pub x_samples: i32, pub y_samples: i32, pub jitter: i32 }
pub enum pb2_scene {}
pub enum pb2_film {}
pub const PB2_MISS: u32 = 0xFFFF_FFFF;

extern "C" {
    pub fn pb2_init(device: c_int) -> c_int;
    pub fn pb2_last_error() -> *const c_char;
    pub fn pb2_scene_create(verts: *const f32, n_verts: u64, indices: *const u32, n_tris: u64, tri_material: *const u32,
                            mats: *const pb2_material, n_mats: u32, lights: *const pb2_light, n_lights: u32,
                            out: *mut *mut pb2_scene) -> c_int;
    pub fn pb2_scene_set_shading_geometry(scene: *mut pb2_scene, normals: *const f32, tangents: *const f32, uvs: *const f32) -> c_int;
    pub fn pb2_scene_destroy(scene: *mut pb2_scene) -> c_int;
    pub fn pb2_scene_build_bvh(scene: *mut pb2_scene, max_prims_in_node: c_int, split_method: c_int) -> c_int;
    pub fn pb2_world_bound(scene: *const pb2_scene, out: *mut f32) -> c_int;
    pub fn pb2_intersect(scene: *mut pb2_scene, rays: *const pb2_ray, n: u64, hits: *mut pb2_hit, b0: *mut f32) -> c_int;
    pub fn pb2_intersect_p(scene: *mut pb2_scene, rays: *const pb2_ray, n: u64, out: *mut u8) -> c_int;
    // asynchronous forms: buffers from pb2_host_alloc, untouched until pb2_scene_wait returns
    pub fn pb2_intersect_async(scene: *mut pb2_scene, rays: *const pb2_ray, n: u64, hits: *mut pb2_hit, b0: *mut f32) -> c_int;
    pub fn pb2_intersect_p_async(scene: *mut pb2_scene, rays: *const pb2_ray, n: u64, out: *mut u8) -> c_int;
    pub fn pb2_scene_wait(scene: *mut pb2_scene) -> c_int;
    pub fn pb2_film_create(desc: *const pb2_film_desc, out: *mut *mut pb2_film) -> c_int;
    pub fn pb2_film_destroy(film: *mut pb2_film) -> c_int;
    pub fn pb2_film_read_xyzw(film: *mut pb2_film, out: *mut f32) -> c_int;
    pub fn pb2_film_resolve_rgb(film: *mut pb2_film, scale: f32, rgb: *mut f32) -> c_int;
    pub fn pb2_film_write_image(film: *mut pb2_film, filename: *const c_char, scale: f32) -> c_int;
    pub fn pb2_bvh_build_stats(scene: *const pb2_scene, ms: *mut f64) -> c_int;
    pub fn pb2_render_path(scene: *mut pb2_scene, cam: *const pb2_camera, path: *const pb2_path_desc, film: *mut pb2_film,
                           stream: *mut c_void) -> c_int;
}

fn check(rc: c_int) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(pb2_last_error()) }.to_string_lossy().into_owned();
        panic!("pbrt_b200 error {}: {}", rc, msg);   // the reference's own error convention is panic!/unwrap
    }
}

/// Stands where `BVHAccel` stands: `Scene::new(aggregate, lights)` takes it as its `PrimitiveDt`.
pub struct B200Accel { scene: *mut pb2_scene }
unsafe impl Send for B200Accel {}
unsafe impl Sync for B200Accel {}

impl B200Accel {
    /// verts / indices: the world-space triangle list that would have been handed to `BVHAccel::new` as
    /// `GeometricPrimitive(Triangle)`s; `max_prims_in_node` and SAH as in src/accelerators/bvh.rs:216-222.
    pub fn new(verts: &[f32], indices: &[u32], max_prims_in_node: usize) -> Self {
        Self::with_split(verts, indices, max_prims_in_node, 0)
    }
    /// `split_method` follows `SplitMethod` (src/accelerators/bvh.rs:199-204): 0 = SAH (built on the host), 1 = HLBVH (built
    /// on the GPU, bvh.rs:475-772), 2 = Middle, 3 = EqualCounts (host, bvh.rs:331-360).
    pub fn with_split(verts: &[f32], indices: &[u32], max_prims_in_node: usize, split_method: c_int) -> Self {
        let mut scene = std::ptr::null_mut();
        unsafe {
            check(pb2_init(0));
            check(pb2_scene_create(verts.as_ptr(), (verts.len() / 3) as u64, indices.as_ptr(), (indices.len() / 3) as u64,
                                   std::ptr::null(), std::ptr::null(), 0, std::ptr::null(), 0, &mut scene));
            check(pb2_scene_build_bvh(scene, max_prims_in_node as c_int, split_method));
        }
        B200Accel { scene }
    }
    /// Batched `Primitive::intersect`: shrinks `t_max` of every ray that hits and returns the hit records.
    pub fn intersect_many(&self, rays: &mut [pb2_ray]) -> Vec<pb2_hit> {
        let mut hits = vec![pb2_hit::default(); rays.len()];
        unsafe { check(pb2_intersect(self.scene, rays.as_ptr(), rays.len() as u64, hits.as_mut_ptr(), std::ptr::null_mut())); }
        for (r, h) in rays.iter_mut().zip(&hits) { if h.prim_id != PB2_MISS { r.t_max = h.t; } }
        hits
    }
    pub fn intersect_p_many(&self, rays: &[pb2_ray]) -> Vec<bool> {
        let mut out = vec![0u8; rays.len()];
        unsafe { check(pb2_intersect_p(self.scene, rays.as_ptr(), rays.len() as u64, out.as_mut_ptr())); }
        out.into_iter().map(|b| b != 0).collect()
    }
}
impl Drop for B200Accel { fn drop(&mut self) { unsafe { pb2_scene_destroy(self.scene); } } }

// impl pbrt::core::primitive::Primitive for B200Accel (src/core/primitive.rs:17-30):
//   fn world_bound(&self) -> Bounds3f            -> pb2_world_bound
//   fn intersect(&self, r: &mut Ray, si: &mut SurfaceInteraction) -> bool
//                                                 -> intersect_many(&mut [ray]) (batch of 1: functional, slow); fills
//                                                    si.p / si.n from (prim_id, b1, b2) as Triangle::intersect does
//   fn intersect_p(&self, r: &Ray) -> bool        -> intersect_p_many(&[ray])[0]
//   get_area_light / get_material / compute_scattering_functions: unimplemented!() exactly like bvh.rs:934-953
//
// impl pbrt::core::integrator::Integrator for B200PathIntegrator (src/core/integrator.rs:29-42):
//   fn render(&mut self, scene: &Scene)           -> one pb2_render_path call, then pb2_film_read_xyzw -> Film::set_image
//   light_sample_strategy (path.rs:43, lightdistrib.rs:222-232): "uniform" -> 0, "power" -> 1, "spatial" -> 2 in
//                                                    pb2_path_desc.light_strategy; any other name panics in the reference and
//                                                    is PB2_ERR_INVALID here
