"""pbrt_rs_b200 — thin ctypes binding over libpbrt_b200.so (include/pbrt_b200.h).

This is the Python-side stand-in for the Rust shim (rust_shim/): it mirrors the reference's trait surface
(`BVHAccel::{world_bound, intersect, intersect_p}`, `PerspectiveCamera::generate_ray`, `PathIntegrator::render`,
`Film::add_sample`) one to one over the C ABI so tests and bench.py exercise exactly what the FFI would bind.
There is no CPU fallback: a missing library or GPU raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PB2_LIB", os.path.join(_HERE, "libpbrt_b200.so"))   # PB2_LIB: tuning builds only
_LIB = None

PB2_MISS = 0xFFFFFFFF
HIT_DTYPE = np.dtype([("prim_id", np.uint32), ("t", np.float32), ("b1", np.float32), ("b2", np.float32)])
NODE_DTYPE = np.dtype([("bounds", np.float32, 6), ("offset", np.uint32), ("n_prims", np.uint16), ("axis", np.uint8),
                       ("pad", np.uint8)])

MAT_MATTE, MAT_PLASTIC, MAT_GLASS, MAT_MIRROR, MAT_METAL, MAT_SUBSTRATE = 0, 1, 2, 3, 4, 5
LIGHT_POINT, LIGHT_AREA, LIGHT_SPOT, LIGHT_DISTANT = 0, 1, 2, 3
FILTER_BOX, FILTER_GAUSSIAN, FILTER_TRIANGLE, FILTER_MITCHELL, FILTER_SINC = 0, 1, 2, 3, 4
LIGHTS_UNIFORM, LIGHTS_POWER, LIGHTS_SPATIAL = 0, 1, 2


class Pb2Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pb2 error {code}: {msg}")
        self.code = code


class Material(C.Structure):
    _fields_ = [("type", C.c_int32), ("kd", C.c_float * 3), ("ks", C.c_float * 3), ("roughness", C.c_float),
                ("remap_roughness", C.c_int32), ("kr", C.c_float * 3), ("kt", C.c_float * 3), ("eta", C.c_float),
                ("sigma", C.c_float), ("metal_eta", C.c_float * 3), ("metal_k", C.c_float * 3)]


class Light(C.Structure):
    _fields_ = [("type", C.c_int32), ("p", C.c_float * 3), ("i", C.c_float * 3), ("prim_id", C.c_uint32),
                ("two_sided", C.c_int32), ("axis", C.c_float * 3), ("total_width", C.c_float), ("falloff_start", C.c_float)]


class Sphere(C.Structure):
    """pb2_sphere: Sphere::new's arguments (src/shapes/sphere.rs:229-248) + the material of its GeometricPrimitive."""
    _fields_ = [("object_to_world", C.c_float * 16), ("radius", C.c_float), ("z_min", C.c_float), ("z_max", C.c_float),
                ("phi_max", C.c_float), ("reverse_orientation", C.c_int32), ("material", C.c_uint32)]


def sphere_from_dict(d):
    """dict(o2w=4x4 row-major affine matrix | center=(x, y, z), radius, z_min, z_max, phi_max (degrees), reverse_orientation, material)"""
    s = Sphere()
    if "o2w" in d:
        m = np.asarray(d["o2w"], dtype=np.float32).reshape(4, 4)
    else:
        m = np.eye(4, dtype=np.float32)
        m[:3, 3] = d.get("center", (0, 0, 0))
    s.object_to_world[:] = [float(v) for v in m.ravel()]
    s.radius = d["radius"]
    s.z_min = d.get("z_min", -d["radius"])
    s.z_max = d.get("z_max", d["radius"])
    s.phi_max = d.get("phi_max", 360.0)
    s.reverse_orientation = int(d.get("reverse_orientation", False))
    s.material = d.get("material", 0)
    return s


class CameraDesc(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("look", C.c_float * 3), ("up", C.c_float * 3), ("fov", C.c_float),
                ("res_x", C.c_int32), ("res_y", C.c_int32), ("lens_radius", C.c_float), ("focal_distance", C.c_float)]


class FilmDesc(C.Structure):
    _fields_ = [("res_x", C.c_int32), ("res_y", C.c_int32), ("filter", C.c_int32), ("radius_x", C.c_float),
                ("radius_y", C.c_float), ("gaussian_alpha", C.c_float), ("mitchell_b", C.c_float), ("mitchell_c", C.c_float),
                ("sinc_tau", C.c_float), ("crop_window", C.c_float * 4), ("max_sample_luminance", C.c_float)]


class PathDesc(C.Structure):
    _fields_ = [("max_depth", C.c_int32), ("rr_threshold", C.c_float), ("light_strategy", C.c_int32),
                ("spp", C.c_int32), ("sample_begin", C.c_int32), ("sample_end", C.c_int32), ("sampler", C.c_int32),
                ("n_sampled_dimensions", C.c_int32), ("x_samples", C.c_int32), ("y_samples", C.c_int32), ("jitter", C.c_int32),
                ("integrator", C.c_int32)]


class Medium(C.Structure):
    """pb2_medium: HomogeneousMedium::new(sigma_a, sigma_s, g) (src/media/homogeneous.rs:20-28)."""
    _fields_ = [("sigma_a", C.c_float * 3), ("sigma_s", C.c_float * 3), ("g", C.c_float)]


NO_MATERIAL = 0xFFFFFFFF       # PB2_NO_MATERIAL: a surface that only separates two media


def medium_from_dict(d):
    m = Medium()
    m.sigma_a[:] = d["sigma_a"]
    m.sigma_s[:] = d["sigma_s"]
    m.g = d.get("g", 0.0)
    return m


_SAMPLER = {"random": 0, "halton": 1, "stratified": 2, "zerotwo": 3, "sobol": 4}


def matte(kd, sigma=0.0):
    """pbrt-v3 MatteMaterial: LambertianReflection, or OrenNayar (reflection.rs:917-971) when sigma (degrees) != 0."""
    m = Material()
    m.type = MAT_MATTE
    m.kd[:] = kd
    m.sigma = sigma
    return m


def substrate(kd, ks, roughness=0.1, remap=True):
    """pbrt-v3 SubstrateMaterial (isotropic): FresnelBlend(Kd, Ks, TrowbridgeReitz) (reflection.rs:1194-1280)."""
    m = plastic(kd, ks, roughness, remap)
    m.type = MAT_SUBSTRATE
    return m


def mirror(kr=(0.9, 0.9, 0.9)):
    """pbrt-v3 MirrorMaterial: SpecularReflection(Kr, FresnelNoOp) (reflection.rs:606-659)."""
    m = Material()
    m.type = MAT_MIRROR
    m.kr[:] = kr
    return m


def metal(eta=(0.2, 0.92, 1.1), k=(3.9, 2.45, 2.14), roughness=0.01, remap=True):
    """pbrt-v3 MetalMaterial (isotropic): MicrofacetReflection(1, TrowbridgeReitz, FresnelConductor(1, eta, k))."""
    m = Material()
    m.type = MAT_METAL
    m.metal_eta[:] = eta
    m.metal_k[:] = k
    m.roughness = roughness
    m.remap_roughness = int(remap)
    return m


def plastic(kd, ks, roughness=0.1, remap=True):
    m = Material()
    m.type = MAT_PLASTIC
    m.kd[:] = kd
    m.ks[:] = ks
    m.roughness = roughness
    m.remap_roughness = int(remap)
    return m


def glass(kr=(1, 1, 1), kt=(1, 1, 1), eta=1.5, roughness=0.0, remap=True):
    """pbrt-v3 GlassMaterial: FresnelSpecular when smooth, MicrofacetReflection + MicrofacetTransmission when roughness > 0."""
    m = Material()
    m.type = MAT_GLASS
    m.kr[:] = kr
    m.kt[:] = kt
    m.eta = eta
    m.roughness = roughness
    m.remap_roughness = int(remap)
    return m


def point_light(p, intensity):
    l = Light()
    l.type = LIGHT_POINT
    l.p[:] = p
    l.i[:] = intensity
    return l


def spot_light(p, axis, intensity, total_width, falloff_start):
    """src/lights/spot.rs SpotLight::new: `axis` is the third row of world_to_light (normalize(to - from) for pbrt's from/to
    spot light), `total_width` / `falloff_start` the cone half-angles in degrees."""
    l = Light()
    l.type = LIGHT_SPOT
    l.p[:] = p
    l.i[:] = intensity
    l.axis[:] = axis
    l.total_width = total_width
    l.falloff_start = falloff_start
    return l


def distant_light(w, radiance):
    """src/lights/distant.rs DistantLight::new: `w` points towards the light; the scene's bounding sphere (pre_process) is
    taken from the BVH's world bound when the scene is built."""
    l = Light()
    l.type = LIGHT_DISTANT
    l.i[:] = radiance
    l.axis[:] = w
    return l


def area_light(prim_id, radiance, two_sided=False):
    l = Light()
    l.type = LIGHT_AREA
    l.i[:] = radiance
    l.prim_id = prim_id
    l.two_sided = int(two_sided)
    return l


def header_symbols():
    """Every pb2_* function include/pbrt_b200.h declares."""
    import re
    text = open(os.path.join(_HERE, "..", "include", "pbrt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pb2_[a-z0-9_]+)\s*\(", text)))


def lib():
    """Load libpbrt_b200.so; raises if it has not been built (no fallback path exists)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise Pb2Error(-2, f"{LIB_PATH} is missing: run `make -C pbrt-rs_b200` (or __graft_entry__.build()); "
                           "there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, u64, i32, u32, f32 = C.c_void_p, C.c_uint64, C.c_int, C.c_uint32, C.c_float
    L.pb2_last_error.restype = C.c_char_p
    sig = {
        "pb2_init": [i32], "pb2_shutdown": [], "pb2_device_count": [vp],
        "pb2_host_alloc": [u64, vp], "pb2_host_free": [vp], "pb2_device_alloc": [u64, vp], "pb2_device_free": [vp],
        "pb2_memcpy_h2d": [vp, vp, u64], "pb2_memcpy_d2h": [vp, vp, u64], "pb2_device_synchronize": [],
        "pb2_set_trace_tuning": [i32, i32, i32, i32],
        "pb2_scene_create": [vp, u64, vp, u64, vp, vp, u32, vp, u32, vp], "pb2_scene_destroy": [vp],
        "pb2_scene_set_shading_geometry": [vp, vp, vp, vp], "pb2_scene_add_spheres": [vp, vp, u32],
        "pb2_scene_set_media": [vp, vp, u32, vp, vp, C.c_int32],
        "pb2_scene_build_bvh": [vp, i32, i32], "pb2_scene_build_bvh_host": [vp, i32, i32], "pb2_world_bound": [vp, vp], "pb2_bvh_info": [vp, vp, vp, vp],
        "pb2_bvh_export": [vp, vp, vp], "pb2_bvh_build_stats": [vp, vp],
        "pb2_intersect": [vp, vp, u64, vp, vp], "pb2_intersect_p": [vp, vp, u64, vp],
        "pb2_intersect_async": [vp, vp, u64, vp, vp], "pb2_intersect_p_async": [vp, vp, u64, vp], "pb2_scene_wait": [vp], "pb2_scene_wait_until": [vp, C.c_uint32],
        "pb2_intersect_device": [vp, vp, u64, vp, vp, vp], "pb2_intersect_p_device": [vp, vp, u64, vp, vp],
        "pb2_camera_generate_rays": [vp, vp, vp, u64, vp], "pb2_camera_primary_rays_device": [vp, vp, vp],
        "pb2_camera_matrices": [vp, vp, vp],
        "pb2_spawn_shadow_rays_device": [vp, vp, vp, u64, vp, vp, vp],
        "pb2_spawn_bounce_rays_device": [vp, vp, vp, u64, vp, vp],
        "pb2_spawn_shadow_bounce_rays_device": [vp, vp, vp, u64, vp, vp, vp, vp],
        "pb2_rng_uniform_floats": [u64, u32, u32, vp], "pb2_film_bounds": [vp, vp, vp],
        "pb2_film_create": [vp, vp], "pb2_film_destroy": [vp], "pb2_film_clear": [vp],
        "pb2_film_add_samples": [vp, vp, vp, vp, u64], "pb2_film_read_xyzw": [vp, vp],
        "pb2_film_resolve_rgb": [vp, f32, vp], "pb2_film_resolve_rgb_splat": [vp, f32, f32, vp],
        "pb2_film_add_splats": [vp, vp, vp, u64], "pb2_film_set_image": [vp, vp], "pb2_film_device_ptr": [vp, vp, vp],
        "pb2_film_write_image": [vp, C.c_char_p, f32],
        "pb2_render_path": [vp, vp, vp, vp, vp], "pb2_path_li": [vp, vp, vp, vp, vp, u64, vp, vp],
        "pb2_render_counters": [vp, vp], "pb2_spatial_light_distribution": [vp, vp, vp, vp, vp],
        "pb2_nccl_unique_id": [vp], "pb2_nccl_init": [vp, i32, i32], "pb2_nccl_shutdown": [],
        "pb2_film_reduce": [vp, i32, vp],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int
    _LIB = L
    return L


def check(rc):
    if rc != 0:
        raise Pb2Error(rc, lib().pb2_last_error().decode())


def init(device=0):
    check(lib().pb2_init(device))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def camera_desc(pos, look, up, fov, res, lens_radius=0.0, focal_distance=1e6):
    c = CameraDesc()
    c.pos[:] = pos
    c.look[:] = look
    c.up[:] = up
    c.fov = fov
    c.res_x, c.res_y = res
    c.lens_radius, c.focal_distance = lens_radius, focal_distance
    return c


class DeviceBuffer:
    """Raw device allocation (pb2_device_alloc) for the device-resident entry points."""

    def __init__(self, nbytes):
        self.ptr = C.c_void_p()
        self.nbytes = int(nbytes)
        check(lib().pb2_device_alloc(self.nbytes, C.byref(self.ptr)))

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        check(lib().pb2_memcpy_h2d(self.ptr, _p(arr), arr.nbytes))
        return self

    def download(self, dtype, count):
        out = np.empty(count, dtype=dtype)
        assert out.nbytes <= self.nbytes
        check(lib().pb2_memcpy_d2h(_p(out), self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr:
            lib().pb2_device_free(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Scene:
    """Triangle list + materials + lights handed to BVHAccel::new (pb2_scene_create); normals / tangents / uvs are
    TriangleMesh's optional per-vertex arrays (src/shapes/triangle.rs:17-26)."""

    def __init__(self, verts, idx, tri_material=None, materials=None, lights=None, normals=None, tangents=None, uvs=None, spheres=None,
                 media=None, prim_inside=None, prim_outside=None, camera_medium=-1):
        verts = _f32(verts).reshape(-1, 3)
        idx = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1, 3)
        self.n_tris = len(idx)
        self.n_lights = len(lights) if lights else 0
        tm = None if tri_material is None else np.ascontiguousarray(tri_material, dtype=np.uint32)
        mats = (Material * len(materials))(*materials) if materials else None
        lts = (Light * len(lights))(*lights) if lights else None
        self.h = C.c_void_p()
        check(lib().pb2_scene_create(_p(verts), len(verts), _p(idx), len(idx), _p(tm),
                                     C.cast(mats, C.c_void_p) if mats else None, len(materials) if materials else 0,
                                     C.cast(lts, C.c_void_p) if lts else None, len(lights) if lights else 0,
                                     C.byref(self.h)))
        if normals is not None or tangents is not None or uvs is not None:
            sg = [None if a is None else _f32(a).reshape(len(verts), k) for a, k in ((normals, 3), (tangents, 3), (uvs, 2))]
            check(lib().pb2_scene_set_shading_geometry(self.h, *[_p(a) for a in sg]))
        self.n_spheres = len(spheres) if spheres else 0
        if spheres:                       # analytic spheres: primitive ids n_tris .. n_tris + n_spheres - 1 (pb2_scene_add_spheres)
            arr = (Sphere * len(spheres))(*spheres)
            check(lib().pb2_scene_add_spheres(self.h, C.cast(arr, C.c_void_p), len(spheres)))
        if media:                         # HomogeneousMedium list + every primitive's MediumInterface (pb2_scene_set_media)
            arr = (Medium * len(media))(*media)
            ins = None if prim_inside is None else np.ascontiguousarray(prim_inside, dtype=np.int32)
            outs = None if prim_outside is None else np.ascontiguousarray(prim_outside, dtype=np.int32)
            check(lib().pb2_scene_set_media(self.h, C.cast(arr, C.c_void_p), len(media), _p(ins), _p(outs), int(camera_medium)))

    def destroy(self):
        if self.h:
            lib().pb2_scene_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class BVHAccel:
    """Mirror of src/accelerators/bvh.rs BVHAccel behind the Primitive trait (world_bound / intersect / intersect_p)."""

    SAH = 0

    def __init__(self, scene_or_verts, idx=None, max_prims_in_node=4, split_method=0, host_only=False):
        self.scene = scene_or_verts if isinstance(scene_or_verts, Scene) else Scene(scene_or_verts, idx)
        build = lib().pb2_scene_build_bvh_host if host_only else lib().pb2_scene_build_bvh
        check(build(self.scene.h, max_prims_in_node, split_method))

    @property
    def h(self):
        return self.scene.h

    def world_bound(self):
        out = np.empty(6, dtype=np.float32)
        check(lib().pb2_world_bound(self.h, _p(out)))
        return out

    def spatial_light_distribution(self, tables=True):
        """SpatialLightDistribution (src/core/lightdistrib.rs:71-220) of the scene: (n_voxels (x, y, z), func [z, y, x, n_lights],
        cdf [z, y, x, n_lights + 1], func_int [z, y, x]); tables=False returns only the grid extents."""
        nv = np.zeros(3, dtype=np.int32)
        check(lib().pb2_spatial_light_distribution(self.h, _p(nv), None, None, None))
        if not tables:
            return tuple(int(v) for v in nv)
        n = self.scene.n_lights
        shape = (int(nv[2]), int(nv[1]), int(nv[0]))
        func = np.empty(shape + (n,), dtype=np.float32)
        cdf = np.empty(shape + (n + 1,), dtype=np.float32)
        func_int = np.empty(shape, dtype=np.float32)
        check(lib().pb2_spatial_light_distribution(self.h, _p(nv), _p(func), _p(cdf), _p(func_int)))
        return tuple(int(v) for v in nv), func, cdf, func_int

    def build_stats(self):
        """HLBVH stage times in ms: upload + bounds + Morton, sort, treelets, upper SAH (host), flatten, device-layout repack."""
        ms = (C.c_double * 6)()
        check(lib().pb2_bvh_build_stats(self.h, ms))
        return list(ms)

    def info(self):
        n_nodes, n_prims, depth = C.c_uint64(), C.c_uint64(), C.c_int()
        check(lib().pb2_bvh_info(self.h, C.byref(n_nodes), C.byref(n_prims), C.byref(depth)))
        return n_nodes.value, n_prims.value, depth.value

    def export(self):
        n_nodes, n_prims, _ = self.info()
        nodes = np.empty(n_nodes, dtype=NODE_DTYPE)
        prims = np.empty(n_prims, dtype=np.uint32)
        check(lib().pb2_bvh_export(self.h, _p(nodes), _p(prims)))
        return nodes, prims

    def intersect(self, rays, want_b0=False):
        """Closest hit per ray -> structured array (prim_id, t, b1, b2) [+ b0]."""
        rays = _f32(rays).reshape(-1, 8)
        hits = np.empty(len(rays), dtype=HIT_DTYPE)
        b0 = np.empty(len(rays), dtype=np.float32) if want_b0 else None
        check(lib().pb2_intersect(self.h, _p(rays), len(rays), _p(hits), _p(b0)))
        return (hits, b0) if want_b0 else hits

    def intersect_p(self, rays):
        rays = _f32(rays).reshape(-1, 8)
        out = np.empty(len(rays), dtype=np.uint8)
        check(lib().pb2_intersect_p(self.h, _p(rays), len(rays), _p(out)))
        return out

    # device-resident variants (pointers are ints / c_void_p; stream is a cudaStream_t value)
    def intersect_device(self, d_rays, n, d_hits, d_b0=None, stream=None):
        check(lib().pb2_intersect_device(self.h, d_rays, n, d_hits, d_b0, stream))

    def intersect_p_device(self, d_rays, n, d_out, stream=None):
        check(lib().pb2_intersect_p_device(self.h, d_rays, n, d_out, stream))

    def spawn_shadow_rays_device(self, d_rays, d_hits, n, light_pos, d_out, stream=None):
        lp = np.asarray(light_pos, dtype=np.float32)
        check(lib().pb2_spawn_shadow_rays_device(self.h, d_rays, d_hits, n, _p(lp), d_out, stream))

    def spawn_bounce_rays_device(self, d_rays, d_hits, n, d_out, stream=None):
        check(lib().pb2_spawn_bounce_rays_device(self.h, d_rays, d_hits, n, d_out, stream))

    def spawn_shadow_bounce_rays_device(self, d_rays, d_hits, n, light_pos, d_out_shadow, d_out_bounce, stream=None):
        lp = np.asarray(light_pos, dtype=np.float32)
        check(lib().pb2_spawn_shadow_bounce_rays_device(self.h, d_rays, d_hits, n, _p(lp), d_out_shadow, d_out_bounce, stream))


class PerspectiveCamera:
    """Mirror of src/cameras/perspective.rs: pinhole, or thin lens when lens_radius > 0 (perspective.rs:101-107)."""

    def __init__(self, pos, look, up, fov, res, lens_radius=0.0, focal_distance=1e6):
        self.desc = camera_desc(pos, look, up, fov, res, lens_radius, focal_distance)
        self.res = tuple(res)

    def matrices(self):
        r2c = np.empty((4, 4), dtype=np.float32)
        c2w = np.empty((4, 4), dtype=np.float32)
        check(lib().pb2_camera_matrices(C.byref(self.desc), _p(r2c), _p(c2w)))
        return r2c, c2w

    def generate_rays(self, p_film, p_lens=None):
        """Camera::generate_ray per CameraSample (p_film [n,2] raster points, p_lens [n,2] lens samples for a thin lens)."""
        p_film = _f32(p_film).reshape(-1, 2)
        p_lens = None if p_lens is None else _f32(p_lens).reshape(-1, 2)
        rays = np.empty((len(p_film), 8), dtype=np.float32)
        check(lib().pb2_camera_generate_rays(C.byref(self.desc), _p(p_film), _p(p_lens), len(p_film), _p(rays)))
        return rays

    def primary_rays_device(self, d_rays, stream=None):
        check(lib().pb2_camera_primary_rays_device(C.byref(self.desc), d_rays, stream))


_MAT = {"matte": MAT_MATTE, "plastic": MAT_PLASTIC, "glass": MAT_GLASS, "mirror": MAT_MIRROR, "metal": MAT_METAL, "substrate": MAT_SUBSTRATE}
_STRATEGY = {"uniform": LIGHTS_UNIFORM, "power": LIGHTS_POWER, "spatial": LIGHTS_SPATIAL}
_FILTER = {"box": FILTER_BOX, "gaussian": FILTER_GAUSSIAN, "triangle": FILTER_TRIANGLE, "mitchell": FILTER_MITCHELL, "sinc": FILTER_SINC}


def material_from_dict(d):
    if d["type"] == "matte":
        return matte(d["kd"], d.get("sigma", 0.0))
    if d["type"] == "mirror":
        return mirror(d.get("kr", (0.9, 0.9, 0.9)))
    if d["type"] == "metal":
        return metal(d.get("metal_eta", (0.2, 0.92, 1.1)), d.get("metal_k", (3.9, 2.45, 2.14)), d.get("roughness", 0.01), d.get("remap", True))
    if d["type"] == "substrate":
        return substrate(d["kd"], d["ks"], d.get("roughness", 0.1), d.get("remap", True))
    if d["type"] == "plastic":
        return plastic(d["kd"], d["ks"], d.get("roughness", 0.1), d.get("remap", True))
    return glass(d.get("kr", (1, 1, 1)), d.get("kt", (1, 1, 1)), d.get("eta", 1.5), d.get("roughness", 0.0), d.get("remap", True))


def light_from_dict(d):
    if d["type"] == "point":
        return point_light(d["p"], d["I"])
    if d["type"] == "spot":
        return spot_light(d["p"], d["axis"], d["I"], d["total_width"], d["falloff_start"])
    if d["type"] == "distant":
        return distant_light(d["w"], d["L"])
    return area_light(d["prim"], d["L"], d.get("two_sided", False))


def scene_from_dict(sc):
    """Scene from the plain-dict description the generators in scenes.py return."""
    return Scene(sc["verts"], sc["idx"], sc["tri_material"], [material_from_dict(m) for m in sc["materials"]],
                 [light_from_dict(l) for l in sc["lights"]], normals=sc.get("normals"), tangents=sc.get("tangents"), uvs=sc.get("uvs"),
                 spheres=[sphere_from_dict(d) for d in sc.get("spheres") or []],
                 media=[medium_from_dict(d) for d in sc.get("media") or []], prim_inside=sc.get("prim_inside"),
                 prim_outside=sc.get("prim_outside"), camera_medium=sc.get("camera_medium", -1))


class Film:
    """Mirror of src/core/film.rs Film (+ FilmTile::add_sample), accumulators resident on the device.  filter: "box",
    "gaussian" (alpha), "triangle", "mitchell" (b, c), "sinc" (tau) — src/filters/*.rs; crop = (x0, y0, x1, y1) fractions of the
    full resolution (Film::new's crop_window); max_sample_luminance as in Film::new (None = infinity).  `res` is the size of
    the stored image (cropped_pixel_bounds), `full_res` the full resolution."""

    def __init__(self, res, filter="box", radius=(0.5, 0.5), alpha=2.0, b=1.0 / 3.0, c=1.0 / 3.0, tau=3.0, crop=None, max_sample_luminance=None):
        self.desc = FilmDesc()
        self.desc.res_x, self.desc.res_y = res
        self.desc.filter = _FILTER[filter]
        self.desc.radius_x, self.desc.radius_y = radius
        self.desc.gaussian_alpha = alpha
        self.desc.mitchell_b, self.desc.mitchell_c, self.desc.sinc_tau = b, c, tau
        if crop is not None:
            self.desc.crop_window[:] = crop
        self.desc.max_sample_luminance = 0.0 if max_sample_luminance is None else max_sample_luminance
        self.full_res = tuple(res)
        self.h = C.c_void_p()
        check(lib().pb2_film_create(C.byref(self.desc), C.byref(self.h)))
        self.pixel_bounds, self.sample_bounds = self.bounds()
        self.res = (self.pixel_bounds[2] - self.pixel_bounds[0], self.pixel_bounds[3] - self.pixel_bounds[1])

    def bounds(self):
        """(cropped_pixel_bounds, sample bounds), each (x0, y0, x1, y1) with exclusive maxima."""
        pb, sb = (C.c_int32 * 4)(), (C.c_int32 * 4)()
        check(lib().pb2_film_bounds(self.h, pb, sb))
        return tuple(pb), tuple(sb)

    def destroy(self):
        if self.h:
            lib().pb2_film_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def clear(self):
        check(lib().pb2_film_clear(self.h))

    def add_samples(self, p_film, L_rgb, weight):
        p_film, L_rgb, weight = _f32(p_film), _f32(L_rgb), _f32(weight)
        check(lib().pb2_film_add_samples(self.h, _p(p_film), _p(L_rgb), _p(weight), len(weight)))

    def read_xyzw(self):
        out = np.empty((self.res[1], self.res[0], 4), dtype=np.float32)
        check(lib().pb2_film_read_xyzw(self.h, _p(out)))
        return out

    def resolve_rgb(self, scale=1.0, splat_scale=1.0):
        out = np.empty((self.res[1], self.res[0], 3), dtype=np.float32)
        check(lib().pb2_film_resolve_rgb_splat(self.h, scale, splat_scale, _p(out)))
        return out

    def add_splats(self, p_film, v_rgb):
        """Film::add_splat (film.rs:137-151) for every (p_film[i], v_rgb[i])."""
        p_film, v_rgb = _f32(p_film).reshape(-1, 2), _f32(v_rgb).reshape(-1, 3)
        check(lib().pb2_film_add_splats(self.h, _p(p_film), _p(v_rgb), len(p_film)))

    def set_image(self, rgb):
        """Film::set_image (film.rs:125-135): rgb [H, W, 3] of the cropped pixel bounds."""
        rgb = _f32(rgb)
        assert rgb.size == self.res[0] * self.res[1] * 3
        check(lib().pb2_film_set_image(self.h, _p(rgb)))

    def write_image(self, filename, scale=1.0):
        """Film::write_image (film.rs:153-180) through to a .pfm / .ppm file."""
        check(lib().pb2_film_write_image(self.h, os.fsencode(filename), scale))

    def device_ptr(self):
        ptr, n = C.c_void_p(), C.c_uint64()
        check(lib().pb2_film_device_ptr(self.h, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def reduce(self, root=0, stream=None):
        check(lib().pb2_film_reduce(self.h, root, stream))


class PathIntegrator:
    """Mirror of src/integrators/path.rs PathIntegrator + SamplerIntegrator::render; sampler = "random" (RandomSampler streams
    per (pixel, sample)), "halton" (HaltonSampler, src/samplers/halton.rs), "stratified" (StratifiedSampler::new(x_samples,
    y_samples, jitter, n_sampled_dimensions), src/samplers/stratified.rs) or "zerotwo" (ZeroTwoSequenceSampler::new(spp,
    n_sampled_dimensions), src/samplers/zerotwosequence.rs: spp is rounded up to a power of two as the constructor does) or
    "sobol" (SobolSampler::new(spp, sample_bounds), src/samplers/sobol.rs; spp rounded up likewise)."""

    def __init__(self, accel, camera, max_depth=5, rr_threshold=1.0, light_strategy="uniform", spp=1, sampler="random",
                 n_sampled_dimensions=4, x_samples=0, y_samples=0, jitter=True, integrator="path"):
        """integrator: "path" = PathIntegrator (wavefront), "volpath" = VolPathIntegrator (src/integrators/volpath.rs; media)."""
        if sampler in ("zerotwo", "sobol"):
            spp = 1 << max(0, int(spp) - 1).bit_length()             # round_up_pow2_i64, zerotwosequence.rs:21, sobol.rs:22-28
        if sampler == "stratified" and x_samples and y_samples:
            spp = x_samples * y_samples                              # stratified.rs:31-32
        self.accel, self.camera = accel, camera
        self.desc = PathDesc()
        self.desc.max_depth = max_depth
        self.desc.rr_threshold = rr_threshold
        self.desc.light_strategy = _STRATEGY[light_strategy]
        self.desc.spp = spp
        self.desc.sample_begin, self.desc.sample_end = 0, spp
        self.desc.sampler = _SAMPLER[sampler]
        self.desc.n_sampled_dimensions = n_sampled_dimensions
        self.desc.x_samples, self.desc.y_samples, self.desc.jitter = x_samples, y_samples, int(jitter)
        self.desc.integrator = {"path": 0, "volpath": 1}[integrator]
        self.spp = spp

    def render(self, film, sample_begin=0, sample_end=None, stream=None):
        """Integrator::render: accumulates sample indices [sample_begin, sample_end) of every pixel into `film`."""
        self.desc.sample_begin = sample_begin
        self.desc.sample_end = self.desc.spp if sample_end is None else sample_end
        check(lib().pb2_render_path(self.accel.h, C.byref(self.camera.desc), C.byref(self.desc), film.h, stream))

    def li(self, pixel_xy, sample_index):
        """PathIntegrator::li per explicit (pixel, sample) -> (L_rgb [n,3], p_film [n,2])."""
        pixel_xy = np.ascontiguousarray(pixel_xy, dtype=np.uint32).reshape(-1, 2)
        sample_index = np.ascontiguousarray(sample_index, dtype=np.uint32)
        n = len(sample_index)
        L = np.empty((n, 3), dtype=np.float32)
        pf = np.empty((n, 2), dtype=np.float32)
        self.desc.sample_begin, self.desc.sample_end = 0, self.desc.spp
        check(lib().pb2_path_li(self.accel.h, C.byref(self.camera.desc), C.byref(self.desc), _p(pixel_xy), _p(sample_index), n,
                                _p(L), _p(pf)))
        return L, pf

    def counters(self):
        out = np.zeros(8, dtype=np.uint64)
        check(lib().pb2_render_counters(self.accel.h, _p(out)))
        return dict(camera_samples=int(out[0]), extend_rays=int(out[1]), shadow_rays=int(out[2]), mis_rays=int(out[3]),
                    kernel_launches=int(out[4]), stray_overflow=int(out[5]))


def partition_samples(spp, rank, world):
    """Sample-index range of every pixel that GPU `rank` of `world` renders: contiguous, disjoint, covering
    [0, spp); the first spp % world ranks take one extra index.  Every (pixel, sample) keeps its own sampler stream, so the
    union over ranks draws exactly the random numbers a single GPU would."""
    if world < 1 or not (0 <= rank < world) or spp < 0:
        raise ValueError("bad partition arguments")
    base, extra = divmod(spp, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    check(lib().pb2_nccl_unique_id(buf))
    return buf.raw


def nccl_init(unique_id, rank, n_ranks):
    check(lib().pb2_nccl_init(C.create_string_buffer(unique_id, 128), rank, n_ranks))


def nccl_shutdown():
    check(lib().pb2_nccl_shutdown())


def rng_uniform_floats(first_sequence, n_sequences, n_per):
    out = np.empty((n_sequences, n_per), dtype=np.float32)
    check(lib().pb2_rng_uniform_floats(first_sequence, n_sequences, n_per, _p(out)))
    return out
