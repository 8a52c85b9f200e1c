"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes wrapper over the path-tracer half of oracle/liboracle.so (oracle_path.hpp): PathIntegrator::li, Film,
SamplerIntegrator::render restated on the CPU.  Parity unpinned (see oracle_core.hpp).
"""
import ctypes as C

import numpy as np

from . import oracle as O


class Material(C.Structure):
    _fields_ = [("type", C.c_int32), ("kd", C.c_float * 3), ("ks", C.c_float * 3), ("roughness", C.c_float),
                ("remap_roughness", C.c_int32), ("kr", C.c_float * 3), ("kt", C.c_float * 3), ("eta", C.c_float),
                ("sigma", C.c_float), ("metal_eta", C.c_float * 3), ("metal_k", C.c_float * 3)]


class Light(C.Structure):
    _fields_ = [("type", C.c_int32), ("p", C.c_float * 3), ("i", C.c_float * 3), ("prim_id", C.c_uint32),
                ("two_sided", C.c_int32), ("axis", C.c_float * 3), ("total_width", C.c_float), ("falloff_start", C.c_float)]


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
    """HomogeneousMedium::new(sigma_a, sigma_s, g) (src/media/homogeneous.rs:20-28)"""
    _fields_ = [("sigma_a", C.c_float * 3), ("sigma_s", C.c_float * 3), ("g", C.c_float)]


NO_MATERIAL = 0xFFFFFFFF       # tri_material / sphere material of a surface that only separates media (GeometricPrimitive.material = None)


def medium(d):
    m = Medium()
    m.sigma_a[:] = d["sigma_a"]
    m.sigma_s[:] = d["sigma_s"]
    m.g = d.get("g", 0.0)
    return m


class Sphere(C.Structure):
    _fields_ = [("object_to_world", C.c_float * 16), ("radius", C.c_float), ("z_min", C.c_float), ("z_max", C.c_float),
                ("phi_max", C.c_float), ("reverse_orientation", C.c_int32), ("material", C.c_uint32)]


def sphere(d):
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


_SAMPLER = {"random": 0, "halton": 1, "stratified": 2, "zerotwo": 3, "sobol": 4}
_MAT = {"matte": 0, "plastic": 1, "glass": 2, "mirror": 3, "metal": 4, "substrate": 5}
_STRAT = {"uniform": 0, "power": 1, "spatial": 2}
_FILTER = {"box": 0, "gaussian": 1, "triangle": 2, "mitchell": 3, "sinc": 4}


def material(d):
    m = Material()
    m.type = _MAT[d["type"]]
    m.kd[:] = d.get("kd", (0, 0, 0))
    m.ks[:] = d.get("ks", (0, 0, 0))
    m.roughness = d.get("roughness", 0.0 if d["type"] == "glass" else 0.1)      # glass: 0 = smooth (FresnelSpecular), > 0 = rough
    m.remap_roughness = int(d.get("remap", True))
    m.kr[:] = d.get("kr", (0, 0, 0))
    m.kt[:] = d.get("kt", (0, 0, 0))
    m.eta = d.get("eta", 1.5)
    m.sigma = d.get("sigma", 0.0)
    m.metal_eta[:] = d.get("metal_eta", (0.2, 0.92, 1.1))       # copper-ish defaults (pbrt's metal.cpp tabulates Cu)
    m.metal_k[:] = d.get("metal_k", (3.9, 2.45, 2.14))
    if d["type"] == "metal":
        m.roughness = d.get("roughness", 0.01)
    return m


def light(d):
    l = Light()
    if d["type"] == "point":
        l.type = 0
        l.p[:] = d["p"]
        l.i[:] = d["I"]
    elif d["type"] == "spot":          # axis = row 2 of world_to_light = normalize(to - from) for pbrt's from/to spot light
        l.type = 2
        l.p[:] = d["p"]
        l.i[:] = d["I"]
        l.axis[:] = d["axis"]
        l.total_width = d["total_width"]
        l.falloff_start = d["falloff_start"]
    elif d["type"] == "distant":
        l.type = 3
        l.i[:] = d["L"]
        l.axis[:] = d["w"]
    else:
        l.type = 1
        l.i[:] = d["L"]
        l.prim_id = d["prim"]
        l.two_sided = int(d.get("two_sided", False))
    return l


def camera_desc(cam):
    c = CameraDesc()
    c.pos[:] = cam["pos"]
    c.look[:] = cam["look"]
    c.up[:] = cam["up"]
    c.fov = cam["fov"]
    c.res_x, c.res_y = cam["res"]
    c.lens_radius = cam.get("lens_radius", 0.0)
    c.focal_distance = cam.get("focal_distance", 1e6)
    return c


def film_desc(res, filt="box", radius=(0.5, 0.5), alpha=2.0, b=1.0 / 3.0, c=1.0 / 3.0, tau=3.0, crop=None, max_sample_luminance=0.0):
    """crop = (x0, y0, x1, y1) fractions of the full resolution (Film::new's crop_window, film.rs:33); None = full image."""
    f = FilmDesc()
    f.res_x, f.res_y = res
    f.filter = _FILTER[filt]
    f.radius_x, f.radius_y = radius
    f.gaussian_alpha = alpha
    f.mitchell_b, f.mitchell_c, f.sinc_tau = b, c, tau
    if crop is not None:
        f.crop_window[:] = crop
    f.max_sample_luminance = max_sample_luminance
    return f


def film_bounds(film):
    """(cropped_pixel_bounds (x0, y0, x1, y1), sample bounds (x0, y0, x1, y1))."""
    out = np.zeros(8, np.int32)
    O.lib().orc_film_bounds(C.byref(film), _p(out))
    return tuple(int(v) for v in out[:4]), tuple(int(v) for v in out[4:])


def film_shape(film):
    (x0, y0, x1, y1), _ = film_bounds(film)
    return (y1 - y0, x1 - x0)


def path_desc(max_depth=5, rr_threshold=1.0, light_strategy="uniform", spp=1, sample_begin=0, sample_end=None, sampler="random",
              n_sampled_dimensions=4, x_samples=0, y_samples=0, jitter=True, integrator="path"):
    p = PathDesc()
    p.integrator = {"path": 0, "volpath": 1}[integrator]
    p.max_depth = max_depth
    p.rr_threshold = rr_threshold
    p.light_strategy = _STRAT[light_strategy]
    p.spp = spp
    p.sample_begin = sample_begin
    p.sample_end = spp if sample_end is None else sample_end
    p.sampler = _SAMPLER[sampler]
    p.n_sampled_dimensions = n_sampled_dimensions
    p.x_samples, p.y_samples, p.jitter = x_samples, y_samples, int(jitter)
    if sampler == "stratified":
        assert x_samples * y_samples == spp, "StratifiedSampler: spp = x_samples * y_samples (stratified.rs:31-32)"
    if sampler == "sobol":
        assert spp & (spp - 1) == 0, "SobolSampler rounds spp up to a power of two (sobol.rs:22-28); pass the rounded value"
    if sampler == "zerotwo":
        assert spp & (spp - 1) == 0, "ZeroTwoSequenceSampler rounds spp up to a power of two (zerotwosequence.rs:21); pass the rounded value"
    return p


def pixel_tables(path, table_sequence):
    """PixelSampler tables of one pixel: (t1 [n_dims, spp], t2 [n_dims, spp, 2])."""
    t1 = np.zeros((path.n_sampled_dimensions, path.spp), np.float32)
    t2 = np.zeros((path.n_sampled_dimensions, path.spp, 2), np.float32)
    O.lib().orc_pixel_tables(C.byref(path), C.c_uint64(int(table_sequence)), _p(t1), _p(t2))
    return t1, t2


def halton_probe(res, pixel, sample_num, n_dims=8, n_perm=32):
    """(sample index, first n_dims sample_dimension values, first n_perm entries of the permutation table)."""
    index = C.c_int64()
    dims = np.zeros(n_dims, np.float32)
    perm = np.zeros(n_perm, np.uint16)
    O.lib().orc_halton_probe(int(res[0]), int(res[1]), int(pixel[0]), int(pixel[1]), C.c_uint64(int(sample_num)), int(n_dims), C.byref(index),
                             _p(dims), int(n_perm), _p(perm))
    return index.value, dims, perm


def sobol_probe(sample_bounds, pixel, sample_num, n_dims=8, raw=False):
    """SobolSampler probe: (sample index, first n_dims dimensions).  sample_bounds = (x0, y0, w, h).  raw=True: sample_num is
    taken as the Sobol' index itself and the dimensions are sobol_sample(index, d) without the pixel remap."""
    index = C.c_int64()
    dims = np.zeros(n_dims, np.float32)
    rc = O.lib().orc_sobol_probe(int(sample_bounds[0]), int(sample_bounds[1]), int(sample_bounds[2]), int(sample_bounds[3]), int(pixel[0]),
                                 int(pixel[1]), C.c_uint64(int(sample_num)), int(n_dims), int(raw), C.byref(index), _p(dims))
    assert rc == 0, "the Sobol' table is not embedded in liboracle.so"
    return index.value, dims


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Scene:
    def __init__(self, sc, max_prims_in_node=4):
        L = O.lib()
        self.verts = np.ascontiguousarray(sc["verts"], dtype=np.float32).reshape(-1, 3)
        self.idx = np.ascontiguousarray(sc["idx"], dtype=np.uint32).reshape(-1, 3)
        tm = np.ascontiguousarray(sc["tri_material"], dtype=np.uint32)
        mats = (Material * len(sc["materials"]))(*[material(m) for m in sc["materials"]])
        lts = (Light * max(1, len(sc["lights"])))(*[light(l) for l in sc["lights"]])
        sph = sc.get("spheres") or []
        sphs = (Sphere * max(1, len(sph)))(*[sphere(d) for d in sph])
        self.n_spheres = len(sph)
        self.h = L.orc_scene_create2(_p(self.verts), len(self.verts), _p(self.idx), len(self.idx), _p(tm),
                                     C.cast(mats, C.c_void_p), len(sc["materials"]), C.cast(lts, C.c_void_p),
                                     len(sc["lights"]), max_prims_in_node, C.cast(sphs, C.c_void_p), len(sph))
        # participating media: sc["media"] = [dict(sigma_a, sigma_s, g)], sc["prim_inside"] / sc["prim_outside"] = medium index per
        # primitive (triangles, then spheres; -1 = none), sc["camera_medium"]
        if sc.get("media"):
            n_prims = len(self.idx) + len(sph)
            med = (Medium * len(sc["media"]))(*[medium(d) for d in sc["media"]])
            self._ins = np.ascontiguousarray(sc.get("prim_inside", np.full(n_prims, -1)), dtype=np.int32)
            self._outs = np.ascontiguousarray(sc.get("prim_outside", np.full(n_prims, -1)), dtype=np.int32)
            assert len(self._ins) == n_prims and len(self._outs) == n_prims
            L.orc_scene_set_media(C.c_void_p(self.h), C.cast(med, C.c_void_p), len(sc["media"]), _p(self._ins), _p(self._outs),
                                  int(sc.get("camera_medium", -1)))
        # TriangleMesh's optional per-vertex normals / tangents / UVs (triangle.rs:17-26)
        self._sg = [None if sc.get(k) is None else np.ascontiguousarray(sc[k], dtype=np.float32) for k in ("normals", "tangents", "uvs")]
        if any(a is not None for a in self._sg):
            L.orc_scene_set_shading_geometry(C.c_void_p(self.h), *[None if a is None else _p(a) for a in self._sg])

    def __del__(self):
        if getattr(self, "h", None):
            O.lib().orc_scene_free(self.h)
            self.h = None

    def bvh(self):
        b = O.BVHAccel(None, None, _handle=O.lib().orc_scene_bvh(self.h))
        b._scene = self          # the tree lives inside the scene: keep it alive as long as the view
        return b

    def render(self, cam, film, path, mode=1, threads=0, out=None):
        """Returns (xyzw [H,W,4] accumulators, seconds)."""
        cd, fd = camera_desc(cam), film
        if out is None:
            out = np.zeros(film_shape(fd) + (4,), dtype=np.float32)
        dt = O.lib().orc_render(self.h, C.byref(cd), C.byref(fd), C.byref(path), mode, threads, _p(out))
        return out, dt

    def render_counted(self, cam, film, path, mode=1, threads=0):
        """render() with the traversal instrumentation on: (xyzw, seconds, counters) where counters = {camera_samples, vertices,
        rays / nodes / tris per ray kind (extend, shadow, mis)} — the inputs of the path roofline."""
        cd = camera_desc(cam)
        out = np.zeros(film_shape(film) + (4,), dtype=np.float32)
        c = np.zeros(11, np.uint64)
        dt = O.lib().orc_render_counted(self.h, C.byref(cd), C.byref(film), C.byref(path), mode, threads, _p(out), _p(c))
        kinds = ("extend", "shadow", "mis")
        cnt = {"camera_samples": int(c[0]), "vertices": int(c[1]),
               "rays": {k: int(c[2 + i]) for i, k in enumerate(kinds)},
               "nodes": {k: int(c[5 + i]) for i, k in enumerate(kinds)},
               "tris": {k: int(c[8 + i]) for i, k in enumerate(kinds)}}
        return out, dt, cnt

    # SpatialLightDistribution probes (lightdistrib.rs:71-220)
    def spatial_grid(self):
        """Voxels per axis of the "spatial" light distribution (selects that strategy on the scene)."""
        nv = np.zeros(3, np.int32)
        O.lib().orc_spatial_grid(C.c_void_p(self.h), _p(nv))
        return tuple(int(v) for v in nv)

    def spatial_voxel(self, pi, n_lights):
        """(func [n_lights], cdf [n_lights + 1], func_int) of voxel pi = (x, y, z); call spatial_grid() first."""
        func = np.zeros(n_lights, np.float32)
        cdf = np.zeros(n_lights + 1, np.float32)
        fi = C.c_float()
        O.lib().orc_spatial_voxel(C.c_void_p(self.h), _p(np.asarray(pi, np.int32)), _p(func), _p(cdf), C.byref(fi))
        return func, cdf, np.float32(fi.value)

    def spatial_voxel_of(self, p):
        pi = np.zeros(3, np.int32)
        O.lib().orc_spatial_voxel_of(C.c_void_p(self.h), _p(np.asarray(p, np.float32)), _p(pi))
        return tuple(int(v) for v in pi)

    def path_li(self, cam, film, path, pixel_xy, sample_index):
        pixel_xy = np.ascontiguousarray(pixel_xy, dtype=np.int32).reshape(-1, 2)
        sample_index = np.ascontiguousarray(sample_index, dtype=np.uint32)
        n = len(pixel_xy)
        L_rgb = np.empty((n, 3), dtype=np.float32)
        pf = np.empty((n, 2), dtype=np.float32)
        cd = camera_desc(cam)
        O.lib().orc_path_li(self.h, C.byref(cd), C.byref(film), C.byref(path), _p(pixel_xy), _p(sample_index), n,
                            _p(L_rgb), _p(pf))
        return L_rgb, pf


def resolve_rgb(xyzw, scale=1.0):
    xyzw = np.ascontiguousarray(xyzw, dtype=np.float32)
    out = np.empty(xyzw.shape[:-1] + (3,), dtype=np.float32)
    O.lib().orc_resolve_rgb(_p(xyzw), xyzw.size // 4, scale, _p(out))
    return out


def film_add_samples(film, p_film, L_rgb, weight):
    p_film = np.ascontiguousarray(p_film, dtype=np.float32)
    L_rgb = np.ascontiguousarray(L_rgb, dtype=np.float32)
    weight = np.ascontiguousarray(weight, dtype=np.float32)
    out = np.zeros(film_shape(film) + (4,), dtype=np.float32)
    O.lib().orc_film_add_samples(C.byref(film), _p(p_film), _p(L_rgb), _p(weight), len(weight), _p(out))
    return out


def film_add_splats(film, p_film, v_rgb, splat_xyz=None):
    """Film::add_splat for every (p_film[i], v_rgb[i]) in array order -> splat_xyz [H, W, 3] (accumulates into splat_xyz if given)."""
    p_film = np.ascontiguousarray(p_film, dtype=np.float32)
    v_rgb = np.ascontiguousarray(v_rgb, dtype=np.float32)
    if splat_xyz is None:
        splat_xyz = np.zeros(film_shape(film) + (3,), dtype=np.float32)
    O.lib().orc_film_add_splats(C.byref(film), _p(p_film), _p(v_rgb), len(v_rgb), _p(splat_xyz))
    return splat_xyz


def resolve_rgb_splat(xyzw, splat_xyz, scale=1.0, splat_scale=1.0):
    """Film::write_image's pixel loop (film.rs:153-178) including the splat term."""
    xyzw = np.ascontiguousarray(xyzw, dtype=np.float32)
    splat_xyz = np.ascontiguousarray(splat_xyz, dtype=np.float32)
    out = np.empty(xyzw.shape[:-1] + (3,), dtype=np.float32)
    O.lib().orc_resolve_rgb_splat(_p(xyzw), _p(splat_xyz), xyzw.size // 4, scale, splat_scale, _p(out))
    return out


def rgb_to_xyz(rgb):
    rgb = np.ascontiguousarray(rgb, dtype=np.float32)
    out = np.empty_like(rgb)
    O.lib().orc_rgb_to_xyz(_p(rgb), rgb.size // 3, _p(out))
    return out


def roughness_to_alpha(r):
    return np.float32(O.lib().orc_roughness_to_alpha(r))


def sincos(x):
    s, c = C.c_float(), C.c_float()
    O.lib().orc_sincos(x, C.byref(s), C.byref(c))
    return np.float32(s.value), np.float32(c.value)


def film_table(film):
    t = np.empty(256, dtype=np.float32)
    O.lib().orc_film_table(C.byref(film), _p(t))
    return t
