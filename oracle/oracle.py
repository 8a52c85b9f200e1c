"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes wrapper over oracle/liboracle.so (the CPU restatement of the pbrt-rs hot path; parity unpinned,
see oracle_core.hpp).  Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

HIT_DTYPE = np.dtype([("prim_id", np.uint32), ("t", np.float32), ("b1", np.float32), ("b2", np.float32)])
NODE_DTYPE = np.dtype([("bounds", np.float32, 6), ("offset", np.uint32), ("n_prims", np.uint16), ("axis", np.uint8),
                       ("pad", np.uint8)])


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_gamma.restype = C.c_float
        L.orc_gamma.argtypes = [C.c_float]
        L.orc_next_float_up.restype = C.c_float
        L.orc_next_float_up.argtypes = [C.c_float]
        L.orc_next_float_down.restype = C.c_float
        L.orc_next_float_down.argtypes = [C.c_float]
        L.orc_slab_widen.restype = C.c_float
        L.orc_bvh_build.restype = C.c_void_p
        L.orc_bvh_build.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int]
        L.orc_bvh_build2.restype = C.c_void_p
        L.orc_bvh_build2.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.c_int]
        L.orc_bvh_free.argtypes = [C.c_void_p]
        for f in ("orc_bvh_num_nodes", "orc_bvh_num_prims"):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_bvh_max_depth.argtypes = [C.c_void_p]
        L.orc_bvh_get_nodes.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_bvh_get_ordered_prims.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_bvh_world_bound.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_intersect.restype = C.c_double
        L.orc_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_intersect_p.restype = C.c_double
        L.orc_intersect_p.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_brute_force.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]
        L.orc_slab_test.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_triangle_test.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_pcg32_u32.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
        L.orc_pcg32_float.argtypes = [C.c_uint64, C.c_int, C.c_void_p]
        L.orc_camera_matrices.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_camera_primary_rays.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p]
        L.orc_camera_rays.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_camera_rays_lens.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        _bind_path(L)
        _LIB = L
    return _LIB


def _bind_path(L):
    vp = C.c_void_p
    L.orc_scene_create.restype = vp
    L.orc_scene_create.argtypes = [vp, C.c_uint64, vp, C.c_uint64, vp, vp, C.c_uint32, vp, C.c_uint32, C.c_int]
    L.orc_scene_create2.restype = vp
    L.orc_scene_create2.argtypes = [vp, C.c_uint64, vp, C.c_uint64, vp, vp, C.c_uint32, vp, C.c_uint32, C.c_int, vp, C.c_uint32]
    L.orc_sphere_sample2.argtypes = [vp, C.c_uint32, vp, C.c_float, C.c_float, vp]
    L.orc_sphere_pdf2.restype = C.c_float
    L.orc_sphere_pdf2.argtypes = [vp, C.c_uint32, vp, vp]
    L.orc_acos.restype = C.c_float
    L.orc_acos.argtypes = [C.c_float]
    L.orc_atan2.restype = C.c_float
    L.orc_atan2.argtypes = [C.c_float, C.c_float]
    L.orc_scene_set_media.argtypes = [vp, vp, C.c_uint32, vp, vp, C.c_int32]
    L.orc_exp.restype = C.c_float
    L.orc_exp.argtypes = [C.c_float]
    L.orc_log.restype = C.c_float
    L.orc_log.argtypes = [C.c_float]
    L.orc_scene_free.argtypes = [vp]
    L.orc_scene_bvh.restype = vp
    L.orc_scene_bvh.argtypes = [vp]
    L.orc_render.restype = C.c_double
    L.orc_render.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, vp]
    L.orc_path_li.argtypes = [vp, vp, vp, vp, vp, vp, C.c_uint64, vp, vp]
    L.orc_render_counted.restype = C.c_double
    L.orc_render_counted.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, vp, vp]
    L.orc_resolve_rgb.argtypes = [vp, C.c_uint64, C.c_float, vp]
    L.orc_film_add_samples.argtypes = [vp, vp, vp, vp, C.c_uint64, vp]
    L.orc_roughness_to_alpha.restype = C.c_float
    L.orc_roughness_to_alpha.argtypes = [C.c_float]
    L.orc_sincos.argtypes = [C.c_float, vp, vp]
    L.orc_film_table.argtypes = [vp, vp]
    L.orc_film_add_splats.argtypes = [vp, vp, vp, C.c_uint64, vp]
    L.orc_resolve_rgb_splat.argtypes = [vp, vp, C.c_uint64, C.c_float, C.c_float, vp]
    L.orc_rgb_to_xyz.argtypes = [vp, C.c_uint64, vp]
    L.orc_spatial_grid.argtypes = [vp, vp]
    L.orc_spatial_voxel.argtypes = [vp, vp, vp, vp, vp]
    L.orc_spatial_voxel_of.argtypes = [vp, vp, vp]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


PCG32_DEFAULT_STATE = 0x853C49E6748FEA9B


def pcg32_u32(sequence, n, init_state=PCG32_DEFAULT_STATE):
    out = np.empty(n, dtype=np.uint32)
    lib().orc_pcg32_u32(sequence, init_state, n, _p(out))
    return out


def pcg32_float(sequence, n):
    out = np.empty(n, dtype=np.float32)
    lib().orc_pcg32_float(sequence, n, _p(out))
    return out


def gamma(n):
    return np.float32(lib().orc_gamma(n))


def next_float_up(v):
    return np.float32(lib().orc_next_float_up(v))


def next_float_down(v):
    return np.float32(lib().orc_next_float_down(v))


def make_ray(o, d, t_max=np.inf, time=0.0):
    return np.array([o[0], o[1], o[2], t_max, d[0], d[1], d[2], time], dtype=np.float32)


def slab_test(bounds6, ray8):
    t = C.c_float(0)
    ok = lib().orc_slab_test(_p(_f32(bounds6)), _p(_f32(ray8)), C.byref(t))
    return bool(ok), np.float32(t.value)


def triangle_test(tri9, ray8):
    out = np.zeros(4, dtype=np.float32)
    ok = lib().orc_triangle_test(_p(_f32(tri9).ravel()), _p(_f32(ray8)), _p(out))
    return bool(ok), out  # b0 b1 b2 t


class BVHAccel:
    """Oracle restatement of src/accelerators/bvh.rs BVHAccel (split_method 0 = SAH, 1 = HLBVH)."""

    def __init__(self, verts, idx, max_prims_in_node=4, _handle=None, split_method=0):
        self._own = _handle is None
        if _handle is not None:
            self.h = _handle
            return
        self.verts = _f32(verts).reshape(-1, 3)
        self.idx = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1, 3)
        self.h = lib().orc_bvh_build2(_p(self.verts), len(self.verts), _p(self.idx), len(self.idx), max_prims_in_node, split_method)

    def __del__(self):
        if getattr(self, "h", None) and self._own:
            lib().orc_bvh_free(self.h)
            self.h = None

    @property
    def num_nodes(self):
        return int(lib().orc_bvh_num_nodes(self.h))

    @property
    def max_depth(self):
        return int(lib().orc_bvh_max_depth(self.h))

    def nodes(self):
        out = np.empty(self.num_nodes, dtype=NODE_DTYPE)
        lib().orc_bvh_get_nodes(self.h, _p(out))
        return out

    def ordered_prims(self):
        out = np.empty(int(lib().orc_bvh_num_prims(self.h)), dtype=np.uint32)
        lib().orc_bvh_get_ordered_prims(self.h, _p(out))
        return out

    def world_bound(self):
        out = np.empty(6, dtype=np.float32)
        lib().orc_bvh_world_bound(self.h, _p(out))
        return out

    def intersect(self, rays, threads=0, counters=False, want_b0=False):
        rays = _f32(rays).reshape(-1, 8)
        hits = np.empty(len(rays), dtype=HIT_DTYPE)
        cnt = np.zeros(2, dtype=np.uint64)
        b0 = np.empty(len(rays), dtype=np.float32) if want_b0 else None
        dt = lib().orc_intersect(self.h, _p(rays), len(rays), _p(hits), _p(b0) if want_b0 else None,
                                 _p(cnt) if counters else None, threads)
        res = [hits]
        if want_b0:
            res.append(b0)
        if counters:
            res.append(cnt)
        res.append(dt)
        return tuple(res)

    def intersect_p(self, rays, threads=0, counters=False):
        rays = _f32(rays).reshape(-1, 8)
        out = np.empty(len(rays), dtype=np.uint8)
        cnt = np.zeros(2, dtype=np.uint64)
        dt = lib().orc_intersect_p(self.h, _p(rays), len(rays), _p(out), _p(cnt) if counters else None, threads)
        return (out, cnt, dt) if counters else (out, dt)

    def brute_force(self, rays, threads=0):
        rays = _f32(rays).reshape(-1, 8)
        hits = np.empty(len(rays), dtype=HIT_DTYPE)
        lib().orc_brute_force(self.h, _p(rays), len(rays), _p(hits), threads)
        return hits


def _cam9(pos, look, up):
    return np.array(list(pos) + list(look) + list(up), dtype=np.float32)


def camera_matrices(pos, look, up, fov, res):
    r2c = np.empty((4, 4), dtype=np.float32)
    c2w = np.empty((4, 4), dtype=np.float32)
    lib().orc_camera_matrices(_p(_cam9(pos, look, up)), fov, res[0], res[1], _p(r2c), _p(c2w))
    return r2c, c2w


def camera_primary_rays(pos, look, up, fov, res):
    rays = np.empty((res[0] * res[1], 8), dtype=np.float32)
    lib().orc_camera_primary_rays(_p(_cam9(pos, look, up)), fov, res[0], res[1], _p(rays))
    return rays


def camera_rays(pos, look, up, fov, res, pfilm, plens=None, lens_radius=0.0, focal_distance=1e6):
    pfilm = _f32(pfilm).reshape(-1, 2)
    rays = np.empty((len(pfilm), 8), dtype=np.float32)
    if plens is None:
        lib().orc_camera_rays(_p(_cam9(pos, look, up)), fov, res[0], res[1], _p(pfilm), len(pfilm), _p(rays))
    else:
        plens = _f32(plens).reshape(-1, 2)
        lib().orc_camera_rays_lens(_p(_cam9(pos, look, up)), fov, res[0], res[1], lens_radius, focal_distance, _p(pfilm), _p(plens),
                                   len(pfilm), _p(rays))
    return rays


def spawn_shadow_rays(bvh, hits, b0, light_pos):
    L = lib()
    L.orc_spawn_shadow_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    out = np.empty((len(hits), 8), dtype=np.float32)
    lp = np.asarray(light_pos, dtype=np.float32)
    L.orc_spawn_shadow_rays(bvh.h, _p(np.ascontiguousarray(hits)), _p(_f32(b0)), len(hits), _p(lp), _p(out))
    return out


def spawn_bounce_rays(bvh, rays, hits, b0):
    L = lib()
    L.orc_spawn_bounce_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    out = np.empty((len(hits), 8), dtype=np.float32)
    L.orc_spawn_bounce_rays(bvh.h, _p(_f32(rays)), _p(np.ascontiguousarray(hits)), _p(_f32(b0)), len(hits), _p(out))
    return out
