#!/usr/bin/env python
"""Development check for the QuadNode collapse (CPU only): quad walk == reference-order binary walk, plus step counts."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

so = os.path.join(ROOT, "build", "quad_sim.so")
os.makedirs(os.path.dirname(so), exist_ok=True)
subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-o", so,
                       os.path.join(ROOT, "tools", "quad_sim.cpp"), os.path.join(ROOT, "pbrt-rs_b200", "csrc", "bvh_build.cpp")])
L = C.CDLL(so)
scenes, orc = ge.load_scenes(), ge.load_oracle()


def run(name, verts, idx, rays):
    verts = np.ascontiguousarray(verts, np.float32)
    idx = np.ascontiguousarray(idx, np.uint32)
    rays = np.ascontiguousarray(rays, np.float32)
    st = np.zeros(8, np.uint64)
    vp = C.c_void_p
    L.quad_sim(verts.ctypes.data_as(vp), C.c_uint64(len(verts)), idx.ctypes.data_as(vp), C.c_uint64(len(idx)), 4,
               rays.ctypes.data_as(vp), C.c_uint64(len(rays)), st.ctypes.data_as(vp))
    n = len(rays)
    print(f"{name}: rays {n} binary nodes/ray {st[0]/n:.2f} tris/ray {st[1]/n:.3f} | quad steps/ray {st[2]/n:.2f} boxes/ray {st[3]/n:.2f} "
          f"tris/ray {st[4]/n:.3f} | mismatches {st[5]} nonplain {st[6]} quads {int(st[7]) >> 8} max_sp {int(st[7]) & 255}")
    assert st[5] == 0 and st[1] == st[4]
    w = np.zeros(13, np.uint64)
    L.wide_sim(verts.ctypes.data_as(vp), C.c_uint64(len(verts)), idx.ctypes.data_as(vp), C.c_uint64(len(idx)), 4,
               rays.ctypes.data_as(vp), C.c_uint64(len(rays)), w.ctypes.data_as(vp))
    for lv, o in ((2, 0), (3, 6)):
        print(f"   {lv}-level record ({1 << lv}-wide): steps/ray {w[o]/n:.2f} boxes/ray {w[o+1]/n:.2f} tris/ray {w[o+2]/n:.3f} pushes/ray {w[o+3]/n:.2f} "
              f"pops/ray {w[o+4]/n:.2f} leaf visits/ray {w[o+5]/n:.2f}")
    print(f"   3-level / 2-level: steps x{w[6]/w[0]:.3f} boxes x{w[7]/w[1]:.3f} tris x{w[8]/max(1, w[2]):.3f}; mismatches vs binary {w[12]}")
    assert w[12] == 0


which = sys.argv[1] if len(sys.argv) > 1 else "c1"
if which == "c1":
    v, i = scenes.scene_c1()
    cam = scenes.C1_CAMERA
elif which == "c3":
    v, i = scenes.scene_c3(int(sys.argv[2]) if len(sys.argv) > 2 else 2237)
    cam = dict(scenes.C3_CAMERA, res=(1024, 1024))
elif which == "c2":
    sc = scenes.scene_c2(); v, i = sc["verts"], sc["idx"]; cam = scenes.C2_CAMERA
elif which == "c4":
    sc = scenes.scene_c4(); v, i = sc["verts"], sc["idx"]; cam = dict(scenes.C4_CAMERA, res=(960, 540))
elif which == "soup":
    v, i = scenes.random_soup(20000, seed=3); cam = scenes.C1_CAMERA
rays = orc.camera_primary_rays(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
run(which + " primary", v, i, rays)
ref = orc.BVHAccel(v, i, 4)
hits, b0, _ = ref.intersect(rays, want_b0=True)
br = orc.spawn_bounce_rays(ref, rays, hits, b0)
run(which + " bounce", v, i, br)
