# compute-sanitizer memcheck over one small frame of each integrator (new kernels of the round: k_select3, k_vol_*, two arenas).
set -x
O=gpurun_out
cat > /tmp/mc.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import __graft_entry__ as ge
pb2, scenes = ge.load_package(), ge.load_scenes()
pb2.init(0)
for name, sc, kw in (("path", scenes.scene_c2(), dict(max_depth=4, rr_threshold=1.0, light_strategy="uniform", spp=4)),
                     ("volpath", scenes.scene_media(), dict(max_depth=4, rr_threshold=1.0, light_strategy="power", spp=4, integrator="volpath"))):
    cam = dict(scenes.C2_CAMERA, res=(48, 48))
    accel = pb2.BVHAccel(pb2.scene_from_dict(sc), max_prims_in_node=4)
    camera = pb2.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
    integ = pb2.PathIntegrator(accel, camera, **kw)
    film = pb2.Film(cam["res"])
    integ.render(film)
    print(name, float(film.resolve_rgb().mean()))
PY
PB2_WAVEFRONT_LOG2_SLOTS=16 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/mc.py > $O/r02_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/r02_memcheck.log
tail -12 $O/r02_memcheck.log
