# Round-2 batch N: per-primitive frame in the SG kernels (scenes with analytic spheres, no mesh attributes).
set -x
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "not full_size" > $O/r02n_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02n_pytest.log
tail -3 $O/r02n_pytest.log
python - > $O/r02n_spheres.log 2>&1 <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import __graft_entry__ as ge, torch, numpy as np
pb2, scenes = ge.load_package(), ge.load_scenes()
pb2.init(0)
st = torch.cuda.current_stream().cuda_stream
sc = scenes.scene_spheres(); cam = dict(scenes.C2_CAMERA, res=(1024, 1024))
accel = pb2.BVHAccel(pb2.scene_from_dict(sc), max_prims_in_node=4)
camera = pb2.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
integ = pb2.PathIntegrator(accel, camera, spp=16, max_depth=5, rr_threshold=1.0, light_strategy="power")
film = pb2.Film(cam["res"])
integ.render(film, 0, 2, stream=st); torch.cuda.synchronize()
for rep in range(3):
    film.clear()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); integ.render(film, stream=st); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print(f"spheres 1024x1024 @ 16 spp: {ms:.3f} ms = {1024*1024*16/ms/1e3:.1f} Msamples/s rgb {film.resolve_rgb().mean():.6f}")
PY
cat $O/r02n_spheres.log
