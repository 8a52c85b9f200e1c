# Round-2 experiment batch H: per-primitive shading frame + per-light constants + branch-free offset_ray_origin step.
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q -k "not full_size" > $O/r02h_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02h_pytest.log
for rep in 1 2; do
  python tools/tune_path.py 12 8 18 2>/dev/null | sed "s/^/frames /" >> $O/r02h_frames.log
done
cat $O/r02h_frames.log; tail -3 $O/r02h_pytest.log
ncu --set full --clock-control none --import-source on -k regex:'k_shade' --launch-skip 3 -c 1 -f -o /tmp/r02h_shade3 python tools/prof_path.py --scene c4 --spp 8 > /dev/null 2>&1
ncu -i /tmp/r02h_shade3.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r02h_c4_shade3_sass.csv.gz
ncu -i /tmp/r02h_shade3.ncu-rep --page raw --csv 2>/dev/null | gzip > $O/r02h_c4_shade3_raw.csv.gz
