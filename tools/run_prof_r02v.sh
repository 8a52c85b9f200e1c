# Round-2 batch V: source-level captures of k_shade<0> (bounce 1) and k_shadow (bounce 1) of one C4 batch, and of
# k_closest_hit on the C3 primary rays.
set -x
O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_shade' --launch-skip 3 -c 1 -f -o /tmp/r02v_shade0 python tools/prof_path.py --scene c4 --spp 8 > /dev/null 2>&1
ncu -i /tmp/r02v_shade0.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r02v_c4_shade0_source.csv.gz
ncu --set full --clock-control none --import-source on -k regex:'k_shadow' --launch-skip 1 -c 1 -f -o /tmp/r02v_shadow python tools/prof_path.py --scene c4 --spp 8 > /dev/null 2>&1
ncu -i /tmp/r02v_shadow.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r02v_c4_shadow_source.csv.gz
ncu --set full --clock-control none --import-source on -k regex:'k_extend' --launch-skip 1 -c 1 -f -o /tmp/r02v_extend python tools/prof_path.py --scene c4 --spp 8 > /dev/null 2>&1
ncu -i /tmp/r02v_extend.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r02v_c4_extend_source.csv.gz
ncu -i /tmp/r02v_extend.ncu-rep --page raw --csv 2>/dev/null | gzip > $O/r02v_c4_extend_raw.csv.gz
ncu --set full --clock-control none --import-source on -k regex:'k_closest_hit' -c 1 -f -o /tmp/r02v_c3 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-path > /dev/null 2>&1
ncu -i /tmp/r02v_c3.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r02v_c3_closest_source.csv.gz
ncu -i /tmp/r02v_c3.ncu-rep --page raw --csv 2>/dev/null | gzip > $O/r02v_c3_closest_raw.csv.gz
ls -la $O/r02v*
