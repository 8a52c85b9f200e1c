# Round-2 final profiling batch: ncu --set full of the wavefront kernels of one C4 batch (first three bounces) and of the
# volpath stages of the fog + smoke scene, source-level view of the second k_extend launch.  .ncu-rep -> CSV on the box.
set -x
O=gpurun_out
python tools/prof_path.py --scene c4 --spp 8 > $O/r02t_prof_c4_plain.log 2>&1 || exit 1
ncu --set full --clock-control none -k regex:'k_extend|k_shade|k_shadow|k_select3' -c 21 -f -o /tmp/r02t_c4_full python tools/prof_path.py --scene c4 --spp 8 > $O/r02t_c4_full.log 2>&1
ncu -i /tmp/r02t_c4_full.ncu-rep --page raw --csv 2>/dev/null | gzip > $O/r02t_c4_full_raw.csv.gz
ncu --set full --clock-control none --import-source on -k regex:'k_extend' --launch-skip 1 -c 1 -f -o /tmp/r02t_c4_extend python tools/prof_path.py --scene c4 --spp 8 > /dev/null 2>&1
ncu -i /tmp/r02t_c4_extend.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r02t_c4_extend_source.csv.gz
python tools/prof_path.py --scene media --spp 8 > $O/r02t_prof_media_plain.log 2>&1 || exit 1
ncu --set full --clock-control none -k regex:'k_vol_' -c 24 -f -o /tmp/r02t_media_full python tools/prof_path.py --scene media --spp 8 > $O/r02t_media_full.log 2>&1
ncu -i /tmp/r02t_media_full.ncu-rep --page raw --csv 2>/dev/null | gzip > $O/r02t_media_full_raw.csv.gz
ls -la $O/r02t* /tmp/*.ncu-rep
