# Profiling batch of round 2 (run on the GPU box through gpurun): ray-sort experiment, C4 launch list, ncu --set full of the
# wavefront kernels of one C4 batch and of the three C3 traversal launches.  .ncu-rep files are exported to CSV on the box
# and only the CSVs (gzip) travel back (gpurun_out/ is capped at 64 MiB).
set -x
O=gpurun_out
python tools/exp_raysort.py > $O/r02_exp_raysort.log 2>&1
python tools/prof_path.py --scene c4 --spp 8 > $O/r02_prof_c4_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_c4_launches.csv python tools/prof_path.py --scene c4 --spp 8 > $O/r02_c4_ncu.log 2>&1
ncu --set full --clock-control none -k regex:'k_extend|k_shade|k_shadow|k_select3' -c 21 -f -o /tmp/r02_c4_full python tools/prof_path.py --scene c4 --spp 8 > $O/r02_c4_full.log 2>&1
ncu -i /tmp/r02_c4_full.ncu-rep --page raw --csv 2>/dev/null | gzip > $O/r02_c4_full_raw.csv.gz
# source-level view of the second-bounce k_extend and k_shade<0> (launches 7 and 9 of the filtered list)
ncu --set full --clock-control none --import-source on -k regex:'k_extend' --launch-skip 1 -c 1 -f -o /tmp/r02_c4_extend python tools/prof_path.py --scene c4 --spp 8 > /dev/null 2>&1
ncu -i /tmp/r02_c4_extend.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r02_c4_extend_source.csv.gz
ncu --set full --clock-control none --import-source on -k regex:'k_shade' --launch-skip 3 -c 1 -f -o /tmp/r02_c4_shade python tools/prof_path.py --scene c4 --spp 8 > /dev/null 2>&1
ncu -i /tmp/r02_c4_shade.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r02_c4_shade_source.csv.gz
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-path > $O/r02_bench_nopath.log 2>&1 || exit 1
ncu --set full --clock-control none -k regex:'k_closest_hit|k_any_hit' -c 3 -f -o /tmp/r02_c3_trace python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-path > $O/r02_c3_trace.log 2>&1
ncu -i /tmp/r02_c3_trace.ncu-rep --page raw --csv 2>/dev/null | gzip > $O/r02_c3_trace_raw.csv.gz
ls -la $O /tmp/*.ncu-rep
