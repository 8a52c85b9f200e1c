# Round-2 final profiling batch 2: launch list of the bench command (C3 steps), ncu --set full of the three C3 traversal launches
# (DRAM bytes + lanes per instruction -> profiles/traffic.json), launch list + DRAM bytes of one C4 / C2 batch (traffic.json path.*).
set -x
O=gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-path > $O/r02ah_bench_nopath.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02ah_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-path > $O/r02ah_bench_ncu.log 2>&1
ncu --set full --clock-control none -k regex:'k_closest_hit|k_any_hit' -c 3 -f -o /tmp/r02ah_c3_trace python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-path > $O/r02ah_c3_trace.log 2>&1
ncu -i /tmp/r02ah_c3_trace.ncu-rep --page raw --csv 2>/dev/null | gzip > $O/r02ah_c3_trace_raw.csv.gz
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r02ah_c4_launches.csv python tools/prof_path.py --scene c4 --spp 8 > $O/r02ah_c4_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r02ah_c2_launches.csv python tools/prof_path.py --scene c2 --spp 32 > $O/r02ah_c2_ncu.log 2>&1
ls -la $O/r02ah*
