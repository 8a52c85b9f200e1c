# Round-2 batch L: DRAM traffic of one C4 / C2 batch (ncu launch list with dram bytes -> profiles/traffic.json "path"), then the
# default bench line of the current tree.
set -x
O=gpurun_out
python tools/prof_path.py --scene c4 --spp 8 > $O/r02l_prof_c4_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r02l_c4_launches.csv python tools/prof_path.py --scene c4 --spp 8 > $O/r02l_c4_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r02l_c2_launches.csv python tools/prof_path.py --scene c2 --spp 32 > $O/r02l_c2_ncu.log 2>&1
python tools/ncu_path_traffic.py $O/r02l_c4_launches.csv c4 $((1920*1080*8)) > $O/r02l_traffic_c4.log 2>&1
python tools/ncu_path_traffic.py $O/r02l_c2_launches.csv c2 $((512*512*32)) > $O/r02l_traffic_c2.log 2>&1
cp profiles/traffic.json $O/r02l_traffic.json
python bench.py > $O/r02l_bench.json 2> $O/r02l_bench.err; echo "bench rc=$?" >> $O/r02l_bench.err
tail -3 $O/r02l_bench.err
