# Round-2 batch AB: k_select3 at 4 CTAs per SM (64 registers); the multi-batch two-stream test.
set -x
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "not full_size" > $O/r02ab_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02ab_pytest.log
tail -3 $O/r02ab_pytest.log
for rep in 1 2; do
  TUNE_C2_SPP=64 TUNE_C4_SPP=32 python tools/tune_path.py 12 8 18 2>/dev/null | sed "s/^/select_4cta /" >> $O/r02ab_frames.log
done
python tools/bench_volpath.py 2>&1 | grep -v Warning >> $O/r02ab_frames.log
cat $O/r02ab_frames.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02ab_c4_launches.csv python tools/prof_path.py --scene c4 --spp 8 > /dev/null 2>&1
python tools/launch_share.py $O/r02ab_c4_launches.csv k_raygen 2>/dev/null | head -8
