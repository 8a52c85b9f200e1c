# Round-2 batch AK: two node steps per vote while few lanes hold a leaf (variant reps2) against one (main).
set -x
O=gpurun_out
PB2_LIB=$PWD/build/libpbrt_b200_reps2.so timeout 900 python -m pytest tests/test_gpu_raycast.py -m gpu -x -q -k "not full_size" > $O/r02ak_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02ak_pytest.log
tail -2 $O/r02ak_pytest.log
for rep in 1 2; do
  for v in main reps2; do
    if [ $v = main ]; then unset PB2_LIB; else export PB2_LIB=$PWD/build/libpbrt_b200_$v.so; fi
    TUNE_C2_SPP=64 TUNE_C4_SPP=32 python tools/tune_path.py 12 8 18 2>/dev/null | sed "s/^/$v /" >> $O/r02ak_frames.log
    python tools/tune_trace.py 12 8 0 18 2>/dev/null | sed "s/^/$v /" >> $O/r02ak_frames.log
  done
done
cat $O/r02ak_frames.log
