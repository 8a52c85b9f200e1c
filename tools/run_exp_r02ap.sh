# Round-2 batch AP: traversal kernels at 9 / 10 CTAs per SM (56 / 51 registers) against 8 (64).
set -x
O=gpurun_out
for rep in 1 2; do
  for v in main mb9 mb10; do
    if [ $v = main ]; then unset PB2_LIB; else export PB2_LIB=$PWD/build/libpbrt_b200_$v.so; fi
    TUNE_C2_SPP=64 TUNE_C4_SPP=32 python tools/tune_path.py 12 8 18 2>/dev/null | sed "s/^/$v /" >> $O/r02ap_frames.log
    python tools/tune_trace.py 12 8 0 18 2>/dev/null | sed "s/^/$v /" >> $O/r02ap_frames.log
  done
done
cat $O/r02ap_frames.log
