set -x
O=gpurun_out
for rep in 1 2; do
  for v in main s128x5 s128x6; do
    if [ $v = main ]; then unset PB2_LIB; else export PB2_LIB=$PWD/build/libpbrt_b200_$v.so; fi
    TUNE_C2_SPP=64 TUNE_C4_SPP=32 python tools/tune_path.py 12 8 18 2>/dev/null | sed "s/^/$v /" >> $O/r02ad_frames.log
  done
done
cat $O/r02ad_frames.log
