# Round-2 batch X: + prefetch of the next vertex's triangle record in k_shade.
set -x
O=gpurun_out
for rep in 1 2; do
  TUNE_C2_SPP=64 TUNE_C4_SPP=32 python tools/tune_path.py 12 8 18 2>/dev/null | sed "s/^/pf_state+tri /" >> $O/r02x_frames.log
  PB2_LIB=$PWD/build/libpbrt_b200_pfl2only.so TUNE_C2_SPP=64 TUNE_C4_SPP=32 python tools/tune_path.py 12 8 18 2>/dev/null | sed "s/^/pf_state     /" >> $O/r02x_frames.log
done
cat $O/r02x_frames.log
