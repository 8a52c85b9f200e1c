# Round-2 batch W: prefetch of the next vertex's path state in k_shade: L2 (main) / none / L1.
set -x
O=gpurun_out
for rep in 1 2; do
  TUNE_C2_SPP=64 TUNE_C4_SPP=32 python tools/tune_path.py 12 8 18 2>/dev/null | sed "s/^/pf_L2 /" >> $O/r02w_frames.log
  PB2_LIB=$PWD/build/libpbrt_b200_nopf.so TUNE_C2_SPP=64 TUNE_C4_SPP=32 python tools/tune_path.py 12 8 18 2>/dev/null | sed "s/^/no_pf /" >> $O/r02w_frames.log
  PB2_LIB=$PWD/build/libpbrt_b200_pfl1.so TUNE_C2_SPP=64 TUNE_C4_SPP=32 python tools/tune_path.py 12 8 18 2>/dev/null | sed "s/^/pf_L1 /" >> $O/r02w_frames.log
done
cat $O/r02w_frames.log
