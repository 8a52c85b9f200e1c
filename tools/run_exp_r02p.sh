# Round-2 batch P: wavefront size with two batches in flight (C4 1920x1080 @ 32 spp, C2 512x512 @ 64 spp).
set -x
O=gpurun_out
for lg in 22 23 24 25; do
  PB2_WAVEFRONT_LOG2_SLOTS=$lg TUNE_C2_SPP=64 TUNE_C4_SPP=32 python tools/tune_path.py 12 8 18 2>/dev/null | sed "s/^/log2_slots $lg /" >> $O/r02p_slots.log
done
cat $O/r02p_slots.log
