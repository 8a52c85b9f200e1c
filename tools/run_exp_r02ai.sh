# Round-2 batch AI: per-warp prepared-ray pool (refill = shared-memory loads) against the in-refill set-up, refill thresholds swept.
set -x
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "not full_size" > $O/r02ai_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02ai_pytest.log
tail -3 $O/r02ai_pytest.log
TUNE_C2_SPP=64 TUNE_C4_SPP=32 python tools/tune_path.py 12,16,20,24,28 8,12 18 2>/dev/null | sed "s/^/pool /" >> $O/r02ai_frames.log
PB2_LIB=$PWD/build/libpbrt_b200_nopool.so TUNE_C2_SPP=64 TUNE_C4_SPP=32 python tools/tune_path.py 12 8 18 2>/dev/null | sed "s/^/nopool /" >> $O/r02ai_frames.log
python tools/tune_trace.py 12,16,20,24,28 8,12 0 18 2>/dev/null | sed "s/^/pool /" >> $O/r02ai_frames.log
PB2_LIB=$PWD/build/libpbrt_b200_nopool.so python tools/tune_trace.py 12 8 0 18 2>/dev/null | sed "s/^/nopool /" >> $O/r02ai_frames.log
cat $O/r02ai_frames.log
