# Round-2 batch K: two batches in flight (Wavefront::peer) against PB2_TWO_STREAMS=0.
set -x
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02k_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02k_pytest.log
tail -5 $O/r02k_pytest.log
for rep in 1 2; do
  TUNE_C2_SPP=64 TUNE_C4_SPP=32 python tools/tune_path.py 12 8 18 2>/dev/null | sed "s/^/two-streams /" >> $O/r02k_frames.log
  PB2_TWO_STREAMS=0 TUNE_C2_SPP=64 TUNE_C4_SPP=32 python tools/tune_path.py 12 8 18 2>/dev/null | sed "s/^/one-stream  /" >> $O/r02k_frames.log
done
cat $O/r02k_frames.log
python tools/bench_volpath.py 2>&1 | grep -v Warning | tee -a $O/r02k_frames.log
