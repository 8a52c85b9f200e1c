set -x
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "not full_size" > $O/r02as_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02as_pytest.log
tail -3 $O/r02as_pytest.log
for rep in 1 2; do TUNE_C2_SPP=64 TUNE_C4_SPP=32 python tools/tune_path.py 12 8 18 2>/dev/null | sed "s/^/fastdiv /" >> $O/r02as_frames.log; done
cat $O/r02as_frames.log
