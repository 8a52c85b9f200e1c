set -x
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_volpath.py tests/test_gpu_path.py -m gpu -x -q -k "volpath or media or many_small" > $O/r02ar_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02ar_pytest.log
tail -2 $O/r02ar_pytest.log
for rep in 1 2; do
  python tools/bench_volpath.py 2>&1 | grep -v Warning | sed "s/^/blind1+cached /" >> $O/r02ar_vol.log
  PB2_LIB=$PWD/build/libpbrt_b200_blind0.so python tools/bench_volpath.py 2>&1 | grep -v Warning | sed "s/^/blind0+cached /" >> $O/r02ar_vol.log
done
cat $O/r02ar_vol.log
