# Round-2 batch I: wavefront VolPathIntegrator against the one-thread-per-path kernel.
set -x
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_volpath.py -m gpu -x -q > $O/r02i_pytest_vol.log 2>&1; echo "pytest rc=$?" >> $O/r02i_pytest_vol.log
tail -15 $O/r02i_pytest_vol.log
for rep in 1 2; do
  timeout 300 python tools/bench_volpath.py 2>&1 | grep -v Warning >> $O/r02i_volpath.log
  PB2_VOLPATH_MEGAKERNEL=1 timeout 300 python tools/bench_volpath.py 2>&1 | grep -v Warning >> $O/r02i_volpath.log
done
cat $O/r02i_volpath.log
