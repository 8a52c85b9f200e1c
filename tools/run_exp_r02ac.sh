set -x
O=gpurun_out
for rep in 1 2; do
  python tools/bench_volpath.py 2>&1 | grep -v Warning | sed "s/^/blocks3 /" >> $O/r02ac_vol.log
  PB2_LIB=$PWD/build/libpbrt_b200_vol4.so python tools/bench_volpath.py 2>&1 | grep -v Warning | sed "s/^/blocks4 /" >> $O/r02ac_vol.log
done
cat $O/r02ac_vol.log
