# Round-2 experiment batch C: k_shade loop-head software pipeline (PB2_SHADE_PIPE 1 vs 0).
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q -k "not full_size" > $O/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02c_pytest.log
for rep in 1 2; do
  for v in main nopipe; do
    if [ $v = main ]; then unset PB2_LIB; else export PB2_LIB=$PWD/build/libpbrt_b200_$v.so; fi
    python tools/tune_path.py 12 8 18 2>/dev/null | sed "s/^/$v /" >> $O/r02c_shade_pipe.log
  done
done
unset PB2_LIB
cat $O/r02c_shade_pipe.log; tail -3 $O/r02c_pytest.log
