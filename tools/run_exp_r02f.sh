# Round-2 experiment batch F: zero-numerator divide shortcut (PB2_DIV_ZN 1 vs 0) + a fresh source-level ncu capture of
# k_shade<0> / k_shade<1> at bounce 1 of a C4 batch (the kernels as they are after the loop-head pipeline).
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q -k "not full_size" > $O/r02f_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02f_pytest.log
for rep in 1 2; do
  for v in main nodivzn; do
    if [ $v = main ]; then unset PB2_LIB; else export PB2_LIB=$PWD/build/libpbrt_b200_$v.so; fi
    python tools/tune_path.py 12 8 18 2>/dev/null | sed "s/^/$v /" >> $O/r02f_divzn.log
  done
done
unset PB2_LIB
cat $O/r02f_divzn.log; tail -3 $O/r02f_pytest.log
# k_shade launches of one batch in order: bounce 0 <0>,<1>,<2>, bounce 1 <0>,<1>,<2> ... -> skip 3 = bounce 1 <0>, skip 4 = bounce 1 <1>
for k in 3 4; do
  ncu --set full --clock-control none --import-source on -k regex:'k_shade' --launch-skip $k -c 1 -f -o /tmp/r02f_shade$k python tools/prof_path.py --scene c4 --spp 8 > /dev/null 2>&1
  ncu -i /tmp/r02f_shade$k.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r02f_c4_shade${k}_sass.csv.gz
  ncu -i /tmp/r02f_shade$k.ncu-rep --page source --csv --print-source cuda 2>/dev/null | gzip > $O/r02f_c4_shade${k}_cuda.csv.gz
  ncu -i /tmp/r02f_shade$k.ncu-rep --page raw --csv 2>/dev/null | gzip > $O/r02f_c4_shade${k}_raw.csv.gz
done
ls -la $O | tail -8
