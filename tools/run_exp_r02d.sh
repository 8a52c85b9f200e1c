# Round-2 experiment batch D: triangle records in primitive order for k_shade (PB2_TRIS_BY_PRIM 1 vs 0).
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q -k "path or sphere or volpath or golden" > $O/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02d_pytest.log
for rep in 1 2; do
  for v in main noprim; do
    if [ $v = main ]; then unset PB2_LIB; else export PB2_LIB=$PWD/build/libpbrt_b200_$v.so; fi
    python tools/tune_path.py 12 8 18 2>/dev/null | sed "s/^/$v /" >> $O/r02d_tris_prim.log
  done
done
unset PB2_LIB
cat $O/r02d_tris_prim.log; tail -3 $O/r02d_pytest.log
