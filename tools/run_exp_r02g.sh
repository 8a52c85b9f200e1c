# Round-2 experiment batch G: plastic-only class-1 shade kernel (k_shade<3>) vs the general class-1 kernel (PB2_GENERAL_CLASS1=1).
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q -k "not full_size" > $O/r02g_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02g_pytest.log
for rep in 1 2; do
  for v in plastic general; do
    if [ $v = plastic ]; then unset PB2_GENERAL_CLASS1; else export PB2_GENERAL_CLASS1=1; fi
    python tools/tune_path.py 12 8 18 2>/dev/null | sed "s/^/$v /" >> $O/r02g_plastic.log
  done
done
unset PB2_GENERAL_CLASS1
cat $O/r02g_plastic.log; tail -3 $O/r02g_pytest.log
