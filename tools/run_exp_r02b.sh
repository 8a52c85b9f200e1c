# Round-2 experiment batch B: deferred result hand-off in trace_persistent (PB2_DEFER_FINISH 1 vs 0), pixel-tile ray orders.
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q -k "not full_size" > $O/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02b_pytest.log
for rep in 1 2; do
  for v in main nodefer; do
    if [ $v = main ]; then unset PB2_LIB; else export PB2_LIB=$PWD/build/libpbrt_b200_$v.so; fi
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-path 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['value'],1), {k: round(x,4) for k,x in d['kernel_ms'].items()}, d['hits_crc32'])" >> $O/r02b_defer.log
    python tools/tune_path.py 12 8 18 2>/dev/null | sed "s/^/$v /" >> $O/r02b_defer.log
  done
done
unset PB2_LIB
python tools/exp_raysort.py > $O/r02b_exp_raysort.log 2>&1
cat $O/r02b_defer.log
