# Round-2 batch Y: L2 prefetch of the ray PB2_RAY_AHEAD indices ahead at every refill (main) against none.
set -x
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "not full_size" > $O/r02y_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02y_pytest.log
tail -3 $O/r02y_pytest.log
for rep in 1 2; do
  for v in ahead noahead; do
    if [ $v = noahead ]; then export PB2_LIB=$PWD/build/libpbrt_b200_noahead.so; else unset PB2_LIB; fi
    TUNE_C2_SPP=64 TUNE_C4_SPP=32 python tools/tune_path.py 12 8 18 2>/dev/null | sed "s/^/$v /" >> $O/r02y_frames.log
    python bench.py --no-path --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('$v C3', round(d['value'], 1), {k: round(v, 4) for k, v in d['kernel_ms'].items()}, d['hits_crc32'], 'e2e', round(d['e2e']['value'], 1))
" >> $O/r02y_frames.log
  done
done
cat $O/r02y_frames.log
