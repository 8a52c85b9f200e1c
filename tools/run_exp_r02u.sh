# Round-2 batch U: rotation selects in tri_test.
set -x
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "not full_size" > $O/r02u_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02u_pytest.log
tail -3 $O/r02u_pytest.log
for rep in 1 2; do
  TUNE_C2_SPP=64 TUNE_C4_SPP=32 python tools/tune_path.py 12 8 18 2>/dev/null | sed "s/^/$1 /" >> $O/r02u_frames.log
  python bench.py --no-path --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('$1 C3', round(d['value'], 1), {k: round(v, 4) for k, v in d['kernel_ms'].items()}, d['hits_crc32'], 'e2e', round(d['e2e']['value'], 1))
" >> $O/r02u_frames.log
done
cat $O/r02u_frames.log
