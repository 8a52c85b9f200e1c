# Round-2 batch S: pb2_scene_wait_until (one step in flight in the e2e loop) — tests + C3-only bench line.
set -x
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_raycast.py -m gpu -x -q -k "not full_size" > $O/r02s_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02s_pytest.log
tail -3 $O/r02s_pytest.log
for rep in 1 2; do python bench.py --no-path --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('C3', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'sync', round(d['e2e']['synchronous_calls_mrays_s'], 1), 'link bound', round(d['e2e']['link_bound_mrays_s'], 1), d['parity'])
" | tee -a $O/r02s_e2e.log; done
