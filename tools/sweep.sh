#!/bin/bash
# tools/sweep.sh NAME...  — bench.py (no CPU baseline) once per tuning build build/libpbrt_b200_NAME.so ("base" = the in-tree library)
for v in "$@"; do
  if [ "$v" = base ]; then unset PB2_LIB; else export PB2_LIB=$PWD/build/libpbrt_b200_$v.so; fi
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/sweep_$v.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('$v', 'value', round(d['value'],1), {k:round(x,4) for k,x in d['kernel_ms'].items()}, 'crc', d['hits_crc32'], 'path', round(d['path']['value'],1), 'Msamples/s', round(d['path']['ms_per_frame'],2),'ms', 'rgb', d['path']['mean_rgb'])
"
done
