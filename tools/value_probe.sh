#!/bin/bash
# what bounds the resident stream?  `value` against the number of frames in flight, and with the warp-per-block
# coders on the small configurations
mkdir -p gpurun_out
for w in 2 4 8 12 16; do
  echo "== workers $w" >> gpurun_out/value_probe.log
  OJB_BENCH_WORKERS=$w OJB_BENCH_EXTRAS=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>>gpurun_out/value_probe.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); t=d['detail']
print('value %.0f e2e %.0f serial %.2f ms  enc %s dec %s' % (d['value'], d['e2e']['value'], t['serial_ms_per_frame'], t['stages_encode_ms'], t['stages_decode_ms']))" >> gpurun_out/value_probe.log
done
cat gpurun_out/value_probe.log
