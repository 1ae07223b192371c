#!/bin/bash
# A/B of the block-decoder variants
mkdir -p gpurun_out
for v in warp destuff serial; do
  OJB_BLOCK_DECODER=$v OJB_BENCH_EXTRAS=${EXTRAS:-0} timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dec_$v.json 2> gpurun_out/bench_dec_$v.err
  python - <<PY
import json
r=json.load(open("gpurun_out/bench_dec_$v.json")); c=r["config"]
print("$v value %.0f e2e %.0f serial %.2f ms enc %s dec %s" % (r["value"], r["e2e"]["value"], c["serial_ms_per_frame"], c["stages_encode_ms"], c["stages_decode_ms"]))
PY
done
