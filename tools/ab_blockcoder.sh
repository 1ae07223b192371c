#!/bin/bash
# A/B of the block-encoder variants (thread-per-block default vs warp-per-block) and of the decoder variants
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for v in serial warp twostepdec; do
  if [ $v = twostepdec ]; then export OJB_BLOCK_ENCODER=serial OJB_BLOCK_DECODER=twostep; else export OJB_BLOCK_ENCODER=$v; unset OJB_BLOCK_DECODER; fi
  OJB_BENCH_EXTRAS=${EXTRAS:-0} timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python - <<PY
import json
r=json.load(open("gpurun_out/bench_$v.json")); c=r["config"]
print("$v value %.0f e2e %.0f serial %.2f ms enc %s dec %s" % (r["value"], r["e2e"]["value"], c["serial_ms_per_frame"], c["stages_encode_ms"], c["stages_decode_ms"]))
x=c.get("irv97_ict_q90")
if x: print("   9/7:", x.get("Mpixels_per_s"), x.get("stages_encode_ms"), x.get("stages_decode_ms"))
PY
done
