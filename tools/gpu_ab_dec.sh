#!/bin/bash
# A/B of the decoder variants + one ncu capture of the thread-per-block decoder (instruction counts)
mkdir -p gpurun_out
OJB_BLOCK_DECODER=twostep timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_gpu_twostep.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_twostep.log
tail -3 gpurun_out/pytest_gpu_twostep.log
for v in twostep serialdec; do
  if [ $v = twostep ]; then export OJB_BLOCK_DECODER=twostep; else unset OJB_BLOCK_DECODER; fi
  OJB_BENCH_EXTRAS=${EXTRAS:-1} timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python - <<PY
import json
r=json.load(open("gpurun_out/bench_$v.json")); c=r["config"]
print("$v value %.0f e2e %.0f serial %.2f ms enc %s dec %s" % (r["value"], r["e2e"]["value"], c["serial_ms_per_frame"], c["stages_encode_ms"], c["stages_decode_ms"]))
x=c.get("irv97_ict_q90")
if x: print("   9/7:", x.get("Mpixels_per_s"), x.get("stages_encode_ms"), x.get("stages_decode_ms"))
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ht_decode_serial" -c 2 \
   -o gpurun_out/prof_serialdec python tools/profile_once.py > gpurun_out/ncu_serialdec.log 2>&1
tail -2 gpurun_out/ncu_serialdec.log
