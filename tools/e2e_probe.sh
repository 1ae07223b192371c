#!/bin/bash
# e2e throughput against the number of frames in flight (PCIe duplex overlap)
mkdir -p gpurun_out
for nw in ${NWS:-1 2 4 6}; do
  OJB_BENCH_EXTRAS=0 OJB_BENCH_WORKERS=$nw timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/e2e_nw$nw.json 2> gpurun_out/e2e_nw$nw.err
  python - <<PY
import json
r=json.load(open("gpurun_out/e2e_nw$nw.json"))
c=r["config"]
print("NW=$nw value %.0f e2e %.0f  e2e_enc %s e2e_dec %s" % (r["value"], r["e2e"]["value"], c["e2e_stages_encode_ms"], c["e2e_stages_decode_ms"]))
PY
done
