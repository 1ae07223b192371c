#!/bin/bash
# row regions on real GPUs: parity against the reference over NCCL, then the bench at N GPUs (its `sharded` block times
# the tile and the row-region partitions against one GPU)
mkdir -p gpurun_out
N=${NGPU:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/region_check.py > gpurun_out/region_n$N.json 2> gpurun_out/region_n$N.err; echo "region exit $?" >> gpurun_out/region_n$N.err
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "n$N exit $?" >> gpurun_out/bench_n$N.err
cat gpurun_out/region_n$N.json; tail -3 gpurun_out/region_n$N.err
python - <<PY
import json
try:
    r = json.load(open("gpurun_out/bench_n$N.json"))
    print("value", r["value"], "e2e", r["e2e"]["value"])
    print(json.dumps(r["detail"].get("sharded"), indent=1)[:6000])
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/bench_n$N.err
