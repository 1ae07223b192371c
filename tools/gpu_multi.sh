#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-2}
if [ -z "$SKIP_REF" ]; then
timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?" >> gpurun_out/bench_ref.err
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "n$N exit $?" >> gpurun_out/bench_n$N.err
cat gpurun_out/bench_ref.json | cut -c1-1200; tail -3 gpurun_out/bench_ref.err; cat gpurun_out/bench_n$N.json | cut -c1-2500; tail -5 gpurun_out/bench_n$N.err
if [ -z "$SKIP_SHARD" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/shard_check.py > gpurun_out/shard_n$N.json 2> gpurun_out/shard_n$N.err; echo "shard exit $?" >> gpurun_out/shard_n$N.err
cat gpurun_out/shard_n$N.json; tail -3 gpurun_out/shard_n$N.err
fi
if [ -z "$SKIP_REGIONS" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/region_check.py > gpurun_out/region_n$N.json 2> gpurun_out/region_n$N.err; echo "region exit $?" >> gpurun_out/region_n$N.err
cat gpurun_out/region_n$N.json; tail -3 gpurun_out/region_n$N.err
fi
