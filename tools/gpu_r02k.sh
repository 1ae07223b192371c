#!/bin/bash
# one gpurun call: smoke, GPU parity tests, the bench line, and the block-coder variants on frames with 1/8, 1/4 and 1/2
# of the headline frame's code-blocks (what one rank of a sharded image codes)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
for ph in 1024 2048 4096; do
  for v in default warpenc twostepdec; do
    unset OJB_BLOCK_ENCODER OJB_BLOCK_DECODER
    if [ $v = warpenc ]; then export OJB_BLOCK_ENCODER=warp; fi
    if [ $v = twostepdec ]; then export OJB_BLOCK_DECODER=twostep; fi
    echo "PH=$ph $v: $(PN=4 PH=$ph timeout 120 python tools/profile_once.py 2>&1 | tail -1)" >> gpurun_out/small_frames_ab.log
  done
done
unset OJB_BLOCK_ENCODER OJB_BLOCK_DECODER
tail -2 gpurun_out/smoke.log; tail -4 gpurun_out/pytest_gpu.log; cut -c1-1500 gpurun_out/bench.json; tail -2 gpurun_out/bench.err; cat gpurun_out/small_frames_ab.log
