#!/bin/bash
# compute-sanitizer over smoke() and the block-coder / stage parity tests on a B200; logs -> gpurun_out/
# usage (under gpurun): bash tools/gpu_sanitize.sh
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  echo "== $tool: smoke" > gpurun_out/sanitize_$tool.log
  timeout 900 $CS --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/sanitize_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitize_$tool.log
  echo "== $tool: block coders + stage parity (tests/test_stage_parity.py -m gpu)" >> gpurun_out/sanitize_$tool.log
  timeout 1500 $CS --tool $tool --print-limit 20 python -m pytest tests/test_stage_parity.py -m gpu -q -x >> gpurun_out/sanitize_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitize_$tool.log
  echo "== $tool: packet headers on the device, Part 2 structures (general DWT kernels)" >> gpurun_out/sanitize_$tool.log
  timeout 1500 $CS --tool $tool --print-limit 20 python -m pytest tests/test_packet_headers_device.py tests/test_part2_structures.py -m gpu -q -x -k "not random_configs" >> gpurun_out/sanitize_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitize_$tool.log
done
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|rc=|passed|failed" gpurun_out/sanitize_*.log
