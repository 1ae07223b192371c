#!/bin/bash
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"ht_encode|ht_dec_step|dwt_fwd|dwt_inv" -c 13 \
   -o gpurun_out/prof_r01a python tools/profile_once.py > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/
