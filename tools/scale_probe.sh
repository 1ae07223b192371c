#!/bin/bash
# block-coder stage time against the number of code-blocks in one launch: one 8K frame, then 2 and 4 frames
# stacked as tiles of one canvas
PN=3 PH=8192 timeout 600 python tools/profile_once.py 2>&1 | tail -1
PN=3 PH=16384 PTILE=8192 timeout 600 python tools/profile_once.py 2>&1 | tail -1
PN=3 PH=32768 PTILE=8192 timeout 600 python tools/profile_once.py 2>&1 | tail -1
