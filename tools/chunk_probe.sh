#!/bin/bash
# DWT stage time against the rows per warp chunk (wave quantisation of the level-1 launches)
for r in ${ROWS:-64 44 56 68 100 136}; do
  OJB_DWT_CHUNK_ROWS=$r PN=3 PREV=${PREV:-1} timeout 600 python tools/profile_once.py 2>&1 | tail -1 > /tmp/cp.txt
  python - <<PY
import re
s=open('/tmp/cp.txt').read()
print("chunk_rows $r", re.findall(r"'dwt': ([0-9.]+)", s), re.findall(r"'dwt_inv': ([0-9.]+)", s))
PY
done
