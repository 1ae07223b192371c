#!/usr/bin/env python3
"""torchrun entry: row-region sharding over NCCL on real GPUs (the configurations of tests/test_regions.py, which the
CPU tier runs under the emulator over gloo, plus two larger ones).  Rank 0 checks the gathered codestream against the
reference's (oracle/_ref) and the decoded samples against the reference's decode; exit code 1 on any difference."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
import openjph_b200 as ob
from openjph_b200 import sharding, _lib
import cases, refharness, test_regions

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
L = _lib.lib(); assert L.ojb_set_device(local) == 0
CASES = [kw for kw, _ in test_regions.REGION_CASES] + [
    dict(width=2048, height=3000, num_comps=3, bit_depth=12, num_decomps=5, reversible=True, color_transform=True),
    dict(width=1920, height=2160, num_comps=3, bit_depth=10, num_decomps=5, reversible=False, color_transform=True, qstep=0.002),
]
res, all_ok = [], True
for ci, kw in enumerate(CASES):
    p = cases.make(kw)
    frame = cases.frame_for(p)
    sh = sharding.NativeShard()
    sh.set_partition("regions")
    sh.configure(p, ob.I32, writer=0)
    ok = True
    for rep in range(2):
        fr = frame if rep == 0 else [np.ascontiguousarray(a[::-1]) for a in frame]
        t0 = time.perf_counter()
        cs = sh.encode(fr)
        t1 = time.perf_counter()
        want = None
        if rank == 0:
            # reversible: the reference's codestream.  Irreversible 9/7 (the reference's own ISA variants differ in the last
            # bit) and DFS (the reference has no writer): the one-GPU encoder is the yardstick -- the
            # regions must reproduce it byte for byte -- and the reference is held to its tests' tolerance on decode
            if p.reversible and p.dfs_num_levels == 0:
                want = refharness.encode(p, fr)
            else:
                e1 = ob.Encoder(p, ob.I32)
                want = e1.encode(fr); e1.close()
            ok = ok and cs == want
        planes = sh.decode(want, sample_type=ob.I32, writer=0)
        if rank == 0:
            d1 = ob.Decoder()
            one_planes = d1.decode(want, ob.I32); d1.close()
            ok = ok and all(np.array_equal(a, b) for a, b in zip(planes, one_planes))
            ref_planes, _ = refharness.decode(want)
            err = max(int(np.abs(a.astype(np.int64) - b).max()) for a, b in zip(planes, ref_planes))
            ok = ok and err <= (0 if p.reversible else 1)
    if rank == 0:
        res.append(dict(case=ci, w=p.width, h=p.height, reversible=bool(p.reversible), identical=bool(ok), bytes=len(cs),
                        encode_ms=round((t1 - t0) * 1e3, 2), rows=sh.region_rows(0)))
        all_ok = all_ok and ok
    sh.close()
if rank == 0:
    print(json.dumps(dict(ranks=world, ok=bool(all_ok), cases=res)))
dist.destroy_process_group()
sys.exit(0 if rank != 0 or all_ok else 1)
