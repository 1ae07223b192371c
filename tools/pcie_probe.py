#!/usr/bin/env python3
"""pinned-memory PCIe probe: H2D alone, D2H alone, both directions at once (GB/s)"""
import time, torch
n = 402653184
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1 = torch.cuda.Stream(); s2 = torch.cuda.Stream()
def run(f, reps=5):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both(): h2d(); d2h()
print("H2D  %.1f GB/s" % (n / run(h2d) / 1e9))
print("D2H  %.1f GB/s" % (n / run(d2h) / 1e9))
t = run(both); print("both %.1f GB/s each direction (%.1f total)" % (n / t / 1e9, 2 * n / t / 1e9))
# split in 4 chunks on 4 streams per direction
ss = [torch.cuda.Stream() for _ in range(8)]
def both4():
    q = n // 4
    for i in range(4):
        with torch.cuda.stream(ss[i]): d1[i*q:(i+1)*q].copy_(h1[i*q:(i+1)*q], non_blocking=True)
        with torch.cuda.stream(ss[4+i]): h2[i*q:(i+1)*q].copy_(d2[i*q:(i+1)*q], non_blocking=True)
t = run(both4); print("both, 4 streams each: %.1f GB/s each direction" % (n / t / 1e9))
import subprocess
print(subprocess.run("nvidia-smi topo -m | head -8; numactl -H 2>/dev/null | head -6; cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c | head", shell=True, capture_output=True, text=True).stdout)
