#!/usr/bin/env python3
"""static SASS instructions per source line of one kernel (no GPU needed): cuobjdump -xelf + nvdisasm -g on the object file.
usage: sass_lines.py <object.o> <kernel name fragment> [top N]   (complements tools/ncu_lines.py, which gives EXECUTED counts)"""
import collections, glob, os, re, subprocess, sys, tempfile
obj, frag = os.path.abspath(sys.argv[1]), sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, capture_output=True)
    cubin = glob.glob(os.path.join(d, "*.cubin"))[0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
sec, cur, cnt, name = False, None, collections.Counter(), None
for l in txt.split("\n"):
    m = re.match(r"//-+ \.text\.(\S+) -+", l)
    if m:
        if sec:
            break
        if frag in m.group(1):
            sec, name = True, m.group(1)
        continue
    if not sec:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
        cnt[cur] += 1
src = {}
def line_of(f, n):
    p = os.path.join(os.path.dirname(obj), "..", f)
    if f not in src:
        try: src[f] = open(p).read().split("\n")
        except OSError: src[f] = []
    return src[f][n - 1].strip()[:110] if 0 < n <= len(src[f]) else ""
tot = sum(cnt.values())
print("# %s\n# %d static SASS instructions; per source line (file:line count share text)" % (name, tot))
for (f, n), c in cnt.most_common(top):
    print("%-18s %5d %5.1f%%  %s" % ("%s:%d" % (f, n), c, 100.0 * c / tot, line_of(f, n)))
