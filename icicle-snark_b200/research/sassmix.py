#!/usr/bin/env python3
"""Per-function SASS opcode histogram of a cubin / object / executable (research tool)."""
import collections, re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
pat = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 8
fn, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1); hist[fn] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        hist[fn][m.group(2)] += 1
for fn, h in hist.items():
    if pat and not re.search(pat, fn): continue
    print(fn, sum(h.values()), " ".join(f"{k}={v}" for k, v in h.most_common(top)))
