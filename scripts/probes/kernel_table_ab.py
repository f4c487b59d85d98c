#!/usr/bin/env python
"""Side-by-side per-kernel ms/step of two bench.py JSON lines. usage: kernel_table_ab.py a.json b.json"""
import json
import sys
a, b = (json.loads(open(p).read().strip().splitlines()[-1]) for p in sys.argv[1:3])
ka, kb = a["kernel_ms_per_step"], b["kernel_ms_per_step"]
print("%-48s %8s %8s" % ("kernel", sys.argv[1][-12:], sys.argv[2][-12:]))
for k in sorted(set(ka) | set(kb), key=lambda k: -max(ka.get(k, 0), kb.get(k, 0)))[:22]:
    print("%-48s %8.3f %8.3f" % (k, ka.get(k, 0), kb.get(k, 0)))
print("%-48s %8.3f %8.3f" % ("ms_per_step", a["ms_per_step"], b["ms_per_step"]))
