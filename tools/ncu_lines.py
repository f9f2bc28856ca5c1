"""Aggregate `ncu --page source --csv` output per source line: where do the warp-state samples of a kernel fall?
usage: ncu -i rep.ncu-rep --page source --csv | python tools/ncu_lines.py [top]"""
import csv
import sys
from collections import defaultdict

top = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rows = list(csv.reader(sys.stdin))
hdr = None
agg = defaultdict(lambda: defaultdict(float))
for r in rows:
    if hdr is None:
        if any(c.startswith("Warp Stall Sampling") or c == "Source" for c in r) and "Source" in r:
            hdr = r
        continue
    if len(r) != len(hdr):
        if any(c.startswith("Warp Stall Sampling") for c in r):
            hdr = r
        continue
    d = dict(zip(hdr, r))
    key = d.get("Source", "?")
    for k, v in d.items():
        if k.startswith("stall_") or k.startswith("Warp Stall Sampling") or k in ("# Samples", "Instructions Executed"):
            try:
                agg[key][k] += float(v.replace(",", ""))
            except ValueError:
                pass
samp = [c for c in (hdr or []) if c.startswith("Warp Stall Sampling (All")]
skey = samp[0] if samp else "# Samples"
tot = sum(v.get(skey, 0) for v in agg.values())
print("columns:", [c for c in (hdr or [])][:40])
print("total samples", tot)
for key, v in sorted(agg.items(), key=lambda kv: -kv[1].get(skey, 0))[:top]:
    stalls = sorted(((x, k) for k, x in v.items() if k.startswith("stall_")), reverse=True)[:4]
    print("%6.2f%%  exec %-9d %s   %s" % (100 * v.get(skey, 0) / max(tot, 1), int(v.get("Instructions Executed", 0)),
                                       key[:110], [(k.replace("stall_", ""), int(x)) for x, k in stalls]))
