"""ncu -i rep --page source --csv (SASS view)  ->  compact "address samples stall..." lines for the instructions that were
sampled; map addresses to source lines offline with `nvdisasm -g` on the same (deterministically built) cubin."""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr = None
out = []
base = None
for r in rows:
    if hdr is None:
        if "Address" in r:
            hdr = r
        continue
    if len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    if base is None:
        base = d.get("Address")
    samp = None
    for k in ("Warp Stall Sampling (All Samples)", "# Samples", "Sampling Data (All)"):
        if k in d:
            samp = d[k]
            break
    try:
        sv = float((samp or "0").replace(",", ""))
    except ValueError:
        sv = 0.0
    if sv <= 0:
        continue
    stalls = {k: d[k] for k in d if k.startswith("stall_") and d[k] not in ("0", "", "0.0")}
    out.append((sv, d.get("Address", "?"), d.get("Source", "")[:60], d.get("Instructions Executed", ""), stalls))
print("columns:", hdr)
print("kernel base address:", base)
print("sampled instructions:", len(out), "total samples:", sum(o[0] for o in out))
for sv, addr, src, ex, st in sorted(out, reverse=True)[:400]:
    print("%8.0f %s exec=%s | %s | %s" % (sv, addr, ex, src, " ".join("%s=%s" % (k.replace("stall_", ""), v) for k, v in st.items())))
