"""Top stall lines of one kernel from an `ncu --page source --csv` export (SASS view).
Usage: python tools/ncu_stalls.py <source.csv[.gz]> <kernel-name-substring> [topn]"""
import csv, gzip, sys
path, pat = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
op = gzip.open if path.endswith(".gz") else open
rows = list(csv.reader(op(path, "rt")))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name" and pat in rows[i][1]:
        hdr = rows[i + 1]
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            body.append(rows[j]); j += 1
        idx = {h: k for k, h in enumerate(hdr)}
        samp = idx["# Samples"]
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(int(r[samp] or 0) for r in body)
        print(f"kernel {rows[i][1][:80]}: {len(body)} SASS lines, {tot} samples")
        agg = {h: sum(int(r[idx[h]] or 0) for r in body) for h in stall_cols}
        print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
        for r in sorted(body, key=lambda r: -int(r[samp] or 0))[:topn]:
            st = sorted(((int(r[idx[h]] or 0), h) for h in stall_cols), reverse=True)[:2]
            print(f"{int(r[samp]):7d} {100*int(r[samp])/max(tot,1):5.1f}%  {r[idx['Source']].strip()[:70]:70s} {st}")
        break
    i += 1
