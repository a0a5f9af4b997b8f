"""Summarise an `ncu --page source --csv` dump: hottest SASS instructions by
stall samples, and samples per opcode."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) < len(hdr) - 5: continue
    try:
        data.append((int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]]), r[ix["Source"]].strip(), float(r[ix["Avg. Threads Executed"]] or 0), r))
    except ValueError:
        pass
tot = sum(d[0] for d in data); toti = sum(d[1] for d in data)
print("total samples", tot, "total warp instr", toti, "instructions in kernel", len(data))
byop = collections.Counter(); byopi = collections.Counter()
for s, n, src, thr, r in data:
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    byop[op] += s; byopi[op] += n
print("by opcode (samples%, instr%):")
for op, s in byop.most_common(18):
    print("  %-10s %5.1f%%  %5.1f%%" % (op, 100.0 * s / tot, 100.0 * byopi[op] / toti))
# stall reason columns
stall_cols = [h for h in hdr if h.startswith("stall_")]
if stall_cols:
    agg = collections.Counter()
    for s, n, src, thr, r in data:
        for h in stall_cols:
            try: agg[h] += int(r[ix[h]])
            except ValueError: pass
    print("stall reasons:", [(h, v) for h, v in agg.most_common(8)])
print("hottest instructions:")
for i, (s, n, src, thr, r) in enumerate(sorted(data, key=lambda d: -d[0])[:top]):
    print("  %5.2f%%  n=%9d thr=%4.1f  %s" % (100.0 * s / tot, n, thr, src[:90]))
