#!/usr/bin/env python
"""Aggregates `ncu --page source --csv` into warp-stall samples by SASS opcode and by source line (developer tool)."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr)]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot, ex, cnt, per = collections.Counter(), collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
total = 0
for r in data:
    src = r[ix["Source"]].strip()
    op = (src.split()[0] if not src.startswith("@") else src.split()[1]).split(".")[0]
    s = int(r[ix["# Samples"]] or 0)
    tot[op] += s; ex[op] += int(r[ix["Instructions Executed"]] or 0); cnt[op] += 1; total += s
    for c in stall_cols:
        v = int(r[ix[c]] or 0)
        if v: per[op][c] += v
print("\n## warp-stall samples by SASS opcode (%d samples, %d static instructions)" % (total, len(data)))
print("%-10s %7s %12s %8s  top stall reasons" % ("opcode", "static", "executed", "samples"))
for op, s in tot.most_common(20):
    print("%-10s %7d %12d %8d  %s" % (op, cnt[op], ex[op], s, ", ".join(f"{k[6:]}:{v}" for k, v in per[op].most_common(4))))
# hottest individual instructions
print("\n## hottest instructions")
hot = sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:25]
for r in hot:
    print("%6s  %s" % (r[ix["# Samples"]], r[ix["Source"]].strip()[:100]))
# samples per barrier-delimited section of the SASS (program order): which PHASE of the kernel the time goes to
print("\n## samples by barrier-delimited SASS section (>= 1.5 % of all samples); ops = static count of the main opcodes")
seg, segs = [], []
for i, r in enumerate(data):
    seg.append(r)
    src = r[ix["Source"]].strip()
    if re.search(r"\b(BAR\.SYNC|BAR\.RED|UCGABAR_WAIT|SYNCS\.PHASECHK)", src):
        segs.append(seg); seg = []
if seg: segs.append(seg)
pos = 0
for sg in segs:
    smp = sum(int(r[ix["# Samples"]] or 0) for r in sg)
    if smp >= 0.015 * max(total, 1):
        ops = collections.Counter()
        why = collections.Counter()
        for r in sg:
            src = r[ix["Source"]].strip()
            op = (src.split()[0] if not src.startswith("@") else src.split()[1]).split(".")[0]
            ops[op] += 1
            for c in stall_cols:
                v = int(r[ix[c]] or 0)
                if v: why[c[6:]] += v
        print("  sass rows %5d-%5d  %6d samples (%4.1f %%)  ops: %s | stalls: %s" % (
            pos, pos + len(sg) - 1, smp, 100.0 * smp / total,
            " ".join(f"{k}:{v}" for k, v in ops.most_common(7)), " ".join(f"{k}:{v}" for k, v in why.most_common(4))))
    pos += len(sg)
