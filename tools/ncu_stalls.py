"""Summarise an `ncu --page source --csv` export: stall samples by SASS opcode and the hottest instructions.
usage: python tools/ncu_stalls.py <src.csv> [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_")]
by_op, by_reason, tot = collections.Counter(), collections.Counter(), 0
ins = []
for k, r in enumerate(rows[2:]):
    if len(r) < len(hdr): continue
    n = int(r[ix["# Samples"]] or 0)
    tot += n
    op = r[ix["Source"]].split()
    op = next((w for w in op if not w.startswith("@")), "?")
    by_op[op.split(".")[0] + ("." + op.split(".")[1] if "." in op else "")] += n
    reasons = {c: int(r[ix[c]] or 0) for c in stall_cols if (r[ix[c]] or "0") != "0"}
    for c, v in reasons.items(): by_reason[c] += v
    ins.append((n, k, r[ix["Source"]], reasons, int(r[ix["Instructions Executed"]] or 0)))
print("total samples", tot)
print("by reason:", ", ".join(f"{c[6:]} {v} ({100*v/tot:.1f}%)" for c, v in by_reason.most_common(12)))
print("by opcode:", ", ".join(f"{o} {v} ({100*v/tot:.1f}%)" for o, v in by_op.most_common(16)))
print("total warp instructions executed", sum(i[4] for i in ins))
for n, k, src, reasons, ex in sorted(ins, reverse=True)[:top]:
    print(f"{n:6d} #{k:5d} x{ex:9d} {src[:60]:60s} {dict(sorted(reasons.items(), key=lambda kv: -kv[1])[:3])}")
