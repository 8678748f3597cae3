"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by CUDA source line (development aid)."""
import csv, sys, collections


def num(x):
    try:
        return int(x)
    except Exception:
        return 0


rows = list(csv.reader(open(sys.argv[1])))
hdr = None; fname = ""; agg = collections.OrderedDict(); cur = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 10: continue
    line, src = r[0], r[1]
    if line != "":
        cur = (fname, line, src.strip()[:100]); agg.setdefault(cur, collections.Counter())
    d = dict(zip(hdr[2:], r[2:]))
    if d.get("Address"):
        a = agg[cur]
        a["samples"] += num(d["# Samples"]); a["inst"] += num(d["Instructions Executed"])
        for k in ("stall_long_sb", "stall_barrier", "stall_short_sb", "stall_wait", "stall_branch_resolving", "stall_no_inst"):
            a[k] += num(d[k])
tot = sum(a["samples"] for a in agg.values()); ti = sum(a["inst"] for a in agg.values())
print("total samples", tot, "inst", ti)
for k, a in sorted(agg.items(), key=lambda x: -x[1]["samples"])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{100*a['samples']/tot:5.1f}% inst {100*a['inst']/ti:4.1f}% lsb {a['stall_long_sb']:5d} bar {a['stall_barrier']:5d} ssb {a['stall_short_sb']:5d} wait {a['stall_wait']:5d} | {k[0]}:{k[1]} {k[2]}")
