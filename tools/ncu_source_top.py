"""Top stall locations of one kernel of an .ncu-rep captured with --import-source on.
usage: ncu_source_top.py REP KERNEL_ID|- [N]   ("-": the report holds one kernel)"""
import csv, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + ([] if kid == "-" else ["--kernel-id", ":::" + kid]), capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(h) and r[h.index("# Samples")].isdigit()]
print(rows[0][1][:120] if rows[0] else "")
isrc, isam, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r[isam] or 0) for r in data)
print("total samples", tot, "sass lines", len(data), "warp-instr", sum(int(r[iex] or 0) for r in data))
agg = {}
for r in data:
    for i, c in stall:
        agg[c] = agg.get(c, 0) + int(r[i] or 0)
print("stall mix:", ", ".join("%s %.1f%%" % (c[6:], 100 * v / max(tot, 1)) for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for r in sorted(data, key=lambda r: -int(r[isam] or 0))[:n]:
    st = sorted([(int(r[i] or 0), c[6:]) for i, c in stall], reverse=True)[:2]
    print("%6s %5.1f%% ex=%9s  %-72s %s" % (r[isam], 100 * int(r[isam]) / max(tot, 1), r[iex], r[isrc][:72], st))
