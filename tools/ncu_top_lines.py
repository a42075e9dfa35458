"""Top stalled SASS instructions of an `ncu --page source --csv --print-source sass` export (sampling data)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = None
for i, r in enumerate(rows):
    if "Address" in r and "Source" in r:
        hdr, data = r, rows[i + 1:]
        break
if hdr is None:
    sys.exit("no source table")
col = {k: hdr.index(k) for k in ("Address", "Source", "# Samples", "stall_long_sb", "stall_barrier", "stall_math", "stall_wait", "stall_lg", "stall_short_sb", "stall_mio") if k in hdr}
data = [r for r in data if len(r) == len(hdr)]
tot = sum(int(r[col["# Samples"]] or 0) for r in data)
print("kernel:", rows[0][1] if len(rows[0]) > 1 else "?", " total samples", tot)
print("%8s %6s  %8s %8s %8s %8s %8s  %s" % ("samples", "share", "long_sb", "barrier", "math", "wait", "mio", "SASS"))
for r in sorted(data, key=lambda r: -int(r[col["# Samples"]] or 0))[:n]:
    g = lambda k: r[col[k]] if k in col else "-"
    print("%8s %5.1f%%  %8s %8s %8s %8s %8s  %s" % (g("# Samples"), 100.0 * int(g("# Samples") or 0) / max(tot, 1), g("stall_long_sb"), g("stall_barrier"),
                                                   g("stall_math"), g("stall_wait"), g("stall_mio"), g("Source")[:100]))
