"""Top source lines by stall samples from an ncu report captured with --import-source on:
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:NAME > src.csv
   python tools/hotlines.py src.csv [top]
Aggregates the per-source-line rows (the rows that carry a line number) over all files/launches of the export."""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path, errors="replace")))
cur_file, hdr, agg = None, None, {}
stall_cols = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        ix_smp = hdr.index("# Samples")
        ix_inst = hdr.index("Instructions Executed")
        stall_cols = {i: h for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
        continue
    if hdr is None or r[0] == "":
        continue
    try:
        line = int(r[0])
        smp = float(r[ix_smp] or 0)
        inst = float(r[ix_inst] or 0)
    except ValueError:
        continue
    key = (cur_file, line)
    a = agg.setdefault(key, {"src": r[1].strip(), "smp": 0.0, "inst": 0.0, "st": {}})
    a["smp"] += smp
    a["inst"] += inst
    for i, h in stall_cols.items():
        try:
            a["st"][h] = a["st"].get(h, 0.0) + float(r[i] or 0)
        except (ValueError, IndexError):
            pass
tot_s = sum(a["smp"] for a in agg.values()) or 1.0
tot_i = sum(a["inst"] for a in agg.values()) or 1.0
print("warp instructions %.3e, stall samples %d" % (tot_i, tot_s))
allst = {}
for a in agg.values():
    for h, v in a["st"].items():
        allst[h] = allst.get(h, 0.0) + v
print("stall mix: " + ", ".join("%s %.1f%%" % (h[6:], 100 * v / tot_s) for h, v in sorted(allst.items(), key=lambda t: -t[1])[:8]))
for (f, line), a in sorted(agg.items(), key=lambda t: -t[1]["smp"])[:top]:
    st = sorted(a["st"].items(), key=lambda t: -t[1])[:2]
    print("%-16s %4d %5.1f%% inst %5.1f%% smp [%s] | %s" % (f, line, 100 * a["inst"] / tot_i, 100 * a["smp"] / tot_s,
          ",".join("%s %.0f%%" % (h[6:], 100 * v / max(a["smp"], 1)) for h, v in st), a["src"][:110]))
