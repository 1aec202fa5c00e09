"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line."""
import csv, sys, collections
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur_file = None
agg = collections.OrderedDict()
hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No":
        hdr = r; i_inst = hdr.index("Instructions Executed"); i_samp = hdr.index("# Samples"); i_tinst = hdr.index("Thread Instructions Executed"); continue
    if hdr is None: continue
    if r[0] != "" and r[0].isdigit():
        key = (cur_file, int(r[0]), r[1].strip())
        try:
            inst, samp, tinst = int(r[i_inst]), int(r[i_samp]), int(r[i_tinst])
        except ValueError:
            continue
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += inst; a[1] += samp; a[2] += tinst
tot = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print("total warp-inst %d, samples %d" % (tot, ts))
for (f, ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%6.2f%% inst %6.2f%% samp  thr/inst %4.1f  %s:%d  %s" % (100 * a[0] / tot, 100 * a[1] / max(ts, 1), a[2] / max(a[0], 1), f, ln, src[:110]))
