"""Print SASS (with exec counts) attributed to given source lines from a cuda,sass ncu dump.
usage: ncu_sass.py dump.csv file:line_lo-line_hi"""
import csv, sys
path, spec = sys.argv[1], sys.argv[2]
fname, rng = spec.split(":")
lo, hi = (int(x) for x in rng.split("-"))
rows = list(csv.reader(open(path)))
cur_file = None; hdr = None; cur_line = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No":
        hdr = r; i_inst = hdr.index("Instructions Executed"); continue
    if hdr is None: continue
    if r[0] != "" and r[0].isdigit():
        cur_line = int(r[0])
        if cur_file == fname and lo <= cur_line <= hi: print("---- %s:%d %s  [inst %s]" % (cur_file, cur_line, r[1].strip()[:100], r[i_inst]))
        continue
    if r[0] == "" and cur_file == fname and cur_line is not None and lo <= cur_line <= hi and len(r) > i_inst and r[3] not in ("...", "-"):
        print("      %-60s %s" % (r[3].strip(), r[i_inst]))
