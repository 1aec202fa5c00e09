"""SASS listing of one launch of an ncu report, with executed counts, threads per instruction and stall samples per instruction.
usage: ncu_sass_listing.py report.ncu-rep [launch] > profiles/<name>_sass.txt"""
import csv, subprocess, sys
rep = sys.argv[1]; want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True, check=True).stdout
launch, cols = -1, None
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == "Kernel Name":
        launch += 1
        if launch == want: print("#", r[1])
        continue
    if r[0] == "Address":
        cols = {c: i for i, c in enumerate(r)}; continue
    if launch != want or cols is None or not r[0].startswith("0x"): continue
    ex = int(r[cols["Instructions Executed"]]); th = int(r[cols["Predicated-On Thread Instructions Executed"]])
    smp = r[cols["Warp Stall Sampling (All Samples)"]] if "Warp Stall Sampling (All Samples)" in cols else "0"
    print(f" {int(r[0], 16) & 0xfffff:5x}  {r[cols['Source']]:70s} exec {ex:10d} thr/inst {th / ex if ex else 0:5.0f} samples {int(smp or 0):5d}")
