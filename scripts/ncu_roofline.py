#!/usr/bin/env python3
"""Regenerate one entry of profiles/traffic.json — what bench.py quotes under `roofline` — from an ncu report.

    python scripts/ncu_roofline.py gpurun_out/prof.ncu-rep cfg2 [--launch 0] [--name r02_draw_fast_cfg2]

Nothing is typed by hand: the kernel name, the executed FP32 operation counts, the pipe / issue utilisation and the
DRAM traffic all come out of the `.ncu-rep` (ncu --set full --import-source on, see scripts/gpu_prof.sh):

  executed FP32 flops per launch = sum over the SASS of the launch of "Predicated-On Thread Instructions Executed" x
      {FADD, FMUL: 1; FFMA: 2; FADD2, FMUL2: 2; FFMA2: 4}            (MUFU, FSETP, FMNMX and integer work count 0)
  cross-check: smsp__sass_thread_inst_executed_op_{fadd,fmul,ffma}_pred_on (raw page, per cycle x cycles elapsed)

bench.py divides that count by the kernel time it measures live and by the FFMA peak it measures live: roofline.frac.
The entry also records a hash of the kernel sources at capture time, so a bench line can say when the capture is stale.
With --name the raw-metric subset, the per-line instruction mix and the opcode mix are written next to it under profiles/.
"""
import argparse
import collections
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLOPS = {"FADD": 1, "FMUL": 1, "FFMA": 2, "FADD2": 2, "FMUL2": 2, "FFMA2": 4}


def source_hash() -> str:
    """sha256 over the CUDA sources of the package (what a capture is valid for)."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "uob_raytracer_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh", ".h")):
            with open(os.path.join(d, name), "rb") as f:
                h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def ncu_page(rep: str, page: str, *extra: str) -> list:
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(out.splitlines()))


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("workload", help="key in profiles/traffic.json: cfg2, cfg2_strict, cfg4, ...")
    ap.add_argument("--launch", type=int, default=0, help="which launch of the report (default: the first)")
    ap.add_argument("--name", default=None, help="also write profiles/<name>_{raw.csv,opcodes.txt,lines.txt}")
    args = ap.parse_args()

    raw = ncu_page(args.rep, "raw")
    hdr, units, data = raw[0], raw[1], raw[2:]
    row = dict(zip(hdr, data[args.launch]))
    unit = dict(zip(hdr, units))

    def num(k, default=None):
        if k not in row or row[k] in ("", "n/a"):
            return default
        v = float(row[k].replace(",", ""))
        return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e3, "ms": 1e6}.get(unit[k], 1)  # bytes / ns

    # --- SASS page: per-instruction executed counts of that launch ---
    sass = ncu_page(args.rep, "source", "--print-source", "sass")
    launch, cols, ops_thread, ops_warp = -1, None, collections.Counter(), collections.Counter()
    for r in sass:
        if not r:
            continue
        if r[0] == "Kernel Name":
            launch += 1
            continue
        if r[0] == "Address":
            cols = {c: i for i, c in enumerate(r)}
            continue
        if launch != args.launch or cols is None or not r[0].startswith("0x"):
            continue
        toks = r[cols["Source"]].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        ops_warp[op] += int(r[cols["Instructions Executed"]])
        ops_thread[op] += int(r[cols["Predicated-On Thread Instructions Executed"]])
    flop = sum(ops_thread[o] * w for o, w in FLOPS.items())
    warp_inst = sum(ops_warp.values())

    # cross-check against the hardware counters of the raw page
    cyc = num("smsp__cycles_elapsed.max") or num("sm__cycles_elapsed.max") or num("sm__cycles_elapsed.avg")
    hw = None
    keys = [f"smsp__sass_thread_inst_executed_op_{o}_pred_on.sum.per_cycle_elapsed" for o in ("fadd", "fmul", "ffma")]
    if cyc and all(k in row for k in keys):
        fa, fm, ff = (num(k) * cyc for k in keys)
        hw = fa + fm + 2 * ff

    entry = {
        "kernel": row["Kernel Name"],
        "source_hash": source_hash(),
        "report": os.path.basename(args.rep),
        "duration_us_under_ncu": round(num("gpu__time_duration.sum") / 1e3, 2),
        "executed_fp32_flop_per_launch": int(flop),
        "executed_fp32_flop_hw_counters": None if hw is None else int(hw),
        "thread_inst": {o: int(ops_thread[o]) for o in FLOPS if ops_thread[o]},
        "warp_instructions": int(warp_inst),
        "fp32_share_of_warp_instructions": round(sum(ops_warp[o] for o in FLOPS) / max(warp_inst, 1), 4),
        "dram_bytes_read": int(num("dram__bytes_read.sum", 0)),
        "dram_bytes_write": int(num("dram__bytes_write.sum", 0)),
        "ncu": {
            "issue_active_pct": round(num("smsp__issue_active.avg.pct_of_peak_sustained_active", 0), 2),
            "fma_pipe_pct": round(num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 0), 2),
            "alu_pipe_pct": round(num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 0), 2),
            "xu_pipe_pct": round(num("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 0), 2),
            "lsu_pipe_pct": round(num("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 0), 2),
            "warps_active_pct": round(num("sm__warps_active.avg.pct_of_peak_sustained_active", 0), 2),
            "registers": int(num("launch__registers_per_thread", 0)),
            "threads_per_warp_instruction": round(num("smsp__thread_inst_executed_per_inst_executed.ratio", 0), 2),
            "dram_throughput_pct": round(num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0), 3),
            "l2_throughput_pct": round(num("lts__throughput.avg.pct_of_peak_sustained_elapsed", 0), 3),
            "l1_throughput_pct": round(num("l1tex__throughput.avg.pct_of_peak_sustained_active", 0), 3),
            "stall_no_instruction_per_issue": round(num("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", 0), 3),
            "stall_barrier_per_issue": round(num("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", 0), 3),
            "stall_short_scoreboard_per_issue": round(num("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", 0), 3),
            "stall_long_scoreboard_per_issue": round(num("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", 0), 3),
        },
    }
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            table = json.load(f)
    except Exception:
        table = {}
    table[args.workload] = entry
    with open(path, "w") as f:
        json.dump(table, f, indent=1)
        f.write("\n")

    if args.name:
        base = os.path.join(ROOT, "profiles", args.name)
        with open(base + "_opcodes.txt", "w") as f:
            f.write(f"{entry['kernel']}\nwarp instructions {warp_inst}, executed FP32 flop {flop}\n")
            for op, n in ops_warp.most_common(40):
                f.write(f"{op:10s} {n:12d} {100 * n / warp_inst:6.2f}%  thr/inst {ops_thread[op] / max(n, 1):5.1f}\n")
        subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "save_profile.py"), os.path.splitext(os.path.basename(args.rep))[0], args.name],
                       cwd=ROOT, check=False)
    print(json.dumps(entry, indent=1))
    return 0


if __name__ == "__main__":
    sys.exit(main())
