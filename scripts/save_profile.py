"""Summarise gpurun_out/<tag>.ncu-rep into profiles/: raw-metric subset, per-source-line instruction mix."""
import csv, os, subprocess, sys, json
tag, name = sys.argv[1], sys.argv[2]
rep = f"gpurun_out/{tag}.ncu-rep"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__t_bytes.sum', 'lts__t_bytes.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
keys += [k for k in hdr if 'issue_stalled' in k and k.endswith('per_issue_active.ratio')]
os.makedirs("profiles", exist_ok=True)
with open(f"profiles/{name}_raw.csv", "w") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(data))])
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            w.writerow([k, units[i]] + [d[i] for d in data])
d = dict(zip(hdr, data[0]))
def num(k):
    v = float(d[k]); u = units[hdr.index(k)]
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(u, 1)
summ = {"kernel": d["Kernel Name"], "duration_us_under_ncu": float(d["gpu__time_duration.sum"]) * (1e-3 if units[hdr.index("gpu__time_duration.sum")] == "ns" else 1),
        "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
        "registers": int(float(d["launch__registers_per_thread"])), "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
        "fma_pipe_pct": float(d["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]), "warp_inst": float(d["smsp__inst_executed.sum"])}
json.dump(summ, open(f"profiles/{name}_summary.json", "w"), indent=1)
cs = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
open(f"gpurun_out/{tag}_cs.csv", "w").write(cs)
out = subprocess.run([sys.executable, "scripts/ncu_lines.py", f"gpurun_out/{tag}_cs.csv", "60"], capture_output=True, text=True).stdout
open(f"profiles/{name}_lines.txt", "w").write(out)
print(json.dumps(summ))
