"""Copy the measurement artefacts a round is judged on from gpurun_out/ (scratch) into profiles/ (tracked).

    python scripts/make_profiles.py r01

Inputs (produced by scripts/gpu_job_profile.sh and scripts/gpu_job_multi.sh under gpurun): bench_j.json, launches_j.csv, raw_j.csv (= `ncu -i
prof_r01_j.ncu-rep --page raw --csv`), configs_j.jsonl, bench_n*.json."""
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_config_size",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def last_json_line(path):
    lines = [l for l in open(path).read().splitlines() if l.startswith("{")]
    return json.loads(lines[-1])


os.makedirs(P, exist_ok=True)
bench = last_json_line(os.path.join(G, "bench_j.json"))
json.dump(bench, open(os.path.join(P, f"{tag}_bench_n1.json"), "w"), indent=1)
# launch list: keep the csv part only
lines = open(os.path.join(G, "launches_j.csv")).read().splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
open(os.path.join(P, f"{tag}_launches_n1.csv"), "w").write("\n".join(lines[start:]) + "\n")
# ncu --set full summary: one row per profiled launch
rows = list(csv.reader(open(os.path.join(G, "raw_j.csv"))))
hdr, units = rows[0], rows[1]
keep = [m for m in METRICS if m in hdr]
with open(os.path.join(P, f"{tag}_ncu_full_summary.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel"] + [f"{m} [{units[hdr.index(m)]}]" for m in keep])
    traffic = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        w.writerow([name] + [r[hdr.index(m)] for m in keep])
        scale = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9}
        rd = float(r[hdr.index("dram__bytes_read.sum")]) * scale[units[hdr.index("dram__bytes_read.sum")]]
        wr = float(r[hdr.index("dram__bytes_write.sum")]) * scale[units[hdr.index("dram__bytes_write.sum")]]
        key = name.split("<")[0]
        traffic.setdefault(key, {"bytes_per_launch_n100k": rd + wr, "read": rd, "write": wr,
                                 "source": f"ncu --set full, {tag}_ncu_full_summary.csv (first profiled launch of the kernel)"})
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
# HBM-bound kernels (cell binning / radix sort, stand-alone integrate, export) at n = 1e6: raw_build.csv from scripts/gpu_job_build_profile.sh
braw = os.path.join(G, "raw_build.csv")
if os.path.exists(braw):
    rows = list(csv.reader(open(braw)))
    hdr, units = rows[0], rows[1]
    col = lambda r, m: r[hdr.index(m)]
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9}
    tscale = {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}
    with open(os.path.join(P, f"{tag}_ncu_build_n1m.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "time_us", "dram_read_MB", "dram_write_MB", "achieved_GBps", "dram_pct_of_peak", "l2_pct_of_peak", "grid", "block"])
        for r in rows[2:]:
            t = float(col(r, "gpu__time_duration.sum")) * tscale[units[hdr.index("gpu__time_duration.sum")]]
            rd = float(col(r, "dram__bytes_read.sum")) * scale[units[hdr.index("dram__bytes_read.sum")]]
            wr = float(col(r, "dram__bytes_write.sum")) * scale[units[hdr.index("dram__bytes_write.sum")]]
            w.writerow([col(r, "Kernel Name").split("(")[0].replace("void ", ""), round(t * 1e6, 2), round(rd / 1e6, 2), round(wr / 1e6, 2),
                        round((rd + wr) / t / 1e9, 1), col(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                        col(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"), col(r, "launch__grid_size"), col(r, "launch__block_size")])
shutil.copy(os.path.join(G, "configs_j.jsonl"), os.path.join(P, f"{tag}_configs.jsonl"))
with open(os.path.join(P, f"{tag}_scaling.jsonl"), "w") as f:
    for name in sorted(os.listdir(G)):
        if name.startswith("bench_n") and name.endswith(".json"):
            try:
                d = last_json_line(os.path.join(G, name))
            except Exception:
                continue
            d["run"] = name[len("bench_"):-len(".json")]
            f.write(json.dumps(d) + "\n")
print("profiles written:", sorted(os.listdir(P)))
