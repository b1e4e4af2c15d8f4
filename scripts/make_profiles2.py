"""Round-2 measurement artefacts from gpurun_out/ (scratch) into profiles/ (tracked).

    ncu -i gpurun_out/prof_r02_step.ncu-rep --page raw --csv > /tmp/step_raw.csv
    ncu -i gpurun_out/prof_r02_gemm.ncu-rep --page raw --csv > /tmp/gemm_raw.csv
    python scripts/make_profiles2.py
"""
import csv
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_config_size",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9}
N_CAPTURE = 99991          # particles of the captured scene (scripts/profile_contact.py 100000)

out_rows, traffic = [], {}
hdr_out = None
for raw in ("/tmp/step_raw.csv", "/tmp/gemm_raw.csv"):
    if not os.path.exists(raw):
        continue
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    keep = [m for m in METRICS if m in hdr]
    if hdr_out is None:
        hdr_out = ["kernel"] + [f"{m} [{units[hdr.index(m)]}]" for m in keep]
        keep_out = keep
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        out_rows.append([name] + [r[hdr.index(m)] if m in hdr else "" for m in keep_out])
        rd = float(r[hdr.index("dram__bytes_read.sum")]) * SCALE[units[hdr.index("dram__bytes_read.sum")]]
        wr = float(r[hdr.index("dram__bytes_write.sum")]) * SCALE[units[hdr.index("dram__bytes_write.sum")]]
        key = name.split("<")[0]
        if key not in traffic:
            traffic[key] = {"bytes_per_launch_n100k": rd + wr, "read": rd, "write": wr,
                            "source": "ncu --set full, r02_ncu_full_summary.csv (first profiled launch of the kernel, scene of %d particles)" % N_CAPTURE}
            if key in ("k_deform_t", "k_deform_fin", "k_force_c"):
                traffic[key]["bytes_per_particle"] = (rd + wr) / N_CAPTURE
with open(os.path.join(P, "r02_ncu_full_summary.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(hdr_out)
    w.writerows(out_rows)
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
lines = open(os.path.join(G, "r02_launches_n1.csv")).read().splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
open(os.path.join(P, "r02_launches_n1.csv"), "w").write("\n".join(lines[start:]) + "\n")
for name in ("r02_pytest_2gpu.log", "r02_launches_rebuild_n1m.csv"):
    if os.path.exists(os.path.join(G, name)):
        shutil.copy(os.path.join(G, name), os.path.join(P, name))
print(open(os.path.join(P, "traffic.json")).read())
