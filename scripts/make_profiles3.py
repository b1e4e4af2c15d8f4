"""Round-2 (final) measurement artefacts from gpurun_out/ (scratch) into profiles/ (tracked).

    python scripts/make_profiles3.py

Reads the .ncu-rep files with `ncu -i ... --page raw --csv` (ncu is in the image; no GPU needed):
  prof_r02_step      ncu --set full, one step of the configs[1] workload (99 991 particles + obstacle)   scripts/gpu_job_final2.sh
  prof_r2w_step1m    ncu --set full, k_deform_t + k_force_p at n = 999 934                                scripts/gpu_job_r2w.sh
  prof_r02_build     ncu --set full, k_tile_walk_bits + k_tile_expand<0|1> at n = 999 934                 scripts/gpu_job_final2.sh
The k_sdf_gemm row (16 384 rows) of the earlier capture is kept.
"""
import csv
import io
import json
import os
import shutil
import subprocess

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
           "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9}
TIME = {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}
CAPTURES = [("prof_r02_step", "configs[1] step, n = 99991 + obstacle", 99991),
            ("prof_r2w_step1m", "n = 999934, no obstacle", 999934),
            ("prof_r02_build", "neighbour build, n = 999934", 999934)]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    k = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    return rows[k], rows[k + 1], rows[k + 2:]


header = ["kernel", "capture"] + METRICS
out_rows, traffic = [], {}
old = os.path.join(P, "traffic.json")
if os.path.exists(old):
    traffic = {k: v for k, v in json.load(open(old)).items() if k in ("k_sdf_gemm",)}
for name, what, n_cap in CAPTURES:
    rep = os.path.join(G, name + ".ncu-rep")
    if not os.path.exists(rep):
        print("missing", rep)
        continue
    hdr, units, rows = raw(rep)
    for r in rows:
        kname = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        vals = []
        for m in METRICS:
            if m not in hdr:
                vals.append(""); continue
            v, u = r[hdr.index(m)], units[hdr.index(m)]
            if m == "gpu__time_duration.sum":
                v = "%.3f" % (float(v) * TIME.get(u, 1.0))                    # always us
            elif m.startswith("dram__bytes"):
                v = "%.3f" % (float(v) * SCALE.get(u, 1.0) / 1e6)             # always MB
            vals.append(v)
        out_rows.append([kname, what] + vals)
        rd = float(r[hdr.index("dram__bytes_read.sum")]) * SCALE[units[hdr.index("dram__bytes_read.sum")]]
        wr = float(r[hdr.index("dram__bytes_write.sum")]) * SCALE[units[hdr.index("dram__bytes_write.sum")]]
        key = kname.split("<")[0]
        if key not in traffic:
            traffic[key] = {"bytes_per_launch": rd + wr, "read": rd, "write": wr, "particles_of_capture": n_cap,
                            "source": "ncu --set full, profiles/r02_ncu_full_summary.csv (%s; first profiled launch of the kernel)" % what}
            if key in ("k_deform_t", "k_deform_fin", "k_force_c", "k_force_p", "k_integrate"):
                traffic[key]["bytes_per_particle"] = (rd + wr) / n_cap
# keep the k_sdf_gemm row of the earlier capture
prev = os.path.join(P, "r02_ncu_full_summary.csv")
if os.path.exists(prev):
    rows = list(csv.reader(open(prev)))
    ph = [c.split(" [")[0] for c in rows[0]]
    for r in rows[1:]:
        if r and r[0].startswith("k_sdf_gemm") and "capture" not in ph:
            out_rows.append([r[0], "16384-row bulk query (earlier round-2 capture)"] + [r[ph.index(m)] if m in ph else "" for m in METRICS])
        elif r and r[0].startswith("k_sdf_gemm"):
            out_rows.append(r)
with open(prev, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(header[:2] + [m + (" [us]" if m == "gpu__time_duration.sum" else " [MB]" if m.startswith("dram__bytes") else "") for m in METRICS])
    w.writerows(out_rows)
json.dump(traffic, open(old, "w"), indent=1)

for src, dst in (("r02_launches_n1.csv", "r02_launches_n1.csv"), ("r2u_launches_rebuild_n1m.csv", "r02_launches_rebuild_n1m.csv")):
    p = os.path.join(G, src)
    if os.path.exists(p):
        lines = open(p).read().splitlines()
        start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
        open(os.path.join(P, dst), "w").write("\n".join(lines[start:]) + "\n")
for name in ("r02_pytest_gpu.log", "r02_pytest_2gpu.log", "r02_bench_n1.json"):
    if os.path.exists(os.path.join(G, name)):
        shutil.copy(os.path.join(G, name), os.path.join(P, name))
print(json.dumps({k: {"bytes_per_particle": v.get("bytes_per_particle"), "MB": v["bytes_per_launch"] / 1e6 if "bytes_per_launch" in v else None} for k, v in traffic.items()}, indent=1))
for r in out_rows:
    print(r[0][:28], r[1][:30], r[2], "us")
