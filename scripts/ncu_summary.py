"""Summarise an .ncu-rep (ncu --set full) into a small JSON for profiles/: per kernel the
duration, DRAM traffic, throughput percentages, occupancy, instruction count and top stalls.
    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/out.json"""
import csv, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
res = []
for r in rows[2:]:
    d = {"kernel": r[idx["Kernel Name"]]}
    for k in keep:
        if k in idx:
            d[k] = f"{r[idx[k]]} {units[idx[k]]}".strip()
    stalls = {h.split("issue_stalled_")[1].split("_per_issue")[0]: float(r[i].replace(",", "") or 0)
              for h, i in idx.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("per_issue_active.ratio")}
    d["top_stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:5])
    res.append(d)
json.dump(res, open(out, "w"), indent=1)
print(out, len(res), "kernels")
