"""Text summary of one `ncu --set full` report: headline metrics of every captured kernel + top stall reasons.
usage: python tools/ncu_summary.py report.ncu-rep > profiles/xxx_summary.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
STALL = "smsp__average_warps_issue_stalled_"
print(f"# {rep}: ncu --set full --clock-control none (cold caches, serialised: read shares and ratios, not absolutes)")
for d in data:
    print()
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:90s} {d[i]} {units[i]}")
    st = sorted(((float(d[i].replace(",", "") or 0), h[len(STALL):].replace("_per_issue_active.ratio", "")) for i, h in enumerate(hdr)
                 if h.startswith(STALL) and h.endswith("_per_issue_active.ratio")), reverse=True)[:7]
    print("stall reasons (warps per issue-active cycle): " + ", ".join(f"{n} {v:.2f}" for v, n in st))
