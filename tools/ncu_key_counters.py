"""Key counters of one kernel from an `ncu --page raw --csv` export: python tools/ncu_key_counters.py RAW.csv [kernel substring] > out.json"""
import csv
import json
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_l1tex2xbar_write_bytes.sum", "sm__cycles_elapsed.max"]
STALLS = "smsp__average_warp_latency_issue_stalled_"      # (ncu: warp cycles per issued instruction, by reason)

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
row = next(r for r in rows[2:] if want in r[hdr.index("Kernel Name")])
out = {"kernel": row[hdr.index("Kernel Name")]}
for k in KEYS:
    if k in hdr:
        i = hdr.index(k)
        out[k] = "%s %s" % (row[i], units[i])
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        out["stall_" + h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(float(row[i] or 0), 2)
print(json.dumps(out, indent=1))
