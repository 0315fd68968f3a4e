"""summarise one kernel of an `ncu --set full` report: key counters + stall ratios -> csv (profiles/)."""
import csv
import json
import subprocess
import sys

rep, out_csv = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, unit, val = rows[0], rows[1], rows[2]
keep = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__icc_request_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "smsp__inst_executed.sum", "lts__t_bytes.sum"]
keep += [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
d = {}
with open(out_csv, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "value", "unit"])
    for h in keep:
        if h in hdr:
            i = hdr.index(h)
            w.writerow([h, val[i], unit[i]])
            d[h] = (val[i], unit[i])


def to_bytes(k):
    v, u = d[k]
    return float(v.replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]


print(json.dumps({"dram_bytes": int(to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")),
                  "time": d["gpu__time_duration.sum"], "dram_pct": d["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"][0],
                  "hmma_pct": d.get("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", ("", ""))[0]}))
