"""Key metrics of every launch in an .ncu-rep as JSON: python tools/ncu_key_metrics.py report.ncu-rep > profiles/x.json"""
import csv, io, json, subprocess, sys
KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ci = {h: i for i, h in enumerate(hdr)}
out = {"report": sys.argv[1].split("/")[-1], "launches": []}
for r in rows[2:]:
    rec = {"kernel": r[ci["Kernel Name"]]}
    for k in KEYS:
        if k in ci:
            rec[k] = {"value": r[ci[k]], "unit": units[ci[k]]}
    out["launches"].append(rec)
print(json.dumps(out, indent=1))
