"""Pull the counters the roofline uses out of `ncu -i X.ncu-rep --page raw --csv` -> JSON (one record per kernel).
usage: ncu -i rep --page raw --csv | python tools/ncu_extract.py > profiles/xxx.json"""
import csv
import json
import sys

rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
want = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__inst_executed.avg.per_cycle_elapsed": "ipc_per_sm",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_busy_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_active_pct",
    "l1tex__data_pipe_lsu_wavefronts.sum": "lsu_wavefronts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "lsu_wavefronts_shared",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "shared_bank_conflict_wavefronts",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "lsu_wavefront_pipe_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_lanes_per_instruction",
}
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0, "msecond": 1e-3,
         "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}
out = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    rec = {"kernel": r[ix["Kernel Name"]].split("(")[0].replace("void ", "")}
    for m, name in want.items():
        if m in ix and r[ix[m]] != "":
            v = float(r[ix[m]].replace(",", ""))
            u = units[ix[m]]
            if u in scale and ("byte" in u or "second" in u or u in ("us", "ms", "ns", "s")):
                v *= scale[u]
                name_u = name + ("_s" if "byte" not in u else "")
            else:
                name_u = name
            rec[name_u] = v
    if "dram_bytes_read" in rec and "duration_s" in rec:
        rec["dram_GBps"] = (rec["dram_bytes_read"] + rec.get("dram_bytes_write", 0.0)) / rec["duration_s"] / 1e9
    out.append(rec)
json.dump(out, sys.stdout, indent=1)
