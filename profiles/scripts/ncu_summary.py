"""Summarise an `ncu --page raw --csv` dump: one line per kernel with the metrics used in DESIGN.md / profiles/."""
import csv, sys
KEYS = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "rd",
    "dram__bytes_write.sum": "wr",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram%",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ%",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "lts__t_sector_hit_rate.pct": "L2hit%",
    "l1tex__t_sector_hit_rate.pct": "L1hit%",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "st_long_sb",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "st_barrier",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio": "st_membar",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "st_short_sb",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "st_lg",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "st_math",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "st_wait",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio": "st_sleep",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64%",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm%",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue%",
}
for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    hdr = rows[0]; units = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]][:60]
        out = [f"{name:60s}"]
        def val(k):
            if k not in idx: return None
            try: return float(r[idx[k]].replace(",", ""))
            except Exception: return None
        t = val("gpu__time_duration.sum"); tu = units[idx["gpu__time_duration.sum"]]
        t_us = t / 1e3 if tu in ("ns", "nsecond") else (t if tu.startswith("us") else t * 1e3)
        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        def tob(v, u):
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        rdb, wrb = tob(rd, units[idx["dram__bytes_read.sum"]]), tob(wr, units[idx["dram__bytes_write.sum"]])
        out.append(f"time {t_us:8.1f} us  dram rd {rdb/1e6:8.1f} MB wr {wrb/1e6:8.1f} MB => {(rdb+wrb)/t_us/1e3:7.1f} GB/s")
        for k, lab in KEYS.items():
            if lab in ("time", "rd", "wr"): continue
            v = val(k)
            if v is not None: out.append(f"{lab} {v:.2f}")
        print("  ".join(out))
