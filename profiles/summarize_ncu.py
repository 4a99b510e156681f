"""Summarise .ncu-rep captures (run on the build box: `python profiles/summarize_ncu.py gpurun_out/x.ncu-rep ...`).
Prints one markdown block per kernel launch with the counters the roofline discussion in DESIGN.md uses."""
import csv, io, subprocess, sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
]
for rep in sys.argv[1:]:
    if rep.endswith(".csv"):     # `ncu -i x.ncu-rep --page raw --csv` exported on the GPU box
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        print(f"## {rep}: no data"); continue
    hdr, units = rows[0], rows[1]
    print(f"## {rep}\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"### `{name[:140]}`\n")
        for k, label in KEYS:
            if k in hdr and r[hdr.index(k)] not in ("", "n/a"):
                print(f"* {label}: {r[hdr.index(k)]} {units[hdr.index(k)]}")
        print()
