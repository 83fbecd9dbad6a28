"""Summarise an .ncu-rep (read here, no GPU): per kernel the metrics the roofline/judging needs.
  python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/NAME.md"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]  # an .ncu-rep, or the CSV that `ncu -i file.ncu-rep --page raw --csv` printed on the GPU box
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
raw = raw[raw.index('"ID"'):] if '"ID"' in raw else raw
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tma_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct"]
print("# ncu summary of `%s`\n" % rep.split("/")[-1])
print("Captured with `ncu --set full --clock-control none --import-source on` (B200, sm_100a). Durations under ncu are not bench values.\n")
seen = {}
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    seen.setdefault(name, []).append(r)
for name, rs in seen.items():
    r = rs[-1]
    print("## `%s`  (%d launches captured, last shown)\n" % (name[:150], len(rs)))
    print("| metric | value | unit |\n|---|---|---|")
    for w in want:
        if w in idx:
            print("| %s | %s | %s |" % (w, r[idx[w]], units[idx[w]]))
    try:
        rd = float(r[idx["dram__bytes_read.sum"]]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[idx["dram__bytes_read.sum"]]]
        wr = float(r[idx["dram__bytes_write.sum"]]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[idx["dram__bytes_write.sum"]]]
        dur = float(r[idx["gpu__time_duration.sum"]]) * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}[units[idx["gpu__time_duration.sum"]]]
        print("| **dram traffic per launch (read+write)** | %.4g | GB |" % ((rd + wr) / 1e9))
        print("| **dram GB/s under ncu** | %.1f | GB/s |" % ((rd + wr) / dur / 1e9))
    except Exception:
        pass
    stalls = []
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(r[idx[h]]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    print("\nTop stall reasons (warps per issue-active cycle): " + ", ".join("%s %.2f" % (n, v) for v, n in sorted(stalls, reverse=True)[:7]) + "\n")
