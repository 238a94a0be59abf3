#!/usr/bin/env python
"""Condenses an Nsight Compute report into the few numbers the roofline argument needs.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers, blocks/SM)"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (shared mem, blocks/SM)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active % (of active cycles)"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "FP64 pipe active % (of elapsed)"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle / SMSP"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per instruction"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print("# ncu summary of `%s`\n" % path.split("/")[-1])
    print("Captured with `ncu --set full --clock-control none --import-source on` under gpurun "
          "(cold caches, serialised launches: use shares and ratios, not absolute times).\n")
    for r in rows[2:]:
        print("## %s\n" % r[col["Kernel Name"]])
        print("| metric | value | unit |\n|---|---|---|")
        for key, label in KEYS:
            if key in col:
                print("| %s | %s | %s |" % (label, r[col[key]], units[col[key]]))
        stalls = []
        for h, i in col.items():
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v >= 0.1:
                    stalls.append((v, h.split("issue_stalled_")[1].split("_per_issue")[0]))
        print("\nwarp stall reasons (warps per issue-active cycle): " +
              ", ".join("%s %.2f" % (n, v) for v, n in sorted(stalls, reverse=True)) + "\n")


def traffic(path, out_json):
    """Per-launch DRAM traffic (read + write bytes) of every profiled kernel -> JSON for bench.py."""
    import json
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    res = {}
    for r in rows[2:]:
        tot = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[col[key]]) * scale.get(units[col[key]], 1.0)
        name = r[col["Kernel Name"]].split("::")[-1].split("(")[0]
        res[name] = {"dram_bytes_per_launch": tot, "duration_us_under_ncu":
                     float(r[col["gpu__time_duration.sum"]]) *
                     {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(units[col["gpu__time_duration.sum"]], 1.0),
                     "source": path.split("/")[-1]}
    with open(out_json, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    if len(sys.argv) > 3 and sys.argv[2] == "--traffic-json":
        traffic(sys.argv[1], sys.argv[3])
    else:
        main(sys.argv[1])
