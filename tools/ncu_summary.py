#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) into a small markdown table + JSON: one row per captured launch.
   python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_sweeps   -> writes .md and .json"""
import csv, io, json, subprocess, sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_pct"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("launch__registers_per_thread", "regs"),
    ("launch__block_size", "block"),
    ("launch__grid_size", "grid"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("smsp__pcsamp_warps_issue_stalled_long_scoreboard", "stall_long_sb"),
    ("smsp__pcsamp_warps_issue_stalled_short_scoreboard", "stall_short_sb"),
    ("smsp__pcsamp_warps_issue_stalled_wait", "stall_wait"),
    ("smsp__pcsamp_warps_issue_stalled_selected", "stall_selected"),
    ("smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "stall_math_throttle"),
    ("smsp__pcsamp_warps_issue_stalled_barrier", "stall_barrier"),
    ("smsp__pcsamp_warps_issue_stalled_branch_resolving", "stall_branch"),
    ("smsp__pcsamp_warps_issue_stalled_not_selected", "stall_not_selected"),
    ("smsp__pcsamp_sample_count", "samples"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k, short in KEYS:
            if k in hdr:
                i = hdr.index(k)
                try:
                    d[short] = float(r[i].replace(",", ""))
                except ValueError:
                    d[short] = r[i]
                d[short + "_unit"] = units[i]
        res.append(d)
    json.dump(res, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write(f"# ncu --set full summary of `{rep}` (per captured launch; times are under the profiler: cold cache, serialised)\n\n")
        for d in res:
            f.write(f"## {d['kernel']}\n\n| metric | value | unit |\n|---|---|---|\n")
            for k, short in KEYS:
                if short in d:
                    v = d[short]
                    f.write(f"| {k} | {v:,.3f} | {d.get(short + '_unit', '')} |\n" if isinstance(v, float) else f"| {k} | {v} | |\n")
            if "dram_read" in d and "dram_write" in d:
                def b(x, u):
                    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                t = b(d["dram_read"], d["dram_read_unit"]) + b(d["dram_write"], d["dram_write_unit"])
                f.write(f"| **dram traffic (read+write)** | {t/1e9:.4f} | GB |\n")
            f.write("\n")
    print("wrote", out + ".md", out + ".json")


if __name__ == "__main__":
    main()
