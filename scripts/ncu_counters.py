"""ncu `--page raw --csv` export -> profiles/counters.json: per-kernel, per-launch hardware counters that bench.py turns into roofline
fractions together with the LIVE kernel durations it measures itself (a number taken under ncu is never a bench value).

    ncu --set full --clock-control none --import-source on -k regex:k_ -o gpurun_out/r02 python scripts/prof_workloads.py
    ncu -i gpurun_out/r02.ncu-rep --page raw --csv > gpurun_out/r02_raw.csv
    python scripts/ncu_counters.py gpurun_out/r02_raw.csv profiles/counters.json "note on the capture"
For a kernel captured several times the LAST launch of each workload is kept (earlier ones include first-touch effects)."""
import csv
import json
import re
import sys

KEYS = {
    "duration_us_under_ncu": ("gpu__time_duration.sum", 1e3),            # ms -> us (unit handled below)
    "warp_instructions": ("smsp__inst_executed.sum", 1),
    "thread_instructions": ("smsp__thread_inst_executed.sum", 1),
    "lanes_active_per_instruction": ("smsp__thread_inst_executed_per_inst_executed.ratio", 1),
    "issue_active_pct_of_active_cycles": ("smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
    "sm_cycles_active_avg": ("smsp__cycles_active.avg", 1),
    "sm_cycles_elapsed_max": ("sm__cycles_elapsed.max", 1),
    "l1_lsu_wavefronts": ("l1tex__data_pipe_lsu_wavefronts.sum", 1),
    "l1_lsu_wavefronts_pct_of_peak": ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 1),
    "l1_sectors": ("l1tex__t_sectors.sum", 1),
    "l1_hit_pct": ("l1tex__t_sector_hit_rate.pct", 1),
    "l2_sectors": ("lts__t_sectors.sum", 1),
    "l2_hit_pct": ("lts__t_sector_hit_rate.pct", 1),
    "dram_read_bytes": ("dram__bytes_read.sum", 1),
    "dram_write_bytes": ("dram__bytes_write.sum", 1),
    "pipe_alu_pct": ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1),
    "pipe_fma_pct": ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1),
    "pipe_lsu_pct": ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1),
    "pipe_fp64_pct": ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", 1),
    "warps_active_pct": ("sm__warps_active.avg.pct_of_peak_sustained_active", 1),
    "registers_per_thread": ("launch__registers_per_thread", 1),
    "grid": ("launch__grid_size", 1),
    "block": ("launch__block_size", 1),
}
UNIT_SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3, "second": 1e6}


def short(name):
    m = re.match(r"(?:void\s+)?([\w:]+)(<[^>(]*>)?", name.strip())
    return (m.group(1) + (m.group(2) or "")) if m else name[:40]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    H, U = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(H)}
    out = {}
    for r in rows[2:]:
        k = short(r[idx["Kernel Name"]])
        if not k.startswith("k_"):
            continue
        d = {}
        for key, (col, _) in KEYS.items():
            if col not in idx or r[idx[col]] in ("", "n/a"):
                continue
            v = float(r[idx[col]].replace(",", ""))
            u = U[idx[col]]
            if key.endswith("bytes") or key.startswith("duration"):
                v *= UNIT_SCALE.get(u, 1.0)
            d[key] = v
        if "l2_sectors" in d:
            d["l2_bytes"] = d["l2_sectors"] * 32
        if "l1_sectors" in d:
            d["l1_bytes"] = d["l1_sectors"] * 32
        if "dram_read_bytes" in d:
            d["dram_bytes"] = d["dram_read_bytes"] + d.get("dram_write_bytes", 0.0)
        k = f"{k}#{int(d.get('grid', 0))}"            # one entry per (kernel, grid size): the same kernel serves several workloads ...
        base, v = k, 1
        while k in out and abs(out[k].get("warp_instructions", 0.0) - d.get("warp_instructions", 0.0)) > 0.1 * max(out[k].get("warp_instructions", 1.0), 1.0):
            v += 1; k = f"{base}~{v}"                 # ... and one per workload when the grid is the same (k_icp_loop: C2, then C4 as `~2`), in order of appearance
        d["launches_captured"] = out.get(k, {}).get("launches_captured", 0) + 1
        out[k] = d
    meta = {"source": sys.argv[1], "note": sys.argv[3] if len(sys.argv) > 3 else "", "per": "launch (last captured launch of each kernel)"}
    json.dump({"meta": meta, "kernels": out}, open(sys.argv[2], "w"), indent=1, sort_keys=True)
    for k, d in out.items():
        print(f"{k:32s} {d.get('duration_us_under_ncu', 0):10.1f} us  inst {d.get('warp_instructions', 0):.3g}  lanes {d.get('lanes_active_per_instruction', 0):.1f}  "
              f"L1wf {d.get('l1_lsu_wavefronts_pct_of_peak', 0):.0f}%  dram {d.get('dram_bytes', 0) / 1e6:.1f} MB")


if __name__ == "__main__":
    main()
