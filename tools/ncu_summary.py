"""Condense an `ncu --set full` report into the rows DESIGN.md / bench.py quote.

    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep profiles/r1_x_ncu_full_summary.csv [kernel-regex]

Reads the report with `ncu -i ... --page raw --csv` (per-launch metrics) and
`--page source --csv` (stall samples, instruction mix, shared-memory wavefronts)."""
import collections
import csv
import io
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
]


def page(rep, name, kernel):
    cmd = ["ncu", "-i", rep, "--page", name, "--csv"]
    if kernel:
        cmd += ["-k", f"regex:{kernel}"]
    return list(csv.reader(io.StringIO(subprocess.run(cmd, capture_output=True, text=True).stdout)))


def main():
    rep, out = sys.argv[1], sys.argv[2]
    kernel = sys.argv[3] if len(sys.argv) > 3 else None
    rows = page(rep, "raw", kernel)
    hdr, units = rows[0], rows[1]
    lines = []
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")]
        lines.append(["kernel", "", re.sub(r"\(.*", "", name)])
        for k in KEEP:
            if k in hdr:
                lines.append([k, units[hdr.index(k)], vals[hdr.index(k)]])
    src = page(rep, "source", kernel)
    starts = [n for n, r in enumerate(src) if r and r[0] == "Address"]
    if starts:
        h = src[starts[-1]]
        data = src[starts[-1] + 1:]
        ix = {c: i for i, c in enumerate(h)}

        def f(r, c):
            try:
                return float(r[ix[c]])
            except (ValueError, KeyError, IndexError):
                return 0.0

        tot = sum(f(r, "# Samples") for r in data) or 1.0
        stalls = {c: sum(f(r, c) for r in data) for c in h if c.startswith("stall_") and "Not Issued" not in c}
        for c, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]:
            lines.append([f"samples.{c}", "%", f"{100 * v / tot:.1f}"])
        mix = collections.Counter()
        for r in data:
            parts = r[ix["Source"]].strip().split()
            if not parts:
                continue
            op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
            mix[op.split(".")[0]] += f(r, "Instructions Executed")
        for op, v in mix.most_common(12):
            lines.append([f"inst.{op}", "warp-inst", f"{v:.0f}"])
        lines.append(["smem_wavefronts", "", f"{sum(f(r, 'L1 Wavefronts Shared') for r in data):.0f}"])
        lines.append(["smem_wavefronts_ideal", "", f"{sum(f(r, 'L1 Wavefronts Shared Ideal') for r in data):.0f}"])
    with open(out, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["metric", "unit", "value"])
        w.writerows(lines)
    print(f"{out}: {len(lines)} rows")


if __name__ == "__main__":
    main()
