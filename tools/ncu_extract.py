#!/usr/bin/env python
"""tools/ncu_extract.py — reads an `ncu --set full` report here (no GPU needed) and prints one markdown row per captured
launch: duration, DRAM bytes and throughput, tensor-pipe activity, registers, grid, executed warp instructions and the top
warp-stall reasons.

    python tools/ncu_extract.py gpurun_out/r02_awq_decode.ncu-rep [more.ncu-rep ...] > profiles/r02_ncu_full.md
"""
import csv
import io
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram read",
    "dram__bytes_write.sum": "dram write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram % of peak",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor pipe %",
    "sm__inst_executed_pipe_tensor.sum": "tensor instr",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__inst_executed.sum": "warp instr",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy %",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm %",
}


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    header, units, data = r[0], r[1], r[2:]
    for d in data:
        yield dict(zip(header, d)), dict(zip(header, units))


def fmt(v, unit):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    if unit in ("byte", "Kbyte", "Mbyte", "Gbyte"):
        x *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        return "%.2f MB" % (x / 1e6)
    if unit in ("ns", "us", "usecond", "nsecond", "ms", "msecond"):
        x *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}[unit]
        return "%.1f us" % x
    if unit == "%":
        return "%.1f" % x
    return "%g" % x


def main():
    for path in sys.argv[1:]:
        print("## %s\n" % path)
        cols = list(WANT.values())
        print("| kernel | " + " | ".join(cols) + " | top stalls (warp samples %) |")
        print("|---|" + "---|" * (len(cols) + 1))
        for row, units in rows_of(path):
            name = row.get("Kernel Name", "?")
            cells = [fmt(row[k], units.get(k, "")) if k in row else "" for k in WANT]
            stalls = []
            for k, v in row.items():
                if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
                    try:
                        stalls.append((float(v.replace(",", "")), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                    except ValueError:
                        pass
            tot = sum(s for s, _ in stalls) or 1.0
            top = ", ".join("%s %.0f" % (n, 100 * s / tot) for s, n in sorted(stalls, reverse=True)[:4])
            print("| `%s` | %s | %s |" % (name[:70], " | ".join(cells), top))
        print()


if __name__ == "__main__":
    main()
