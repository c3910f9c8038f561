#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): python scripts/ncu_summary.py file.ncu-rep [more.ncu-rep ...]
Prints, per profiled launch, the metrics the roofline discussion in DESIGN.md cites."""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor", "sm__pipe_tensor_subpipe", "sm__pipe_tensor_op",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        idx = [i for i, h in enumerate(hdr) if any(w in h for w in WANT)]
        kn = hdr.index("Kernel Name")
        print(f"# {path}")
        for r in rows[2:]:
            print(f"## {r[kn][:110]}")
            for i in idx:
                if r[i] not in ("", "n/a"):
                    print(f"   {hdr[i]:78s} {units[i]:12s} {r[i]}")


if __name__ == "__main__":
    main()
