"""Summarise an .ncu-rep: one line per launch with the metrics the roofline needs.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--md]
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("Kernel Name", "kernel"),
    ("gpu__time_duration.sum", "us"),
    ("sm__cycles_elapsed.max", "cycles"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = [(short, hdr.index(full)) for full, short in WANT if full in hdr]
    print(" | ".join(s for s, _ in idx))
    for r in rows[2:]:
        vals = []
        for s, i in idx:
            v = r[i]
            if s == "kernel":
                v = v.split("(")[0][-48:]
            elif s in ("dram_rd", "dram_wr"):
                v = f"{float(v):.1f}{units[i]}"
            else:
                try:
                    v = f"{float(v):.1f}"
                except ValueError:
                    pass
            vals.append(v)
        print(" | ".join(vals))


if __name__ == "__main__":
    main()
