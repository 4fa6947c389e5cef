#!/usr/bin/env python
"""Summarise .ncu-rep files (ncu --set full) as a markdown table: duration, DRAM traffic, hit rates, stalls.
Usage: python tools/ncu_summary.py report1.ncu-rep [report2.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdGB"), ("dram__bytes_write.sum", "wrGB"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__t_sector_hit_rate.pct", "L2hit%"),
        ("l1tex__t_sector_hit_rate.pct", "L1hit%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dsmem")]


def to_gb(v, unit):
    v = float(v.replace(",", ""))
    u = unit.lower()
    return v * {"byte": 1e-9, "kbyte": 1e-6, "mbyte": 1e-3, "gbyte": 1.0, "tbyte": 1e3}.get(u, 1e-9)


def main():
    print("| kernel | " + " | ".join(k for _, k in KEYS) + " | traffic GB | top stalls |")
    print("|---|" + "---|" * (len(KEYS) + 2))
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
            cells, traffic = [], 0.0
            for k, short in KEYS:
                if k not in hdr:
                    cells.append("-")
                    continue
                i = hdr.index(k)
                if short in ("rdGB", "wrGB"):
                    g = to_gb(r[i], units[i])
                    traffic += g
                    cells.append(f"{g:.3f}")
                elif short == "us":
                    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[i].lower().replace("second", "s").replace("usecond", "us"), 1.0)
                    scale = {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(units[i].lower(), scale)
                    cells.append(f"{float(r[i].replace(',', '')) * scale:.1f}")
                else:
                    try:
                        cells.append(f"{float(r[i].replace(',', '')):.1f}")
                    except ValueError:
                        cells.append(r[i])
            stalls = []
            for i, h in enumerate(hdr):
                if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                    try:
                        stalls.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            st = ", ".join(f"{n} {v:.1f}" for v, n in stalls[:3])
            print(f"| {name} | " + " | ".join(cells) + f" | {traffic:.3f} | {st} |")


if __name__ == "__main__":
    main()
