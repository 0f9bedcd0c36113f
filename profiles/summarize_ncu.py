"""Turn an .ncu-rep (brought back in gpurun_out/) into the markdown table kept under profiles/.
usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/<name>.md"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs/thread"),
    ("smsp__inst_executed.sum", "warp instr"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes/instr"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"ncu --set full --clock-control none, report `{path.split('/')[-1]}` (cold caches, kernels serialised)\n")
    print("| kernel | " + " | ".join(n for _, n in METRICS) + " |")
    print("|---|" + "---|" * len(METRICS))
    for r in data:
        name = r[idx["Kernel Name"]].split("::")[-1].split("(")[0]
        cells = []
        for m, _ in METRICS:
            if m in idx:
                v, u = r[idx[m]], units[idx[m]]
                try:
                    v = f"{float(v.replace(',', '')):.4g}"
                except ValueError:
                    pass
                cells.append(f"{v} {u}".strip())
            else:
                cells.append("-")
        print(f"| {name} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
