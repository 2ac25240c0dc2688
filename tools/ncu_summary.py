#!/usr/bin/env python
"""Summarises `ncu --set full` reports (.ncu-rep) as a markdown table: duration, DRAM bytes and throughput, pipe utilisation,
occupancy and the top stall reasons. Usage: python tools/ncu_summary.py out.md rep1.ncu-rep rep2.ncu-rep ..."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor(legacy) %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
]
STALL = "smsp__average_warps_issue_stalled_"


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, row = rows[0], rows[1], rows[2]
    return {h: (row[i], units[i]) for i, h in enumerate(hdr)}


def main():
    out_md, reps = sys.argv[1], sys.argv[2:]
    lines = ["| kernel | " + " | ".join(k[1] for k in KEYS) + " | DRAM GB/s | tcgen05 / TMA evidence | top stalls (warps per issue) |",
             "|---|" + "---|" * (len(KEYS) + 3)]
    for rep in reps:
        d = load(rep)
        name = d.get("Kernel Name", ("?", ""))[0].split("(")[0].replace("magpo::<unnamed>::", "").replace("void ", "")
        cells = []
        for k, _ in KEYS:
            v, u = d.get(k, ("", ""))
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.3g}" if abs(f) < 1e5 else f"{f:.4g}"
            except ValueError:
                pass
            cells.append(f"{v} {u}".strip())
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
        num = lambda k: float(d[k][0].replace(",", "")) * scale.get(d[k][1], 1.0) if k in d and d[k][0] else 0.0
        t = num("gpu__time_duration.sum")
        cells.append(f"{(num('dram__bytes_read.sum') + num('dram__bytes_write.sum')) / t / 1e9:.0f}" if t else "")
        ev = []
        for k, lab in (("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "TMEM/tensor-mem active %"),
                       ("smsp__inst_executed_pipe_uniform.sum", "uniform-pipe inst"),
                       ("sm__inst_executed_pipe_tma.sum", "TMA inst"), ("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "TMA ld bytes")):
            if k in d and d[k][0] not in ("", "0"):
                ev.append(f"{lab} {d[k][0]}")
        stalls = sorted(((float(v[0].replace(",", "")), h[len(STALL):].replace("_per_issue_active.ratio", ""))
                         for h, v in d.items() if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and v[0] not in ("", "n/a")),
                        reverse=True)[:3]
        lines.append(f"| `{name}` | " + " | ".join(cells) + " | " + "; ".join(ev) + " | " + ", ".join(f"{n} {x:.2f}" for x, n in stalls) + " |")
    open(out_md, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


main()
