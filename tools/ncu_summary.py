"""Summarise an .ncu-rep (ncu --set full) into a small table: one row per captured launch with the metrics that the
roofline discussion needs. Usage: python tools/ncu_summary.py rep.ncu-rep [out.md]"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "dur_us", 1e-3),
    ("dram__bytes_read.sum", "dram_rd_MB", 1e-6),
    ("dram__bytes_write.sum", "dram_wr_MB", 1e-6),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct", 1),
    ("lts__t_sectors.avg.pct_of_peak_sustained_elapsed", "l2_pct", 1),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct", 1),
    ("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "upipe_pct", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("launch__grid_size", "grid", 1),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct", 1),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_pct", 1),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    cols = ["kernel"] + [k[1] for k in KEYS]
    out.append("| " + " | ".join(cols) + " |")
    out.append("|" + "---|" * len(cols))
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")[:60]
        vals = []
        for key, _, scale in KEYS:
            i = idx.get(key)
            if i is None or r[i] == "":
                vals.append("")
                continue
            v = float(r[i].replace(",", ""))
            u = units[i]
            if key.startswith("gpu__time_duration") and u == "ns":
                v *= 1e-3
            elif key.startswith("gpu__time_duration") and u == "us":
                pass
            elif key.startswith("dram__bytes"):
                v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
            vals.append(f"{v:.1f}" if abs(v) < 1e5 else f"{v:.3g}")
        out.append("| " + " | ".join([name] + vals) + " |")
    text = "\n".join(out)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")


if __name__ == "__main__":
    main()
