"""Condense an `ncu --page raw --csv` export to one line per captured launch with the roofline counters.
usage: python tools/ncu_summary.py raw.csv [> profiles/<name>_summary.txt]"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "t_us", 1e-3),
    ("dram__bytes_read.sum", "dram_rd_MB", 1e-6),
    ("dram__bytes_write.sum", "dram_wr_MB", 1e-6),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct", 1),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct", 1),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct", 1),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_cyc_pct", 1),
    ("sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active", "dmma_pct", 1),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_cyc_pct", 1),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_wave_pct", 1),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem", 1),
    ("launch__grid_size", "grid", 1),
]


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.reader(lines))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    name_i = col.get("Kernel Name")
    avail = [(k, lab, sc) for k, lab, sc in KEYS if k in col]
    print("# " + path)
    print("# kernel | " + " | ".join(lab for _, lab, _ in avail))
    for r in data:
        vals = []
        for k, lab, sc in avail:
            v = num(r[col[k]])
            u = units[col[k]]
            if v is not None and k == "gpu__time_duration.sum":
                v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
                sc = 1
            if v is not None and k.startswith("dram__bytes"):
                v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
                sc = 1
            vals.append("-" if v is None else f"{v * sc:.4g}")
        print(r[name_i][:64].ljust(64) + " | " + " | ".join(vals))


if __name__ == "__main__":
    main(sys.argv[1])
