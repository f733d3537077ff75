"""Summarise an `ncu --page raw --csv` export: one line per profiled launch with the metrics the
roofline needs (duration, DRAM bytes, DRAM/SM throughput %, occupancy, registers, tensor pipe)."""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "dur_us"),
    ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "inst"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_pct"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
]


def to_float(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return None


def main(path, flt=None):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    units = rows[1]
    name_i = hdr.index("Kernel Name")
    cols = {}
    for k, short in KEYS:
        if k in hdr:
            cols[short] = hdr.index(k)
    print("kernel," + ",".join(cols))
    for r in rows[2:]:
        if len(r) <= name_i:
            continue
        nm = r[name_i].split("(")[0].replace("void ssd::", "").replace("ssd::", "")
        if flt and flt not in nm:
            continue
        out = []
        for short, i in cols.items():
            v = to_float(r[i])
            u = units[i]
            if v is None:
                out.append("")
                continue
            if short == "dur_us":
                v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
            if short.endswith("_MB"):
                v = v / 1e6 if u == "byte" else v / 1e3 if u == "Kbyte" else v * 1e3 if u == "Gbyte" else v
            out.append(f"{v:.1f}" if isinstance(v, float) and abs(v) < 1e7 else f"{v:.3g}")
        print(nm[:48] + "," + ",".join(out))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
