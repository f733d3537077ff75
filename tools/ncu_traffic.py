"""Aggregate an `ncu --page raw --csv` export of ONE inference step into per-kind DRAM traffic:
    python tools/ncu_traffic.py gpurun_out/conv_<tag>_raw.csv profiles/r1_traffic.json
kinds: stemblock (stem_dwproj_kernel), irblock (irblock_mma_kernel, irblock_mma_grouped_kernel, conv_irblock_tcgen05_kernel),
dwproj, conv (conv_tcgen05_kernel + conv_igemm_kernel + conv_splitk_reduce_kernel), dw (depthwise3x3_kernel), chain, decode_nms."""
import csv
import json
import sys


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    ri, wi, ti = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    l1key = "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"
    li = hdr.index(l1key) if l1key in hdr else None

    def to_bytes(v, u):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

    out = {}
    for r in rows[2:]:
        if len(r) <= ti:
            continue
        nm = r[ki]
        kind = ("stemblock" if "stem_dwproj" in nm else "irblock" if "irblock" in nm else "dwproj" if "dwproj" in nm
                else "chain" if "conv_chain" in nm else "stem" if "stem_conv" in nm else "decode_nms" if "nms_" in nm
                else "conv" if ("conv_tcgen05" in nm or "conv_igemm" in nm or "splitk" in nm) else "dw" if "depthwise" in nm else None)
        if kind is None:
            continue
        e = out.setdefault(kind, {"launches": 0, "dram_bytes": 0.0, "time_ns": 0.0, "l1tex_pct_x_ns": 0.0})
        t_ns = float(r[ti].replace(",", "")) * {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(units[ti], 1)
        e["launches"] += 1
        e["dram_bytes"] += to_bytes(r[ri], units[ri]) + to_bytes(r[wi], units[wi])
        e["time_ns"] += t_ns
        if li is not None:
            e["l1tex_pct_x_ns"] += float(r[li].replace(",", "")) * t_ns
    for e in out.values():
        e["dram_bytes_per_launch"] = e["dram_bytes"] / max(e["launches"], 1)
        # time-weighted L1/TEX (= shared-memory pipe) throughput of the group, % of peak: what bounds the block kernels
        e["l1tex_throughput_pct"] = e.pop("l1tex_pct_x_ns") / max(e["time_ns"], 1.0) if li is not None else None
        e["source"] = "ncu --set full --clock-control none, one SSD300-MobileNetV2 B=32 step (" + src.split("/")[-1] + ")"
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
