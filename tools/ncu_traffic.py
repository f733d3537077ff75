"""Aggregate an `ncu --page raw --csv` export of ONE inference step into per-kind DRAM traffic:
    python tools/ncu_traffic.py gpurun_out/conv_<tag>_raw.csv profiles/r1_traffic.json
kinds: conv (conv_tcgen05_kernel + conv_igemm_kernel + conv_splitk_reduce_kernel), dw (depthwise3x3_kernel)."""
import csv
import json
import sys


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    ri, wi, ti = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")

    def to_bytes(v, u):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

    out = {}
    for r in rows[2:]:
        if len(r) <= ti:
            continue
        nm = r[ki]
        kind = ("irblock" if "irblock" in nm else "dwproj" if "dwproj" in nm else "chain" if "conv_chain" in nm
                else "stemblock" if "stem_dwproj" in nm else "stem" if "stem_conv" in nm else "decode_nms" if "nms_" in nm
                else "conv" if ("conv_tcgen05" in nm or "conv_igemm" in nm or "splitk" in nm) else "dw" if "depthwise" in nm else None)
        if kind is None:
            continue
        e = out.setdefault(kind, {"launches": 0, "dram_bytes": 0.0, "time_ns": 0.0})
        e["launches"] += 1
        e["dram_bytes"] += to_bytes(r[ri], units[ri]) + to_bytes(r[wi], units[wi])
        e["time_ns"] += float(r[ti].replace(",", "")) * {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(units[ti], 1)
    for e in out.values():
        e["dram_bytes_per_launch"] = e["dram_bytes"] / max(e["launches"], 1)
        e["source"] = "ncu --set full --clock-control none, one SSD300-MobileNetV2 B=32 step (" + src.split("/")[-1] + ")"
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
